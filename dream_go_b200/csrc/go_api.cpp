// go_api.cpp -- C ABI (include/dg_go.h) over go_board.h.
#include <atomic>
#include <cmath>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/dg_go.h"
#include "go_board.h"
#include "thread_pool.h"

using dg::Board;

static inline Board* B(dg_board* b) { return reinterpret_cast<Board*>(b); }
static inline const Board* B(const dg_board* b) { return reinterpret_cast<const Board*>(b); }

static inline float f16_bits_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int s = 0;
            while (!(m & 0x400u)) { m <<= 1; ++s; }
            x = sign | ((uint32_t)(113 - s) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

extern "C" {

// Argument checks of the boundary: the reference's types make these states unrepresentable (`Point`, `Color`,
// `Transform`); here a caller can hand over any integer.
static inline bool on_board(int32_t point) { return point >= 0 && point < dg::N_POINTS; }
static inline bool is_color(int32_t color) { return color == dg::BLACK || color == dg::WHITE; }
static inline bool is_transform(int32_t t) { return t >= 0 && t < 8; }
static inline bool is_search(int32_t s) { return s == dg::STANDARD_SEARCH || s == dg::SCORING_SEARCH; }
// Outputs of a call with arguments that are not a colour / transform / search kind: all zero (no plane set, no legal move,
// no candidate), for the prior all -inf -- defined, and nothing is read or written out of bounds.

void dg_go_set_zobrist(const uint64_t* table) {
    if (table) dg::mutable_tables().load_zobrist(table);
    else dg::mutable_tables().default_zobrist();
}

dg_board* dg_board_new(float komi) {
    Board* b = new Board();
    b->init(komi);
    return reinterpret_cast<dg_board*>(b);
}
dg_board* dg_board_clone(const dg_board* board) { return reinterpret_cast<dg_board*>(new Board(*B(board))); }
void dg_board_copy(dg_board* dst, const dg_board* src) { *B(dst) = *B(src); }
void dg_board_free(dg_board* board) { delete B(board); }
void dg_board_set_komi(dg_board* board, float komi) { B(board)->komi = komi; }
float dg_board_komi(const dg_board* board) { return B(board)->komi; }
int32_t dg_board_count(const dg_board* board) { return B(board)->count; }
uint64_t dg_board_zobrist_hash(const dg_board* board) { return B(board)->hash; }
int32_t dg_board_to_move(const dg_board* board) { return B(board)->to_move(); }
int32_t dg_board_at(const dg_board* board, int32_t point) { return on_board(point) ? B(board)->color[point] : -1; }
int32_t dg_board_is_valid(const dg_board* board, int32_t color, int32_t point) {
    return on_board(point) && is_color(color) ? B(board)->is_valid(color, point) : 0;
}
void dg_board_place(dg_board* board, int32_t color, int32_t point) {
    if (on_board(point) && is_color(color)) B(board)->place(color, point);      // 361 (pass) is not played (self_play.rs:442-451)
}
int32_t dg_board_get_n_liberty(const dg_board* board, int32_t point) { return on_board(point) ? B(board)->n_liberty(point) : 0; }
int32_t dg_board_get_n_liberty_if(const dg_board* board, int32_t color, int32_t point) {
    if (!on_board(point) || !is_color(color)) return -1;
    return B(board)->color[point] ? -1 : B(board)->liberties_if(color, point);
}
int32_t dg_board_is_ladder_capture(const dg_board* board, int32_t color, int32_t point) {
    return on_board(point) && is_color(color) ? dg::is_ladder_capture(*B(board), color, point) : 0;
}
int32_t dg_board_is_ladder_escape(const dg_board* board, int32_t color, int32_t point) {
    return on_board(point) && is_color(color) ? dg::is_ladder_escape(*B(board), color, point) : 0;
}
int32_t dg_board_is_symmetric(const dg_board* board, int32_t transform) {
    if (!is_transform(transform)) return 0;
    const uint16_t* t = dg::tables().sym[transform];
    const Board* b = B(board);
    for (int p = 0; p < dg::N_POINTS; ++p)
        if (b->color[p] != b->color[t[p]]) return 0;
    return 1;
}
void dg_board_legal_moves(const dg_board* board, int32_t color, uint8_t* out) {
    const Board* b = B(board);
    for (int p = 0; p < dg::N_POINTS; ++p) out[p] = is_color(color) ? (uint8_t)b->is_valid(color, p) : 0;
}
int32_t dg_symmetry_apply(int32_t transform, int32_t point) {
    return is_transform(transform) && point >= 0 && point <= dg::PASS ? dg::tables().sym[transform][point] : -1;
}
int32_t dg_symmetry_inverse(int32_t transform) { return is_transform(transform) ? dg::tables().sym_inverse[transform] : -1; }

void dg_board_features_packed(const dg_board* board, int32_t to_move, int32_t symmetry, dg_packed_position* out, uint8_t* legal) {
    if (!is_color(to_move) || !is_transform(symmetry)) {
        memset(out, 0, sizeof(*out));
        if (legal) memset(legal, 0, dg::N_POINTS);
        return;
    }
    dg::features_v1(*B(board), to_move, symmetry, out->planes, &out->k_bits, legal);
    out->reserved = 0;
}

void dg_board_raw_position(const dg_board* board, int32_t to_move, int32_t symmetry, dg_raw_position* out) {
    if (!is_color(to_move) || symmetry < 0 || symmetry > 0xff) { memset(out, 0, sizeof(*out)); return; }      // (the symmetry byte carries flags, dg_engine.h)
    dg::raw_position(*B(board), to_move, symmetry, out);
}

void dg_board_features_f16(const dg_board* board, int32_t to_move, int32_t symmetry, uint16_t* out) {
    if (!is_color(to_move) || !is_transform(symmetry)) { memset(out, 0, sizeof(uint16_t) * 32 * dg::N_POINTS); return; }
    dg_packed_position pos;
    dg::features_v1(*B(board), to_move, symmetry, pos.planes, &pos.k_bits, nullptr);
    for (int p = 0; p < dg::N_POINTS; ++p) {
        uint32_t m = pos.planes[p];
        uint16_t* o = out + 32 * p;
        o[0] = (m & 1u) ? pos.k_bits : 0;
        o[1] = (m & 2u) ? pos.k_bits : 0;
        for (int c = 2; c < 32; ++c) o[c] = ((m >> c) & 1u) ? 0x3c00 : 0;
    }
}

void dg_go_extract_batch(const dg_board* const* boards, const uint8_t* to_move, const uint8_t* symmetry, int32_t count,
                         dg_packed_position* out, uint8_t* legal, int32_t threads) {
    int hw = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = hw > 0 ? hw : 1;
    std::atomic<int> next{0};
    std::function<void()> work = [&] {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= count) break;
            dg_board_features_packed(boards[i], to_move[i], symmetry ? symmetry[i] : 0, out + i, legal ? legal + (size_t)i * 361 : nullptr);   // (checks its arguments)
        }
    };
    if (threads <= 1 || count <= 1) { work(); return; }
    // one process-wide pool (created on first use, sized to the machine); callers take turns
    static std::mutex turn;
    static dg::Helpers* pool = new dg::Helpers(std::max(1, hw - 1));
    std::lock_guard<std::mutex> g(turn);
    pool->run(work);
}

int32_t dg_go_replay(float komi, const uint8_t* colors, const uint16_t* moves, int32_t n, dg_packed_position* features,
                     uint8_t* legal, uint64_t* hashes) {
    Board board;
    board.init(komi);
    for (int i = 0; i < n; ++i) {
        int c = colors[i];
        if (!is_color(c)) return -(i + 1);                       // like an illegal move: the replay stops at this ply
        uint8_t* lg = legal ? legal + (size_t)i * 361 : nullptr;
        if (features) {
            dg::features_v1(board, c, 0, features[i].planes, &features[i].k_bits, lg);
            features[i].reserved = 0;
        } else if (lg) {
            for (int p = 0; p < dg::N_POINTS; ++p) lg[p] = (uint8_t)board.is_valid(c, p);
        }
        if (moves[i] < dg::N_POINTS) {
            if (!board.is_valid(c, moves[i])) return -(i + 1);
            board.place(c, moves[i]);
        }
        if (hashes) hashes[i] = board.hash;
    }
    return n;
}

int32_t dg_board_is_scorable(const dg_board* board) { return dg::is_scorable(*B(board)); }

void dg_board_territory(const dg_board* board, uint8_t* out) { dg::territory_status(*B(board), out); }

void dg_board_benson(const dg_board* board, int32_t color, uint8_t* out) {
    if (!is_color(color)) { memset(out, 0, dg::N_POINTS); return; }
    dg::Bits alive, eyes;
    dg::benson(*B(board), color, alive, eyes);
    for (int p = 0; p < dg::N_POINTS; ++p) out[p] = alive.test(p) ? 1 : eyes.test(p) ? 2 : 0;
}

void dg_board_policy_candidates(const dg_board* board, int32_t to_move, int32_t search, const uint8_t* legal, uint8_t* out) {
    if (!is_color(to_move) || !is_search(search)) { memset(out, 0, dg::N_POINTS + 1); return; }
    const Board* b = B(board);
    uint8_t local_legal[dg::N_POINTS];
    if (!legal) {
        for (int p = 0; p < dg::N_POINTS; ++p) local_legal[p] = (uint8_t)b->is_valid(to_move, p);
        legal = local_legal;
    }
    dg::policy_candidates(*b, to_move, search, legal, out);
}

void dg_board_prior(const dg_board* board, int32_t to_move, int32_t search, const uint8_t* legal, const uint16_t* policy,
                    int32_t symmetry, float sum_to, float* prior) {
    const dg::Tables& T = dg::tables();
    const float NEG_INF = -std::numeric_limits<float>::infinity();
    if (!is_color(to_move) || !is_search(search) || !is_transform(symmetry)) { for (int i = 0; i < 368; ++i) prior[i] = NEG_INF; return; }
    uint8_t candidates[dg::N_POINTS + 1];
    dg_board_policy_candidates(board, to_move, search, legal, candidates);
    // policy_helper.rs:36-47: candidates start at 0, everything else (and the padding) at -inf
    for (int i = 0; i < 368; ++i) prior[i] = NEG_INF;
    for (int p = 0; p <= dg::N_POINTS; ++p) if (candidates[p]) prior[p] = 0.0f;
    // :54-72: on a symmetric board keep the smallest index of each orbit
    int syms[8], ns = 0;
    for (int t = 0; t < 8; ++t) if (dg_board_is_symmetric(board, t)) syms[ns++] = t;
    uint16_t rep[dg::N_POINTS + 1];
    for (int p = 0; p < dg::N_POINTS; ++p) {
        int best = p;                                       // Identity is always in the group
        for (int k = 0; k < ns; ++k) { int q = T.sym[syms[k]][p]; if (q < best) best = q; }
        rep[p] = (uint16_t)best;
        if (best != p) prior[p] = NEG_INF;
    }
    // :87-104 add_valid_candidates: un-transform the network's policy and fold it onto the representatives
    prior[361] += f16_bits_to_f32(policy[361]);
    const uint16_t* inv = T.sym[T.sym_inverse[symmetry]];
    for (int i = 0; i < dg::N_POINTS; ++i) prior[rep[inv[i]]] += f16_bits_to_f32(policy[i]);
    // :113-134 normalize_policy.  The reference sums the finite entries in 8 interleaved lanes that are then
    // added pairwise (asm/sum_finite.rs:23-57) and multiplies by a reciprocal (asm/normalize_finite.rs:23-40);
    // same order here, with an IEEE reciprocal where the reference uses the 12-bit `rcpps` estimate.
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int finite = 0;
    for (int i = 0; i < 368; ++i) if (std::isfinite(prior[i])) { lane[i & 7] += prior[i]; ++finite; }
    float sum = ((lane[0] + lane[1]) + (lane[2] + lane[3])) + ((lane[4] + lane[5]) + (lane[6] + lane[7]));
    if (sum < 1e-6f) {
        for (int i = 0; i < 368; ++i) if (std::isfinite(prior[i])) prior[i] = sum_to / (float)finite;
    } else {
        float recip = 1.0f / (sum / sum_to);
        for (int i = 0; i < 368; ++i) prior[i] *= recip;
    }
}

}  // extern "C"
