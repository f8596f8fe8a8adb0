// features.cu -- V1 feature planes computed on the device from raw stones.
//
// Replaces the host's per-point work of `features::V1::get_features` (src/libdg_go/utils/features.rs:154-250) and the
// 361 x `Board::is_valid` of the prior construction (pool/policy_helper.rs:39-43) for the engine's self-play path: the
// host sends 384 bytes per leaf (stone / visited masks, hashes, last moves, and the two ladder planes, which stay a
// sequential search on the host) and gets the evaluation plus the legal-move mask back.
//
// One CTA per position, one thread per board point, everything in shared memory:
//   1. chains by min-label propagation with pointer jumping (board_fast.rs keeps linked lists; here they are rebuilt),
//   2. one 361-bit liberty set per chain (atomicOr from the empty points) and the XOR of its stones' zobrist keys,
//   3. per point the 32 plane bits: liberties of stones, legality + liberties-after-playing for both colours
//      (board_fast.rs:216-243, 484-539: union of the joined chains' liberty sets, the point's empty neighbours and any
//      captured stones that touch the new chain), super-ko against the last 16 hashes (board.rs:132-141),
//   4. written as `dg_packed_position` (index = symmetry[p]) plus the legal mask, and expanded in the same kernel into
//      the tower's input rows (400 x 64 fp16 per position).
// Integer work end to end; tests/test_features_gpu.py checks it bit for bit against the host code and the oracle.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "layout.h"

namespace dg {

struct RawPosition {                 // == dg_raw_position (include/dg_engine.h)
    uint32_t black[12], white[12], visited[12], ladder_capture[12], ladder_escape[12];
    unsigned long long hash;
    unsigned long long hash_history[16];
    int16_t last_move[2];
    uint16_t k_bits;
    uint8_t to_move, symmetry;
};
static_assert(sizeof(RawPosition) == 384, "dg_raw_position layout");

__constant__ unsigned long long c_zobrist[2][361];   // [colour - 1][point]
__constant__ uint16_t c_symmetry[8][361];

__device__ __forceinline__ bool bit(const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }


// ---- candidate mask + symmetry representatives on the device ---------------------------------------------------------------
struct PlanScratch {                      // lives in the liberty-set storage once the planes are done (< 17 KB)
    uint16_t rlab[384];                   // region label (smallest point index) of a point, 0xffff = none
    uint16_t vch[361][4];                 // chains a region is vital to
    int chain_cnt[361];                   // vital regions per chain (this iteration)
    int size[361];                        // points per region
    int touch[361][4];                    // points of the region that touch candidate chain j
    uint8_t other[384], dead[384], seeded[384], reg_on[384], nv[384], chain_alive[384], veto[384];
    uint8_t eye[2][384];
    int symm[8];
};
static_assert(sizeof(PlanScratch) <= 361 * 12 * 4, "plan scratch must fit into the liberty-set storage");

// Benson's unconditional life for colour c (utils/benson.rs as restated in csrc/go_board.h: benson()): marks eye[t] for the
// points of the surviving vital regions.  All 384 threads call it; `flag` is a shared int.
__device__ void device_benson(PlanScratch* S, const uint8_t* col, const uint16_t* lab, const int (&nb)[4], int t, bool on, int c,
                              uint8_t* eye, int* flag) {
    const bool is_other = on && col[t] != c;
    bool near = false;
    if (on)
        for (int k = 0; k < 4; k++) near |= nb[k] >= 0 && col[nb[k]] == c;
    S->other[t] = is_other;
    S->dead[t] = is_other && !near;       // touches no stone of c: its whole region is vital to nobody
    S->seeded[t] = 0;
    S->reg_on[t] = 0;
    S->nv[t] = 0;
    S->veto[t] = 0;
    S->chain_alive[t] = on && col[t] == c;
    eye[t] = 0;
    __syncthreads();
    for (;;) {                            // flood the dead regions
        if (t == 0) *flag = 0;
        __syncthreads();
        if (is_other && !S->dead[t]) {
            bool d = false;
            for (int k = 0; k < 4; k++) d |= nb[k] >= 0 && S->dead[nb[k]];
            if (d) { S->dead[t] = 1; *flag = 1; }
        }
        __syncthreads();
        const int again = *flag;
        __syncthreads();
        if (!again) break;
    }
    const bool rest = is_other && !S->dead[t];
    S->rlab[t] = rest ? static_cast<uint16_t>(t) : 0xffffu;
    __syncthreads();
    for (;;) {                            // regions = connected components of what is left
        if (t == 0) *flag = 0;
        __syncthreads();
        if (rest) {
            int best = S->rlab[t];
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0 && S->rlab[nb[k]] != 0xffffu) best = min(best, static_cast<int>(S->rlab[nb[k]]));
            best = min(best, static_cast<int>(S->rlab[best]));
            if (best < S->rlab[t]) { S->rlab[t] = static_cast<uint16_t>(best); *flag = 1; }
        }
        __syncthreads();
        const int again = *flag;
        __syncthreads();
        if (!again) break;
    }
    if (rest && col[t] == 0) S->seeded[S->rlab[t]] = 1;       // a region starts from an empty point (benson.rs:297-301)
    if (on) { S->size[t] = 0; for (int j = 0; j < 4; j++) S->touch[t][j] = 0; }
    __syncthreads();
    const bool root = rest && S->rlab[t] == t && S->seeded[t];
    int ncand = 0;
    if (root) {                           // the chains next to the root point are the only ones the region can be vital to
        for (int k = 0; k < 4; k++)
            if (nb[k] >= 0 && col[nb[k]] == c) {
                const uint16_t ch = lab[nb[k]];
                bool dup = false;
                for (int j = 0; j < ncand; j++) dup |= S->vch[t][j] == ch;
                if (!dup) S->vch[t][ncand++] = ch;
            }
        S->nv[t] = static_cast<uint8_t>(ncand);
    }
    __syncthreads();
    // every point of a region votes: which of the region's candidate chains does it touch?
    const int my_region = rest ? S->rlab[t] : 0xffff;
    const bool in_region = rest && S->seeded[my_region];
    if (in_region) {
        atomicAdd(&S->size[my_region], 1);
        for (int j = 0; j < S->nv[my_region]; j++) {
            const int ch = S->vch[my_region][j];
            bool touches = false;
            for (int k = 0; k < 4; k++) touches |= nb[k] >= 0 && col[nb[k]] == c && lab[nb[k]] == ch;
            if (touches) atomicAdd(&S->touch[my_region][j], 1);
        }
    }
    __syncthreads();
    if (root) {                           // vital = every point of the region touches the chain (benson.rs:188-208)
        int nv = 0;
        uint16_t keep[4];
        for (int j = 0; j < ncand; j++)
            if (S->touch[t][j] == S->size[t]) keep[nv++] = S->vch[t][j];
        for (int j = 0; j < nv; j++) S->vch[t][j] = keep[j];
        S->nv[t] = static_cast<uint8_t>(nv);
        S->reg_on[t] = nv > 0;            // regions that are vital to nobody go first (benson.rs:128-143)
    }
    __syncthreads();
    for (;;) {
        if (t == 0) *flag = 0;
        if (on) S->chain_cnt[t] = 0;
        __syncthreads();
        if (root && S->reg_on[t])
            for (int j = 0; j < S->nv[t]; j++)
                if (S->chain_alive[S->vch[t][j]]) atomicAdd(&S->chain_cnt[S->vch[t][j]], 1);
        __syncthreads();
        // a chain stays alive with two vital regions (benson.rs:95-111); chain_alive is indexed by the chain's label
        if (on && col[t] == c && lab[t] == t && S->chain_alive[t] && S->chain_cnt[t] < 2) { S->chain_alive[t] = 0; *flag = 1; }
        __syncthreads();
        // a region stays while every stone around it is alive (benson.rs:115-131): any point that sees a dead neighbour vetoes
        if (in_region && S->reg_on[my_region]) {
            bool bad = false;
            for (int k = 0; k < 4; k++) bad |= nb[k] >= 0 && col[nb[k]] == c && !S->chain_alive[lab[nb[k]]];
            if (bad) S->veto[my_region] = 1;
        }
        __syncthreads();
        if (root && S->reg_on[t] && S->veto[t]) { S->reg_on[t] = 0; *flag = 1; }
        __syncthreads();
        const int again = *flag;
        __syncthreads();
        if (!again) break;
    }
    if (rest && S->seeded[S->rlab[t]] && S->reg_on[S->rlab[t]]) eye[t] = 1;
    __syncthreads();
}

__device__ void plan_candidates(PlanScratch* S, const uint8_t* col, const uint16_t* lab, const int (&nb)[4], int t, bool on, bool legal,
                                int to_move, int search, uint8_t* out_cand, uint16_t* out_rep, int* flag) {
    // chain labels of stones only matter per colour; chain_alive[] is indexed by label, which is a stone of that chain
    if (t < 8) S->symm[t] = 1;
    __syncthreads();
    if (on)
        for (int tr = 1; tr < 8; tr++)
            if (col[t] != col[c_symmetry[tr][t]]) S->symm[tr] = 0;        // symmetry::is_symmetric (utils/symmetry.rs:139-146)
    bool cand = legal;
    if (search == 1) {                    // ScoringSearch (libdg_mcts/options.rs:109-138)
        device_benson(S, col, lab, nb, t, on, 1, S->eye[0], flag);
        device_benson(S, col, lab, nb, t, on, 2, S->eye[1], flag);
        if (cand && (S->eye[0][t] || S->eye[1][t])) cand = false;
        if (cand) {                       // the own-eye heuristic (options.rs:180-214)
            bool all_cross = true;
            int n_cross = 0;
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0) { n_cross++; all_cross &= col[nb[k]] == to_move; }
            if (all_cross) {
                const int x = t % 19, y = t / 19;
                int diag = 0;
                for (int dy = -1; dy <= 1; dy += 2)
                    for (int dx = -1; dx <= 1; dx += 2) {
                        const int xx = x + dx, yy = y + dy;
                        if (xx >= 0 && xx <= 18 && yy >= 0 && yy <= 18 && col[19 * yy + xx] == to_move) diag++;
                    }
                if (diag >= (n_cross == 2 ? 1 : n_cross == 3 ? 2 : 3)) cand = false;
            }
        }
    }
    __syncthreads();
    if (on) {
        int rep = t;                      // policy_helper.rs:54-72: the smallest index of the point's orbit
        for (int tr = 1; tr < 8; tr++)
            if (S->symm[tr]) rep = min(rep, static_cast<int>(c_symmetry[tr][t]));
        out_rep[t] = static_cast<uint16_t>(rep);
        out_cand[t] = cand && rep == t;
    }
    if (t == 361) {                       // bit 15 of the pass entry: the board has a non-trivial symmetry (orbits to fold)
        int any = 0;
        for (int tr = 1; tr < 8; tr++) any |= S->symm[tr];
        out_rep[361] = static_cast<uint16_t>(361 | (any ? 0x8000 : 0));
        out_cand[361] = search == 0;
    }
}


// ---- ladder reading on the device (utils/ladder.rs:53-179) ------------------------------------------------------------------
// The two ladder planes (features.rs:223-231) are a search: play the atari, extend the chain, count its liberties, recurse
// while it has exactly two.  One WARP per reading.  A board is two registers per lane -- lane y holds the 19-bit row masks
// of the black and of the white stones of board line y (lanes 19..31 hold zeros) -- so a board copy is two moves, a chain is
// a flood fill over rows (occluded Kogge-Stone fill along the row, one shuffle up and one down per sweep), liberties are a
// dilation, captures are "flood the neighbouring chain, no liberty -> clear it".  Nothing is kept incrementally: every
// answer is derived from the two masks, which is what makes the per-step board copies of the reference free here.
// Branches (both liberties of the extended chain are tried, ladder.rs:112-118) go on a per-lane stack in local memory; the
// answer is "any branch captures", so the order in which they are read does not matter.  Neighbour order E, S, W, N
// (iter/adjacent_iter.rs:42-43) decides which chain is extended first, as in the reference.
namespace lad {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t ROW = 0x7ffffu;
constexpr int CAP = 128;      // pending branches: one per step of the line being read; a 19 x 19 board cannot hold a ladder this long

struct B2 { uint32_t s[2]; };                                    // [0] black rows, [1] white rows

__device__ __forceinline__ uint32_t up(uint32_t v, int lane) { const uint32_t r = __shfl_up_sync(FULL, v, 1); return lane ? r : 0u; }
__device__ __forceinline__ uint32_t down(uint32_t v) { return __shfl_down_sync(FULL, v, 1); }     // lane 31 reads its own 0
__device__ __forceinline__ uint32_t dilate(uint32_t v, int lane, uint32_t rm) { return ((v << 1) | (v >> 1) | up(v, lane) | down(v)) & rm; }
__device__ __forceinline__ uint32_t pt(int p, int lane) { return lane == p / 19 ? 1u << (p % 19) : 0u; }
__device__ __forceinline__ int count(uint32_t v) { return static_cast<int>(__reduce_add_sync(FULL, static_cast<unsigned>(__popc(v)))); }
__device__ __forceinline__ bool any(uint32_t v) { return __any_sync(FULL, v != 0); }
__device__ __forceinline__ int first_point(uint32_t v) {          // lowest point of a non-empty set (warp-uniform)
    const int y = __ffs(__ballot_sync(FULL, v != 0)) - 1;
    return y * 19 + __ffs(__shfl_sync(FULL, v, y)) - 1;
}

// connected component of `mask` that contains `seed`
__device__ uint32_t flood(uint32_t seed, uint32_t mask, int lane) {
    uint32_t cur = seed & mask;
    for (;;) {
        uint32_t g = cur, p = mask;
        g |= p & (g << 1); p &= p << 1;
        g |= p & (g << 2); p &= p << 2;
        g |= p & (g << 4); p &= p << 4;
        g |= p & (g << 8); p &= p << 8;
        g |= p & (g << 16);
        uint32_t h = cur;
        p = mask;
        h |= p & (h >> 1); p &= p >> 1;
        h |= p & (h >> 2); p &= p >> 2;
        h |= p & (h >> 4); p &= p >> 4;
        h |= p & (h >> 8); p &= p >> 8;
        h |= p & (h >> 16);
        const uint32_t x = g | h;
        const uint32_t nxt = x | ((up(x, lane) | down(x)) & mask);
        const bool changed = __any_sync(FULL, nxt != cur);
        cur = nxt;
        if (!changed) return cur;
    }
}

// BoardFast::place (board_fast.rs:441-474) without the hash: plays colour index ci at p, removes the enemy chains that lose
// their last liberty, returns the liberties of the chain the stone belongs to (0 = the move was suicide).
__device__ int place(B2& b, int ci, int p, int lane, uint32_t rm) {
    const uint32_t stone = pt(p, lane);
    const uint32_t own = b.s[ci] | stone;
    uint32_t opp = b.s[ci ^ 1];
    uint32_t nb = dilate(stone, lane, rm) & opp;
    while (any(nb)) {
        const uint32_t f = flood(pt(first_point(nb), lane), opp, lane);
        if (!any(dilate(f, lane, rm) & ~(own | opp))) opp &= ~f;
        nb &= ~f;
    }
    b.s[ci] = own;
    b.s[ci ^ 1] = opp;
    return count(dilate(flood(stone, own, lane), lane, rm) & ~(own | opp));
}

// `_can_escape_with_capture` (ladder.rs:33-41): some enemy chain next to the chain `f` (colour index ci) is in atari
__device__ bool chain_can_capture(const B2& b, int ci, uint32_t f, int lane, uint32_t rm) {
    const uint32_t enemy = b.s[ci ^ 1], occupied = b.s[0] | b.s[1];
    uint32_t a = dilate(f, lane, rm) & enemy;
    while (any(a)) {
        const uint32_t g = flood(pt(first_point(a), lane), enemy, lane);
        if (count(dilate(g, lane, rm) & ~occupied) < 2) return true;
        a &= ~g;
    }
    return false;
}

// `_is_ladder_capture` (ladder.rs:53-119) once the attacker (colour index ci) has played p on `start`
__device__ bool capture_search(const B2& start, int ci, int p0, int lane, uint32_t rm, uint32_t* st0, uint32_t* st1, int16_t* stp) {
    const int oi = ci ^ 1;
    int top = 1;
    st0[0] = start.s[0];
    st1[0] = start.s[1];
    stp[0] = static_cast<int16_t>(p0);
    while (top > 0) {
        --top;
        B2 b;
        b.s[0] = st0[top];
        b.s[1] = st1[top];
        const int p = stp[top];
        const int px = p % 19, py = p / 19;
        // the first neighbouring enemy chain that is in atari, cannot capture its way out and may extend into its liberty
        int run = -1, nl = 0;
        B2 t = b;
#pragma unroll 1
        for (int k = 0; k < 4 && run < 0; k++) {
            const int qx = px + (k == 0) - (k == 2), qy = py - (k == 1) + (k == 3);
            if (qx < 0 || qx > 18 || qy < 0 || qy > 18) continue;
            if (!((__shfl_sync(FULL, b.s[oi], qy) >> qx) & 1u)) continue;
            const uint32_t f = flood(pt(qy * 19 + qx, lane), b.s[oi], lane);
            const uint32_t libs = dilate(f, lane, rm) & ~(b.s[0] | b.s[1]);
            if (count(libs) != 1) continue;                                   // not in atari
            if (chain_can_capture(b, oi, f, lane, rm)) continue;
            const int lib = first_point(libs);
            t = b;
            nl = place(t, oi, lib, lane, rm);                                 // 0: the extension is not a legal move
            if (nl > 0) run = lib;
        }
        if (run < 0) continue;                                                // this line of play captures nothing
        if (nl < 2) return true;                                              // still in atari after extending: captured
        if (nl >= 3) continue;                                                // escaped
        b = t;
        const int rx = run % 19, ry = run / 19;
        // the extension put one of the attacker's chains into atari: no ladder (ladder.rs:97-103)
        bool atari = false;
        uint32_t seen = 0;
#pragma unroll 1
        for (int k = 0; k < 4 && !atari; k++) {
            const int qx = rx + (k == 0) - (k == 2), qy = ry - (k == 1) + (k == 3);
            if (qx < 0 || qx > 18 || qy < 0 || qy > 18) continue;
            const uint32_t q = pt(qy * 19 + qx, lane);
            if (!any(q & b.s[ci]) || any(q & seen)) continue;
            const uint32_t f = flood(q, b.s[ci], lane);
            seen |= f;
            atari = count(dilate(f, lane, rm) & ~(b.s[0] | b.s[1])) < 2;
        }
        if (atari) continue;
        // every legal attacker move next to the extension is a branch (ladder.rs:109-118)
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const int qx = rx + (k == 0) - (k == 2), qy = ry - (k == 1) + (k == 3);
            if (qx < 0 || qx > 18 || qy < 0 || qy > 18) continue;
            const int q = qy * 19 + qx;
            if (any(pt(q, lane) & (b.s[0] | b.s[1]))) continue;
            B2 c = b;
            if (place(c, ci, q, lane, rm) == 0) continue;                     // suicide
            if (top < CAP) {
                st0[top] = c.s[0];
                st1[top] = c.s[1];
                stp[top] = static_cast<int16_t>(q);
                ++top;
            }
        }
    }
    return false;
}

// Ladder::is_ladder_capture (ladder.rs:131-135); p is a legal move of colour index ci
__device__ bool is_capture(const B2& b, int ci, int p, int lane, uint32_t rm, uint32_t* st0, uint32_t* st1, int16_t* stp) {
    B2 c = b;
    place(c, ci, p, lane, rm);
    return capture_search(c, ci, p, lane, rm, st0, st1, stp);
}

// Ladder::is_ladder_escape (ladder.rs:144-178); p is a legal move of colour index ci next to an own chain in atari
__device__ bool is_escape(const B2& b, int ci, int p, int lane, uint32_t rm, uint32_t* st0, uint32_t* st1, int16_t* stp) {
    B2 c = b;
    if (place(c, ci, p, lane, rm) != 2) return false;
    const int px = p % 19, py = p / 19;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int qx = px + (k == 0) - (k == 2), qy = py - (k == 1) + (k == 3);
        if (qx < 0 || qx > 18 || qy < 0 || qy > 18) continue;
        const int q = qy * 19 + qx;
        if (any(pt(q, lane) & (c.s[0] | c.s[1]))) continue;
        B2 d = c;
        if (place(d, ci ^ 1, q, lane, rm) == 0) continue;                     // not a legal move of the attacker
        if (capture_search(d, ci ^ 1, q, lane, rm, st0, st1, stp)) return false;
    }
    return true;
}

// row y of a 361-bit point mask (bit p = point p = 19 y + x)
__device__ __forceinline__ uint32_t row_of(const uint32_t* m, int y) {
    const int at = 19 * y, w = at >> 5, sh = at & 31;
    const uint32_t lo = m[w], hi = w + 1 < 12 ? m[w + 1] : 0u;
    return __funnelshift_r(lo, hi, sh) & ROW;
}

}  // namespace lad

// Block n < batch handles position n and also expands its planes into the tower's input rows (what pack_compact_kernel
// does for host-made planes: 400 rows x 64 fp16 channels, halo rows and channels 32..63 zero); block `batch` zeroes the
// rows between the last position and the end of the last 128-row tile (stale after a larger batch).
__global__ void __launch_bounds__(384) planes_from_stones_kernel(const RawPosition* __restrict__ in, uint32_t* __restrict__ out_planes,
                                                                 uint8_t* __restrict__ out_legal, uint4* __restrict__ out_rows,
                                                                 uint8_t* __restrict__ out_cand, uint16_t* __restrict__ out_rep,
                                                                 int batch, int total_rows) {
    __shared__ uint32_t planes_s[361];
    if (static_cast<int>(blockIdx.x) == batch) {
        const long first = static_cast<long>(batch) * DG_POS_ROWS * 8, last = static_cast<long>(total_rows) * 8;
        for (long i = first + threadIdx.x; i < last; i += 384) out_rows[static_cast<long>(DG_GUARD_ROWS) * 8 + i] = make_uint4(0, 0, 0, 0);
        return;
    }
    __shared__ uint32_t lib[361][12];             // liberty set of the chain whose smallest point index is the row
    __shared__ unsigned long long chash[361];     // XOR of the zobrist keys of its stones
    __shared__ uint16_t lab[384];
    __shared__ uint16_t nlib[361];
    __shared__ uint8_t col[384];
    __shared__ int changed, any_ko;

    const RawPosition& r = in[blockIdx.x];
    const int t = threadIdx.x;
    const bool on = t < 361;
    const int x = t % 19, y = t / 19;
    int nb[4];
    nb[0] = (on && x < 18) ? t + 1 : -1;
    nb[1] = (on && y > 0) ? t - 19 : -1;
    nb[2] = (on && x > 0) ? t - 1 : -1;
    nb[3] = (on && y < 18) ? t + 19 : -1;
    const int c = !on ? 0 : bit(r.black, t) ? 1 : bit(r.white, t) ? 2 : 0;
    col[t] = static_cast<uint8_t>(c);
    lab[t] = static_cast<uint16_t>(t);
    for (int i = t; i < 361 * 12; i += 384) (&lib[0][0])[i] = 0;
    if (on) chash[t] = 0;
    if (t == 0) any_ko = 0;
    __syncthreads();

    // 1. chains: every stone takes the smallest label among itself, its same-coloured neighbours and the label's label
    for (;;) {
        if (t == 0) changed = 0;
        __syncthreads();
        if (c) {
            int best = lab[t];
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0 && col[nb[k]] == c) best = min(best, static_cast<int>(lab[nb[k]]));
            best = min(best, static_cast<int>(lab[best]));
            if (best < lab[t]) { lab[t] = static_cast<uint16_t>(best); changed = 1; }
        }
        __syncthreads();
        const int again = changed;
        __syncthreads();
        if (!again) break;
    }

    // 2. liberty sets and stone hashes per chain
    if (on) {
        if (c) atomicXor(&chash[lab[t]], c_zobrist[c - 1][t]);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0 && col[nb[k]]) atomicOr(&lib[lab[nb[k]]][t >> 5], 1u << (t & 31));
        }
    }
    __syncthreads();
    if (on && c && lab[t] == t) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < 12; w++) n += __popc(lib[t][w]);
        nlib[t] = static_cast<uint16_t>(n);
    }
    __syncthreads();

    // 3. plane bits of this point
    const int tm = r.to_move, opp = 3 - tm;
    const bool device_ladders = (r.symmetry & 8) != 0;     // the host did not read the ladders (DG_RAW_DEVICE_LADDERS)
    int ladder_kinds = 0;                                  // 1: a ladder capture can start here, 2: a ladder escape
    uint32_t m = 0;
    bool legal = false;
    if (on && c) {
        const int n = min(static_cast<int>(nlib[lab[t]]), 6);
        m = ((1u << n) - 1u) << (c == tm ? 5 : 17);
    } else if (on) {
        bool ko = false;
        int counts[2];
#pragma unroll
        for (int side_i = 0; side_i < 2; side_i++) {
            const int side = side_i == 0 ? tm : opp;
            uint32_t L[12];
#pragma unroll
            for (int w = 0; w < 12; w++) L[w] = 0;
            int n_empty = 0, nf = 0, nc = 0, fr[4], cap[4];
            bool ok = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int q = nb[k];
                if (q < 0) continue;
                const int cq = col[q];
                if (!cq) {
                    n_empty++;
#pragma unroll
                    for (int w = 0; w < 12; w++) if (w == (q >> 5)) L[w] |= 1u << (q & 31);
                    continue;
                }
                const int root = lab[q];
                const int n = nlib[root];
                if (cq == side) {
                    ok |= n >= 2;
                    bool dup = false;
                    for (int j = 0; j < nf; j++) dup |= fr[j] == root;
                    if (!dup) fr[nf++] = root;
                } else if (n == 1) {
                    bool dup = false;
                    for (int j = 0; j < nc; j++) dup |= cap[j] == root;
                    if (!dup) cap[nc++] = root;
                }
            }
            int count = -1;
            if (ok || n_empty || nc) {
                if (!nf && !nc) count = n_empty;
                else {
                    for (int j = 0; j < nf; j++)
#pragma unroll
                        for (int w = 0; w < 12; w++) L[w] |= lib[fr[j]][w];
#pragma unroll
                    for (int w = 0; w < 12; w++) if (w == (t >> 5)) L[w] &= ~(1u << (t & 31));
                    if (nc) {                    // rare: captured stones that touch the new chain become liberties
                        for (int s = 0; s < 361; s++) {
                            if (col[s] != 3 - side) continue;
                            const int root = lab[s];
                            bool is_cap = false;
                            for (int j = 0; j < nc; j++) is_cap |= cap[j] == root;
                            if (!is_cap) continue;
                            const int sx = s % 19, sy = s / 19;
                            const int sn[4] = {sx < 18 ? s + 1 : -1, sy > 0 ? s - 19 : -1, sx > 0 ? s - 1 : -1, sy < 18 ? s + 19 : -1};
                            bool touches = false;
                            for (int k = 0; k < 4; k++) {
                                const int u = sn[k];
                                if (u < 0) continue;
                                if (u == t) touches = true;
                                else if (col[u] == side)
                                    for (int j = 0; j < nf; j++) touches |= fr[j] == lab[u];
                            }
                            if (touches)
#pragma unroll
                                for (int w = 0; w < 12; w++) if (w == (s >> 5)) L[w] |= 1u << (s & 31);
                        }
                    }
                    count = 0;
#pragma unroll
                    for (int w = 0; w < 12; w++) count += __popc(L[w]);
                }
                if (side_i == 0 && bit(r.visited, t)) {  // super-ko: the position after the move is one of the last 16
                    unsigned long long h = r.hash ^ c_zobrist[tm - 1][t];
                    for (int j = 0; j < nc; j++) h ^= chash[cap[j]];
#pragma unroll
                    for (int j = 0; j < 16; j++) ko |= r.hash_history[j] == h;
                }
            }
            counts[side_i] = count;
        }
        if (counts[0] >= 0) m |= ((1u << min(counts[0], 6)) - 1u) << 11;
        if (counts[1] >= 0) m |= ((1u << min(counts[1], 6)) - 1u) << 23;
        if (ko) { m |= 1u << 29; any_ko = 1; }
        if (counts[0] >= 0) {
            if (device_ladders) {
                // a ladder can start here if the move puts a neighbouring enemy chain in atari (it has two liberties now)
                // or extends an own chain that is in atari
                bool captures = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int q = nb[k];
                    if (q < 0 || !col[q]) continue;
                    const int n = nlib[lab[q]];
                    if (col[q] == opp && n == 2) ladder_kinds |= 1;
                    if (col[q] == opp && n == 1) captures = true;
                    if (col[q] == tm && n == 1) ladder_kinds |= 2;
                }
                // decided without reading anything (csrc/go_board.h: ladder_capture_first_step, is_ladder_escape): an attacker
                // stone that neither captures nor has two liberties is captured itself; an extension that does not end up
                // with exactly two liberties is no ladder escape
                if (!captures && counts[0] < 2) ladder_kinds &= ~1;
                if (counts[0] != 2) ladder_kinds &= ~2;
            } else {
                if (bit(r.ladder_capture, t)) m |= 1u << 30;
                if (bit(r.ladder_escape, t)) m |= 1u << 31;
            }
        }
        legal = counts[0] >= 0 && !ko;
    }
    if (on) {
        if (r.last_move[0] == t) m |= 1u << 3;
        if (r.last_move[1] == t) m |= 1u << 4;
    }
    __syncthreads();
    if (device_ladders) {
        // 3b. the ladder planes: a work list of (point, kind) in the liberty-set storage (free now), one warp per reading
        uint32_t* scratch = &lib[0][0];
        uint32_t* result = scratch;                                   // [361] bits 30 / 31
        uint16_t* work = reinterpret_cast<uint16_t*>(scratch + 384);  // [722]
        int* counters = reinterpret_cast<int*>(scratch + 384 + 384);  // [0] items, [1] next
        if (on) result[t] = 0;
        if (t < 2) counters[t] = 0;
        __syncthreads();
        if (ladder_kinds & 1) work[atomicAdd(&counters[0], 1)] = static_cast<uint16_t>(t);
        if (ladder_kinds & 2) work[atomicAdd(&counters[0], 1)] = static_cast<uint16_t>(t | 512);
        __syncthreads();
        const int items = counters[0];
        if (items > 0) {
            const int lane = t & 31;
            const uint32_t rm = lane < 19 ? lad::ROW : 0u;
            lad::B2 board;
            board.s[0] = lane < 19 ? lad::row_of(r.black, lane) : 0u;
            board.s[1] = lane < 19 ? lad::row_of(r.white, lane) : 0u;
            uint32_t st0[lad::CAP], st1[lad::CAP];
            int16_t stp[lad::CAP];
            for (;;) {
                int i = 0;
                if (lane == 0) i = atomicAdd(&counters[1], 1);
                i = __shfl_sync(lad::FULL, i, 0);
                if (i >= items) break;
                const int item = work[i], p = item & 511;
                const bool yes = (item & 512) ? lad::is_escape(board, tm - 1, p, lane, rm, st0, st1, stp)
                                              : lad::is_capture(board, tm - 1, p, lane, rm, st0, st1, stp);
                if (yes && lane == 0) atomicOr(&result[p], (item & 512) ? 1u << 31 : 1u << 30);
            }
        }
        __syncthreads();
        if (on) m |= result[t];
        __syncthreads();
    }
    // 4. output
    if (on) {
        const uint32_t global = (tm == 1 ? 1u : 2u) | (any_ko ? 4u : 0u);
        uint32_t* out = out_planes + static_cast<size_t>(blockIdx.x) * 362;
        const int target = c_symmetry[r.symmetry & 7][t];
        out[target] = m | global;
        planes_s[target] = m | global;
        if (t == 0) out[361] = r.k_bits;
        out_legal[static_cast<size_t>(blockIdx.x) * 361 + t] = legal ? 1 : 0;
    }
    __syncthreads();
    // 5. the tower's input rows of this position: one 16-byte chunk (8 channels) per thread and step, coalesced
    const uint32_t kbits = r.k_bits;
    uint4* rows = out_rows + (static_cast<long>(DG_GUARD_ROWS) + static_cast<long>(blockIdx.x) * DG_POS_ROWS) * 8;
    for (int idx = t; idx < DG_POS_ROWS * 8; idx += 384) {
        const int q = idx >> 3, chunk = idx & 7;
        const int qx = q % DG_LINE_STRIDE, qy = q / DG_LINE_STRIDE;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (chunk < 4 && qx < 19 && qy < 19) {
            const uint32_t mask = planes_s[qy * 19 + qx] >> (chunk * 8);
            uint32_t h[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t one = (chunk == 0 && j < 2) ? kbits : 0x3c00u;
                h[j] = ((mask >> j) & 1u) ? one : 0u;
            }
            v = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
        }
        rows[idx] = v;
    }
    // 6. optional: what create_initial_policy derives from the board (pool/policy_helper.rs:28-75) -- the candidate mask
    //    of the position's search options and the orbit representatives of a symmetric board -- for the prior kernel
    if (out_cand) {
        __syncthreads();                              // everybody is done with lib[][]: it becomes scratch space
        plan_candidates(reinterpret_cast<PlanScratch*>(&lib[0][0]), col, lab, nb, t, on, legal, tm, r.symmetry >> 4,
                        out_cand + static_cast<size_t>(blockIdx.x) * 362, out_rep + static_cast<size_t>(blockIdx.x) * 362, &changed);
    }
}

// add_valid_candidates + normalize_policy (pool/policy_helper.rs:87-134) on the device: the network's policy (fp16, in the
// orientation the position was evaluated in) is un-transformed, folded onto the orbit representatives, masked by the
// candidates and scaled to sum 1 -- in the host code's operation order (search_task.h: PriorPlan::apply), so that the result
// is bit-identical and the search that consumes it builds the same tree.  One block per position.
__constant__ uint8_t c_sym_inverse[8] = {0, 1, 2, 3, 4, 7, 6, 5};

__global__ void __launch_bounds__(384) prior_from_policy_kernel(const RawPosition* __restrict__ raw, const __half* __restrict__ policy,
                                                                const uint8_t* __restrict__ cand, const uint16_t* __restrict__ rep,
                                                                float* __restrict__ prior_out) {
    __shared__ float pol[362];
    __shared__ float pri[368];
    __shared__ uint16_t target[361];
    __shared__ float lane[8];
    __shared__ int n_finite;
    const int n = blockIdx.x, t = threadIdx.x;
    const int sym = raw[n].symmetry & 7, inv = c_sym_inverse[sym];
    const bool folded = (rep[static_cast<size_t>(n) * 362 + 361] & 0x8000u) != 0;
    if (t < 362) pol[t] = __half2float(policy[static_cast<size_t>(n) * 362 + t]);
    if (t < 361 && folded) target[t] = rep[static_cast<size_t>(n) * 362 + c_symmetry[inv][t]];
    if (t == 0) n_finite = 0;
    __syncthreads();
    if (t < 368) {
        float v = (t < 362 && cand[static_cast<size_t>(n) * 362 + t]) ? 0.0f : -INFINITY;
        if (t == 361) v += pol[361];
        if (t < 361) {
            if (!folded) v += pol[c_symmetry[sym][t]];    // asymmetric board: point t receives exactly its own image
            else
                for (int i = 0; i < 361; i++)
                    if (target[i] == t) v += pol[i];      // ascending source index, as the host loop adds them
        }
        pri[t] = v;
        if (isfinite(v)) atomicAdd(&n_finite, 1);
    }
    __syncthreads();
    if (t < 8) {                                           // asm/sum_finite.rs:23-57: eight interleaved lanes ...
        float s = 0.0f;
        for (int i = t; i < 368; i += 8)
            if (isfinite(pri[i])) s += pri[i];
        lane[t] = s;
    }
    __syncthreads();
    if (t < 368) {
        const float sum = ((lane[0] + lane[1]) + (lane[2] + lane[3])) + ((lane[4] + lane[5]) + (lane[6] + lane[7]));   // ... added pairwise
        float v = pri[t];
        if (sum < 1e-6f) {
            if (isfinite(v)) v = __fdiv_rn(1.0f, static_cast<float>(n_finite));
        } else {
            v = __fmul_rn(v, __fdiv_rn(1.0f, sum));
        }
        prior_out[static_cast<size_t>(n) * 368 + t] = v;
    }
}

cudaError_t launch_prior_from_policy(const void* raw, const void* policy, const void* cand, const void* rep, float* prior, int batch,
                                     cudaStream_t s) {
    prior_from_policy_kernel<<<batch, 384, 0, s>>>(static_cast<const RawPosition*>(raw), static_cast<const __half*>(policy),
                                                   static_cast<const uint8_t*>(cand), static_cast<const uint16_t*>(rep), prior);
    return cudaGetLastError();
}

cudaError_t upload_feature_tables(const unsigned long long* zobrist /* [2][361] */, const uint16_t* symmetry /* [8][361] */) {
    cudaError_t rc = cudaMemcpyToSymbol(c_zobrist, zobrist, sizeof(c_zobrist));
    if (rc != cudaSuccess) return rc;
    return cudaMemcpyToSymbol(c_symmetry, symmetry, sizeof(c_symmetry));
}

cudaError_t launch_planes_from_stones(const void* raw, void* planes, void* legal, void* rows64, void* cand, void* rep, int batch,
                                      cudaStream_t s) {
    const int total = dg_num_tiles(batch) * DG_TILE_M;
    planes_from_stones_kernel<<<batch + 1, 384, 0, s>>>(static_cast<const RawPosition*>(raw), static_cast<uint32_t*>(planes),
                                                        static_cast<uint8_t*>(legal), static_cast<uint4*>(rows64),
                                                        static_cast<uint8_t*>(cand), static_cast<uint16_t*>(rep), batch, total);
    return cudaGetLastError();
}

}  // namespace dg
