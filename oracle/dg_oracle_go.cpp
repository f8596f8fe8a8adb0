/*
 * dg_oracle_go.cpp -- CPU restatement of dream-go's `libdg_go` rules, ladder reader, symmetry tables and
 * V1 feature planes.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may load this; the product (dream_go_b200/csrc) never links or calls it.
 *
 * Every function follows one function of the reference (paths relative to /root/reference/src/libdg_go)
 * statement by statement, including iteration orders (they decide which liberty `get_a_liberty` returns
 * and therefore how a ladder is read) and the data representation (one u32 word per vertex).
 *
 * Parity pins (tests/test_oracle_go.py): board.rs:281-388, board_fast.rs:548-562, utils/ladder.rs:187-351,
 * utils/symmetry.rs:148-201, dg_tests/tests/real_games.rs:49,74,117 (zobrist hashes, with the reference's
 * table supplied as a data fixture), and the legal replay of dg_tests/fixtures/example_games.sgf.
 * Feature VALUES are not pinned by any reference test (features.rs:477-493 only asserts the length);
 * they follow features.rs:154-250 and are additionally pinned by hand-built positions in the tests.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace {

// ---- point.rs:23-32 -------------------------------------------------------------------------------
constexpr int STRIDE = 20;
constexpr int MAXP = STRIDE * 20 + 20;   // Point::MAX = 420
constexpr int BLACK = 1, WHITE = 2;      // color.rs:18-21

inline int point_new(int x, int y) { return STRIDE * (y + 1) + (x + 1); }       // point.rs:26-34
inline int point_x(int p) { int c = p % STRIDE; return c == 0 ? 0 : c - 1; }     // point.rs:57-88 (table TO_X)
inline int point_y(int p) { int r = p / STRIDE; return r == 0 ? 0 : r - 1; }     // point.rs:90-122 (table TO_Y)
inline int point_offset(int p, int dx, int dy) {                                 // point.rs:124-135 (saturating u16)
    int delta = STRIDE * dy + dx;
    int q = p + delta;
    if (q < 0) q = 0;
    if (q > 65535) q = 65535;
    return q;
}
inline int to_packed_index(int p) { return p == 0 ? 361 : 19 * point_y(p) + point_x(p); }   // point.rs:137-143
inline int opposite(int c) { return c == BLACK ? WHITE : BLACK; }                // color.rs:36-41

// iter/adjacent_iter.rs:42-43 -- Right, Down (-y), Left, Up (+y)
const int ADJ_DX[4] = {1, 0, -1, 0};
const int ADJ_DY[4] = {0, -1, 0, 1};

// ---- zobrist.rs:18 -- [3][420] random u64; the VALUES are data the caller may override -------------
uint64_t g_zobrist[3][MAXP];
bool g_zobrist_ready = false;

uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
void zobrist_default() {
    uint64_t s = 0x6472656d2d676f21ull;
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < MAXP; ++i) g_zobrist[c][i] = splitmix64(s);
    g_zobrist_ready = true;
}

// ---- point_state.rs:39-107 -- one u32 per vertex ------------------------------------------------------
inline bool v_is_valid(uint32_t v) { return (v & 3u) != 3u; }
inline int v_color(uint32_t v) { int c = v & 3u; return c == 3 ? 0 : c; }     // COLORS[3] = None
inline int v_next(uint32_t v) { return (v & 0x00000ffcu) >> 2; }
inline int v_head(uint32_t v) { return (v & 0x003ff000u) >> 12; }
inline int v_libs(uint32_t v) { return (v & 0x7fc00000u) >> 22; }
inline bool v_visited(uint32_t v) { return (v & 0x80000000u) != 0; }
inline void v_set_color(uint32_t& v, int c) { v = (v & 0xfffffffcu) | (uint32_t)c; }
inline void v_set_next(uint32_t& v, int p) { v = (v & 0xfffff003u) | ((uint32_t)p << 2); }
inline void v_set_head(uint32_t& v, int p) { v = (v & 0xffc00fffu) | ((uint32_t)p << 12); }
inline void v_set_libs(uint32_t& v, int n) { v = (v & 0x803fffffu) | ((uint32_t)n << 22); }
inline void v_set_visited(uint32_t& v, bool b) { v = (v & 0x7fffffffu) | (b ? 0x80000000u : 0u); }

inline bool contains4(const int a[4], int x) { return a[0] == x || a[1] == x || a[2] == x || a[3] == x; }

// ---- board_fast.rs --------------------------------------------------------------------------------------
struct BoardFast {
    uint32_t vertices[MAXP];

    BoardFast() {                                           // board_fast.rs:95-105
        for (int i = 0; i < MAXP; ++i) vertices[i] = 3u;    // invalid()
        for (int y = 0; y < 19; ++y)
            for (int x = 0; x < 19; ++x) vertices[point_new(x, y)] = 0u;
    }
    bool is_part_of(int p) const { return p >= 0 && p < MAXP && v_is_valid(vertices[p]); }   // :73-79

    // adjacent_to (:115-120): AdjacentIter filtered by is_part_of; returns the count, points in out[]
    int adjacent_to(int p, int out[4]) const {
        int n = 0;
        for (int d = 0; d < 4; ++d) {
            int q = point_offset(p, ADJ_DX[d], ADJ_DY[d]);
            if (is_part_of(q)) out[n++] = q;
        }
        return n;
    }
    int get_n_liberty(int p) const { return v_libs(vertices[v_head(vertices[p])]); }          // :170-174
    bool has_n_liberty(int p, int n) const { return get_n_liberty(p) >= n; }                  // :204-208

    // block_at (:130-132) + iter/chain_iter.rs:33-52: start, then follow `next` until back at the start
    template <class F> void for_block(int start, F&& f) const {
        int cur = start;
        do {
            if (!f(cur)) return;
            cur = v_next(vertices[cur]);
        } while (cur != start);
    }

    int get_a_liberty(int p) const {                        // :182-192; 0 = None
        int found = 0;
        for_block(p, [&](int cur) {
            int adj[4];
            int n = adjacent_to(cur, adj);
            for (int i = 0; i < n; ++i)
                if (v_color(vertices[adj[i]]) == 0) { found = adj[i]; return false; }
            return true;
        });
        return found;
    }

    bool is_valid(int color, int p) const {                 // :216-243
        if (v_color(vertices[p]) != 0) return false;
        int adj[4];
        int n = adjacent_to(p, adj);
        for (int i = 0; i < n; ++i) {
            int value = v_color(vertices[adj[i]]);
            if (value == 0) return true;
            if ((value == color) == has_n_liberty(adj[i], 2)) return true;
        }
        return false;
    }

    bool is_liberty_of(int liberty, int block_at) const {   // :253-264
        int block_color = v_color(vertices[block_at]);
        int adj[4];
        int n = adjacent_to(liberty, adj);
        for (int i = 0; i < n; ++i) {
            bool same_color = v_color(vertices[adj[i]]) == block_color;
            bool same_block = v_head(vertices[adj[i]]) == block_at;
            if (same_color && same_block) return true;
        }
        return false;
    }

    void join_blocks(int one, int two) {                    // :277-327
        int head_one = v_head(vertices[one]);
        int head_two = v_head(vertices[two]);
        if (head_one == head_two) return;
        vertices[head_two] -= (1u << 22);                   // sub_liberties(1)

        bool already_added[MAXP];
        memset(already_added, 0, sizeof(already_added));
        int extra = 0;
        for_block(one, [&](int point) {
            int adj[4];
            int n = adjacent_to(point, adj);
            for (int i = 0; i < n; ++i) {
                bool is_empty = v_color(vertices[adj[i]]) == 0;
                bool is_new = !already_added[adj[i]];
                if (is_empty && is_new && !is_liberty_of(adj[i], head_two)) {
                    already_added[adj[i]] = true;
                    extra += 1;
                }
            }
            return true;
        });
        for_block(one, [&](int point) { v_set_head(vertices[point], head_two); return true; });
        vertices[head_two] += ((uint32_t)extra << 22);

        int one_prev = v_next(vertices[one]);
        int two_prev = v_next(vertices[two]);
        v_set_next(vertices[two], one_prev);
        v_set_next(vertices[one], two_prev);
    }

    void incr_adjacent_liberties(int start) {               // :337-355 (raw AdjacentIter, index i of the raw iterator)
        int already_changed[4] = {0, 0, 0, 0};
        int head = v_head(vertices[start]);
        for (int i = 0; i < 4; ++i) {
            int adj = point_offset(start, ADJ_DX[i], ADJ_DY[i]);
            uint32_t st = vertices[adj];
            if (v_color(st) != 0) {
                int adj_head = v_head(st);
                if (adj_head != head && !contains4(already_changed, adj_head)) {
                    already_changed[i] = adj_head;
                    vertices[adj_head] += (1u << 22);
                }
            }
        }
    }

    uint64_t capture_if(int color, int p) const {           // :365-374
        uint64_t adjust = 0;
        for_block(p, [&](int cur) { adjust ^= g_zobrist[color][cur]; return true; });
        return adjust;
    }

    uint64_t capture(int color, int p) {                    // :385-396
        // the chain links of removed stones stay intact, so the walk can read `next` after the update
        uint64_t hash = 0;
        int cur = p;
        do {
            hash ^= g_zobrist[color][cur];
            v_set_color(vertices[cur], 0);
            incr_adjacent_liberties(cur);
            cur = v_next(vertices[cur]);
        } while (cur != p);
        return hash;
    }

    uint64_t place_if(int color, int p) const {             // :406-424 (index i of the FILTERED iterator)
        int opponent = opposite(color);
        int seen[4] = {0, 0, 0, 0};
        uint64_t adjust = g_zobrist[color][p];
        int adj[4];
        int n = adjacent_to(p, adj);
        for (int i = 0; i < n; ++i) {
            int head = v_head(vertices[adj[i]]);
            if (v_color(vertices[head]) == opponent && !has_n_liberty(head, 2)) {
                if (!contains4(seen, head)) {
                    seen[i] = head;
                    adjust ^= capture_if(opponent, head);
                }
            }
        }
        return adjust;
    }

    uint64_t place(int color, int p) {                      // :434-475
        int adj[4];
        int n = adjacent_to(p, adj);
        int immediate = 0;
        for (int i = 0; i < n; ++i) immediate += v_color(vertices[adj[i]]) == 0;

        v_set_color(vertices[p], color);
        v_set_next(vertices[p], p);
        v_set_head(vertices[p], p);
        v_set_libs(vertices[p], immediate);
        v_set_visited(vertices[p], true);

        uint64_t hash = g_zobrist[color][p];
        int seen[4] = {0, 0, 0, 0};
        int opponent = opposite(color);
        for (int i = 0; i < 4; ++i) {                       // raw AdjacentIter
            int other = point_offset(p, ADJ_DX[i], ADJ_DY[i]);
            int value = v_color(vertices[other]);
            if (value == color) {
                join_blocks(p, other);
            } else if (value == opponent) {
                int head = v_head(vertices[other]);
                if (!contains4(seen, head)) {
                    vertices[head] -= (1u << 22);
                    seen[i] = head;
                    if (!has_n_liberty(head, 1)) hash ^= capture(opponent, head);
                }
            }
        }
        return hash;
    }

    int get_n_liberty_if(int color, int p) const {          // :484-539
        bool already_seen[MAXP];
        memset(already_seen, 0, sizeof(already_seen));
        int num = 0;
        already_seen[p] = true;
        int captured[4] = {0, 0, 0, 0};
        int connected[4] = {0, 0, 0, 0};
        int opp = opposite(color);
        int adj[4];
        int n = adjacent_to(p, adj);
        for (int i = 0; i < n; ++i) {
            int other = adj[i];
            int value = v_color(vertices[other]);
            if (value == color) {
                int head = v_head(vertices[other]);
                if (!contains4(connected, head)) connected[i] = head;
            } else if (value == opp) {
                int head = v_head(vertices[other]);
                if (get_n_liberty(head) == 1) {
                    captured[i] = head;
                    already_seen[other] = true;
                    num += 1;
                }
            } else {
                already_seen[other] = true;
                num += 1;
            }
        }
        for (int k = 0; k < 4; ++k) {
            int head = connected[k];
            if (head == 0) continue;
            // adjacencies_of (:153-158): every valid neighbour of every stone of the chain
            for_block(head, [&](int cur) {
                int a2[4];
                int m = adjacent_to(cur, a2);
                for (int j = 0; j < m; ++j) {
                    int ap = a2[j];
                    if (!already_seen[ap]) {
                        already_seen[ap] = true;
                        if (v_color(vertices[ap]) == 0) num += 1;
                        else if (contains4(captured, v_head(vertices[ap]))) num += 1;
                    }
                }
                return true;
            });
        }
        return num;
    }
};

// ---- utils/ladder.rs ----------------------------------------------------------------------------------------
bool can_escape_with_capture(const BoardFast& b, int color, int p) {    // ladder.rs:33-41
    int opponent = opposite(color);
    bool any = false;
    b.for_block(p, [&](int cur) {
        int adj[4];
        int n = b.adjacent_to(cur, adj);
        for (int i = 0; i < n; ++i)
            if (v_color(b.vertices[adj[i]]) == opponent && !b.has_n_liberty(adj[i], 2)) { any = true; return false; }
        return true;
    });
    return any;
}

long g_ladder_nodes = 0;   // instrumentation for the CPU baseline (share of time spent reading ladders)

bool is_ladder_capture_rec(BoardFast board, int color, int p) {          // ladder.rs:53-119 (board by value = clone)
    ++g_ladder_nodes;
    board.place(color, p);
    int opponent = opposite(color);
    int adj[4];
    int n = board.adjacent_to(p, adj);
    int opponent_index = 0;
    for (int i = 0; i < n && !opponent_index; ++i) {
        int other = adj[i];
        if (v_color(board.vertices[other]) != opponent) continue;
        bool in_atari = !board.has_n_liberty(other, 2);
        if (in_atari && !can_escape_with_capture(board, opponent, other)) {
            int lib = board.get_a_liberty(other);
            if (lib != 0 && board.is_valid(opponent, lib)) opponent_index = lib;
        }
    }
    if (!opponent_index) return false;
    board.place(opponent, opponent_index);
    if (!board.has_n_liberty(opponent_index, 2)) return true;
    if (board.has_n_liberty(opponent_index, 3)) return false;

    n = board.adjacent_to(opponent_index, adj);
    for (int i = 0; i < n; ++i)
        if (v_color(board.vertices[adj[i]]) == color && !board.has_n_liberty(adj[i], 2)) return false;
    for (int i = 0; i < n; ++i)
        if (board.is_valid(color, adj[i]) && is_ladder_capture_rec(board, color, adj[i])) return true;
    return false;
}

bool is_ladder_capture(const BoardFast& b, int color, int p) { return is_ladder_capture_rec(b, color, p); }  // :131-135

bool is_ladder_escape(const BoardFast& self, int color, int p) {         // ladder.rs:144-178
    int adj[4];
    int n = self.adjacent_to(p, adj);
    bool connected_to_one = false;
    for (int i = 0; i < n; ++i)
        if (v_color(self.vertices[adj[i]]) == color && !self.has_n_liberty(adj[i], 2)) { connected_to_one = true; break; }
    if (!connected_to_one) return false;
    BoardFast board = self;
    board.place(color, p);
    if (board.get_n_liberty(p) != 2) return false;
    for (int i = 0; i < n; ++i) {
        int other = adj[i];
        if (board.is_valid(opposite(color), other) && is_ladder_capture_rec(board, opposite(color), other)) return false;
    }
    return true;
}

// ---- board.rs ------------------------------------------------------------------------------------------------
struct Board {
    BoardFast inner;
    int history[8];          // circular_buf.rs:48-88 (CircularBuf<Point>)
    int history_pos;
    uint64_t zobrist_hash;
    uint64_t zobrist_history[16];   // small_set.rs:17-47 (SmallSet64)
    int zobrist_count;
    float komi;
    int count;
    int last_played;         // 0 = None

    explicit Board(float k) : history_pos(0), zobrist_hash(0), zobrist_count(0), komi(k), count(0), last_played(0) {   // board.rs:51-62
        memset(history, 0, sizeof(history));
        memset(zobrist_history, 0, sizeof(zobrist_history));
    }
    int to_move() const { return last_played ? opposite(last_played) : BLACK; }     // board.rs:102-107
    bool zobrist_contains(uint64_t h) const {
        for (int i = 0; i < 16; ++i) if (zobrist_history[i] == h) return true;
        return false;
    }
    bool is_ko(int color, int p) const {                   // board.rs:132-141
        return v_visited(inner.vertices[p]) && zobrist_contains(zobrist_hash ^ inner.place_if(color, p));
    }
    bool is_valid(int color, int p) const { return inner.is_valid(color, p) && !is_ko(color, p); }   // board.rs:151-153
    void place(int color, int p) {                         // board.rs:164-188
        zobrist_hash ^= inner.place(color, p);
        last_played = color;
        count += 1;
        history[history_pos] = p;
        history_pos = (history_pos + 1) & 7;
        zobrist_history[zobrist_count] = zobrist_hash;
        zobrist_count = (zobrist_count + 1) & 15;
    }
    int history_at(int i) const { return history[(history_pos + 7 - i) & 7]; }     // circular_buf.rs:32-45,94-100
};

// ---- utils/symmetry.rs:21-119 ---------------------------------------------------------------------------------
// order of symmetry::ALL (:121-130): Identity, FlipLR, FlipUD, Transpose, TransposeAnti, Rot90, Rot180, Rot270
int g_sym[8][MAXP];
bool g_sym_ready = false;
void symmetry_init() {
    for (int t = 0; t < 8; ++t) {
        for (int i = 0; i < MAXP; ++i) g_sym[t][i] = 0;
        for (int yy = 0; yy < 19; ++yy)
            for (int xx = 0; xx < 19; ++xx) {
                int x = xx - 9, y = yy - 9, tx = 0, ty = 0;
                switch (t) {
                    case 0: tx = x;  ty = y;  break;   // :40
                    case 1: tx = -x; ty = y;  break;   // :43
                    case 2: tx = x;  ty = -y; break;   // :46
                    case 3: tx = y;  ty = x;  break;   // :49
                    case 4: tx = -y; ty = -x; break;   // :52
                    case 5: tx = y;  ty = -x; break;   // :55
                    case 6: tx = -x; ty = -y; break;   // :58
                    case 7: tx = -y; ty = x;  break;   // :61
                }
                g_sym[t][point_new(xx, yy)] = point_new(tx + 9, ty + 9);
            }
    }
    g_sym_ready = true;
}
const int SYM_INVERSE[8] = {0, 1, 2, 3, 4, 7, 6, 5};      // symmetry.rs:78-89

// fp16 (libdg_utils/types/fp16.rs:55-69, llvm.convert.to.fp16.f32 = round to nearest even)
uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - exp;
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

// ---- utils/features.rs:154-250 (V1, HWC order, T = f16) ------------------------------------------------------
void features_v1(const Board& b, int to_move, int symmetry, uint16_t* out /* [361*32] */) {
    const uint16_t c_1 = 0x3c00;
    memset(out, 0, 361 * 32 * sizeof(uint16_t));
    const int* table = g_sym[symmetry];
    int opponent = opposite(to_move);
    auto idx = [](int c, int point) { return 32 * to_packed_index(point) + c; };   // features.rs:52-60

    for (int i = 0; i < 2; ++i) {                           // :168-175
        int point = b.history_at(i);
        if (point != 0) out[idx(3 + i, table[point])] = c_1;
    }
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {     // :178-205
        int index = point_new(x, y);
        int other = table[index];
        int color = v_color(b.inner.vertices[index]);
        if (color != 0) {
            int start = color == to_move ? 5 : 17;
            int nl = b.inner.get_n_liberty(index);
            if (nl > 6) nl = 6;
            for (int i = 0; i < nl; ++i) out[idx(start + i, other)] = c_1;
        } else {
            if (b.inner.is_valid(to_move, index)) {
                int nl = b.inner.get_n_liberty_if(to_move, index);
                if (nl > 6) nl = 6;
                for (int i = 0; i < nl; ++i) out[idx(11 + i, other)] = c_1;
            }
            if (b.inner.is_valid(opponent, index)) {
                int nl = b.inner.get_n_liberty_if(opponent, index);
                if (nl > 6) nl = 6;
                for (int i = 0; i < nl; ++i) out[idx(23 + i, other)] = c_1;
            }
        }
    }
    uint16_t is_ko = 0;                                     // :208-233
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
        int index = point_new(x, y);
        int other = table[index];
        if (v_color(b.inner.vertices[index]) != 0) continue;
        if (!b.inner.is_valid(to_move, index)) continue;
        if (b.is_ko(to_move, index)) { is_ko = c_1; out[idx(29, other)] = c_1; }
        if (is_ladder_capture(b.inner, to_move, index)) out[idx(30, other)] = c_1;
        if (is_ladder_escape(b.inner, to_move, index)) out[idx(31, other)] = c_1;
    }
    float k = 0.5f + (0.5f * b.komi) / 7.5f;                // :236
    if (k > 1.0f) k = 1.0f;
    if (k < 0.0f) k = 0.0f;
    uint16_t c_komi = f32_to_f16(k);
    uint16_t is_black = to_move == BLACK ? c_komi : 0;
    uint16_t is_white = to_move == WHITE ? c_komi : 0;
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {     // :241-247
        int other = table[point_new(x, y)];
        out[idx(0, other)] = is_black;
        out[idx(1, other)] = is_white;
        out[idx(2, other)] = is_ko;
    }
}

// ---- utils/flood_fill.rs + utils/benson.rs (unconditional life) ------------------------------------------------
// PointStatus (benson.rs:48-53)
enum { ST_NONE = 0, ST_BLOCK = 1, ST_REGION = 2 };

struct Benson {
    uint8_t points[MAXP];

    struct Region { std::vector<int> points, neighbours; };
    struct Block { int at; std::vector<int> stones; std::vector<uint8_t> adjacent; };   // adjacent[p] = p in adjacencies_of(at)

    static bool is_vital(const Region& r, const Block& b) {          // benson.rs:188-190, 197-208
        for (int p : r.points) if (!b.adjacent[p]) return false;
        return true;
    }

    Benson(const BoardFast& board, int to_move) {                    // benson.rs:62-78
        memset(points, ST_NONE, sizeof(points));
        // AllRegionsImpl::all (benson.rs:293-320): flood from every empty point through everything that is not
        // `to_move` coloured (flood_fill.rs:47-88); regions without a `to_move` neighbour are dropped
        std::vector<Region> regions;
        {
            int head[MAXP];
            for (int i = 0; i < MAXP; ++i) head[i] = -1;
            for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
                int start = point_new(x, y);
                if (head[start] != -1 || v_color(board.vertices[start]) != 0) continue;
                Region r;
                std::vector<int> queue{start};
                head[start] = start;
                for (size_t qi = 0; qi < queue.size(); ++qi) {
                    int point = queue[qi];
                    r.points.push_back(point);
                    int adj[4];
                    int n = board.adjacent_to(point, adj);
                    for (int i = 0; i < n; ++i) {
                        if (v_color(board.vertices[adj[i]]) == to_move) r.neighbours.push_back(adj[i]);
                        else if (head[adj[i]] == -1) { head[adj[i]] = start; queue.push_back(adj[i]); }
                    }
                }
                if (!r.neighbours.empty()) regions.push_back(std::move(r));
            }
        }
        // AllBlocksImpl::all (benson.rs:214-232)
        std::vector<Block> blocks;
        {
            bool visited[MAXP];
            memset(visited, 0, sizeof(visited));
            for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
                int p = point_new(x, y);
                if (v_color(board.vertices[p]) != to_move) continue;
                int hd = v_head(board.vertices[p]);
                if (visited[hd]) continue;
                visited[hd] = true;
                Block b;
                b.at = p;
                b.adjacent.assign(MAXP, 0);
                board.for_block(p, [&](int cur) {
                    b.stones.push_back(cur);
                    int adj[4];
                    int n = board.adjacent_to(cur, adj);
                    for (int i = 0; i < n; ++i) b.adjacent[adj[i]] = 1;
                    return true;
                });
                blocks.push_back(std::move(b));
            }
        }
        for (auto& b : blocks) for (int p : b.stones) points[p] = ST_BLOCK;          // mark_all_blocks
        for (auto& r : regions) for (int p : r.points) points[p] = ST_REGION;        // mark_all_regions
        {   // remove_non_vital_regions (benson.rs:128-143)
            std::vector<Region> keep;
            for (auto& r : regions) {
                bool vital = false;
                for (auto& b : blocks) if (is_vital(r, b)) { vital = true; break; }
                if (vital) keep.push_back(std::move(r));
                else for (int p : r.points) points[p] = ST_NONE;
            }
            regions.swap(keep);
        }
        for (;;) {                                                                   // benson.rs:73-75 (non-short-circuit |)
            bool changed = false;
            {   // remove_non_alive_blocks (:95-111)
                std::vector<Block> keep;
                for (auto& b : blocks) {
                    int vital = 0;
                    for (auto& r : regions) vital += is_vital(r, b);
                    if (vital >= 2) keep.push_back(std::move(b));
                    else { for (int p : b.stones) points[p] = ST_NONE; changed = true; }
                }
                blocks.swap(keep);
            }
            {   // remove_non_surrounded_regions (:115-131); `points` is updated while the list is filtered
                std::vector<Region> keep;
                for (auto& r : regions) {
                    bool healthy = true;
                    for (int p : r.neighbours) if (points[p] != ST_BLOCK) { healthy = false; break; }
                    if (healthy) keep.push_back(std::move(r));
                    else { for (int p : r.points) points[p] = ST_NONE; changed = true; }
                }
                regions.swap(keep);
            }
            if (!changed) break;
        }
    }
    bool is_alive(int p) const { return points[p] == ST_BLOCK; }     // :152-154
    bool is_eye(int p) const { return points[p] == ST_REGION; }      // :163-165
};

// utils/score.rs:97-110
bool is_scorable(const Board& b) {
    Benson black(b.inner, BLACK), white(b.inner, WHITE);
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
        int p = point_new(x, y);
        int c = v_color(b.inner.vertices[p]);
        bool ok = c == 0 ? (black.is_eye(p) || white.is_eye(p))
                : c == BLACK ? (black.is_alive(p) || white.is_eye(p))
                             : (white.is_alive(p) || black.is_eye(p));
        if (!ok) return false;
    }
    return true;
}

// utils/score.rs:252-282: distance (through empty points) to the closest stone of `color`, 0xff = unreachable
void territory_distance(const BoardFast& board, int color, uint8_t* territory /* [MAXP] */) {
    memset(territory, 0xff, MAXP);
    std::vector<int> probes;
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
        int p = point_new(x, y);
        if (v_color(board.vertices[p]) == color) { territory[p] = 0; probes.push_back(p); }
    }
    for (size_t i = 0; i < probes.size(); ++i) {
        int index = probes[i];
        int t = territory[index] + 1;
        int adj[4];
        int n = board.adjacent_to(index, adj);
        for (int k = 0; k < n; ++k)
            if (v_color(board.vertices[adj[k]]) == 0 && territory[adj[k]] > t) { probes.push_back(adj[k]); territory[adj[k]] = (uint8_t)t; }
    }
}

// utils/score.rs:148-195 `get_stone_status(&self, finished = self)` reduced to what the game record needs: per point
// 1 = counts as black territory, 2 = as white territory, 0 = neither (game_result.rs:45-93).
void territory_status(const Board& b, uint8_t* out /* [361] */) {
    Benson black(b.inner, BLACK), white(b.inner, WHITE);
    BoardFast cleaned = b.inner;                                  // clear_board (score.rs:208-222)
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
        int p = point_new(x, y);
        int c = v_color(b.inner.vertices[p]);
        if (c == WHITE && !white.is_alive(p)) v_set_color(cleaned.vertices[p], 0);
        if (c == BLACK && !black.is_alive(p)) v_set_color(cleaned.vertices[p], 0);
    }
    uint8_t db[MAXP], dw[MAXP];
    territory_distance(cleaned, BLACK, db);
    territory_distance(cleaned, WHITE, dw);
    for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
        int p = point_new(x, y);
        int c = v_color(b.inner.vertices[p]);
        uint8_t st = 0;
        if (c == WHITE) st = white.is_alive(p) ? 2 : black.is_eye(p) ? 1 : 2;      // Alive / Dead / Seki
        else if (c == BLACK) st = black.is_alive(p) ? 1 : white.is_eye(p) ? 2 : 1;
        else if (db[p] != 0xff && dw[p] == 0xff) st = 1;
        else if (dw[p] != 0xff && db[p] == 0xff) st = 2;
        out[19 * y + x] = st;
    }
}

// libdg_mcts/options.rs:180-214 (the "7 of 8 neighbours" own-eye heuristic of ScoringSearch)
bool is_vertex_filled(const Board& b, int color, int p, int dx, int dy) {
    int other = point_offset(p, dx, dy);
    return b.inner.is_part_of(other) && v_color(b.inner.vertices[other]) == color;
}
bool is_simple_eye(const Board& b, int color, int p) {
    static const int CROSS[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    static const int DIAG[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
    int nc = 0, nd = 0;
    for (auto& d : CROSS) nc += is_vertex_filled(b, color, p, d[0], d[1]);
    for (auto& d : DIAG) nd += is_vertex_filled(b, color, p, d[0], d[1]);
    int x = point_x(p), y = point_y(p);
    if ((x == 0 || x == 18) && (y == 0 || y == 18)) return nc >= 2 && nd >= 1;
    if (x == 0 || x == 18 || y == 0 || y == 18) return nc >= 3 && nd >= 2;
    return nc >= 4 && nd >= 3;
}

void ensure_init() {
    if (!g_zobrist_ready) zobrist_default();
    if (!g_sym_ready) symmetry_init();
}

}  // namespace

// =====================================================================================================================
// C interface for tests / bench (ctypes).  Points cross the boundary as packed indices 19*y + x (361 = pass / none).
// =====================================================================================================================
extern "C" {

struct dgo_board;   // opaque = Board

static inline Board* B(dgo_board* b) { return reinterpret_cast<Board*>(b); }
static inline int from_packed(int i) { return point_new(i % 19, i / 19); }   // point.rs:36-45

void dgo_set_zobrist_table(const uint64_t* table /* [3][420] */) {
    memcpy(g_zobrist, table, sizeof(g_zobrist));
    g_zobrist_ready = true;
}
void dgo_reset_zobrist_table(void) { zobrist_default(); }
dgo_board* dgo_board_new(float komi) { ensure_init(); return reinterpret_cast<dgo_board*>(new Board(komi)); }
dgo_board* dgo_board_clone(dgo_board* b) { return reinterpret_cast<dgo_board*>(new Board(*B(b))); }
void dgo_board_free(dgo_board* b) { delete B(b); }
void dgo_board_set_komi(dgo_board* b, float komi) { B(b)->komi = komi; }
float dgo_board_komi(dgo_board* b) { return B(b)->komi; }
void dgo_board_place(dgo_board* b, int color, int index) { B(b)->place(color, from_packed(index)); }
int dgo_board_is_valid(dgo_board* b, int color, int index) { return B(b)->is_valid(color, from_packed(index)); }
int dgo_board_is_valid_fast(dgo_board* b, int color, int index) { return B(b)->inner.is_valid(color, from_packed(index)); }
int dgo_board_is_ko(dgo_board* b, int color, int index) { return B(b)->is_ko(color, from_packed(index)); }
int dgo_board_at(dgo_board* b, int index) { return v_color(B(b)->inner.vertices[from_packed(index)]); }
uint64_t dgo_board_zobrist_hash(dgo_board* b) { return B(b)->zobrist_hash; }
int dgo_board_to_move(dgo_board* b) { return B(b)->to_move(); }
int dgo_board_count(dgo_board* b) { return B(b)->count; }
int dgo_board_get_n_liberty(dgo_board* b, int index) { return B(b)->inner.get_n_liberty(from_packed(index)); }
int dgo_board_get_n_liberty_if(dgo_board* b, int color, int index) { return B(b)->inner.get_n_liberty_if(color, from_packed(index)); }
int dgo_board_is_ladder_capture(dgo_board* b, int color, int index) { return is_ladder_capture(B(b)->inner, color, from_packed(index)); }
int dgo_board_is_ladder_escape(dgo_board* b, int color, int index) { return is_ladder_escape(B(b)->inner, color, from_packed(index)); }
void dgo_board_stones(dgo_board* b, uint8_t* out /* [361] 0/1/2 */) {
    for (int i = 0; i < 361; ++i) out[i] = (uint8_t)v_color(B(b)->inner.vertices[from_packed(i)]);
}
/* Board::is_valid over all 361 points (policy_helper.rs:39-43 / options.rs:53-57 for StandardSearch) */
void dgo_board_legal_mask(dgo_board* b, int color, uint8_t* out /* [361] */) {
    for (int i = 0; i < 361; ++i) out[i] = (uint8_t)B(b)->is_valid(color, from_packed(i));
}
void dgo_board_features_v1(dgo_board* b, int to_move, int symmetry, uint16_t* out /* [11552] */) {
    features_v1(*B(b), to_move, symmetry, out);
}
/* symmetry::is_symmetric (symmetry.rs:139-146) */
int dgo_board_is_symmetric(dgo_board* b, int transform) {
    for (int i = 0; i < 361; ++i) {
        int p = from_packed(i);
        if (v_color(B(b)->inner.vertices[p]) != v_color(B(b)->inner.vertices[g_sym[transform][p]])) return 0;
    }
    return 1;
}
int dgo_symmetry_apply(int transform, int index) { ensure_init(); return index == 361 ? 361 : to_packed_index(g_sym[transform][from_packed(index)]); }
int dgo_symmetry_inverse(int transform) { return SYM_INVERSE[transform]; }
/* Benson status of every point for `color`: 0 none, 1 unconditionally alive block, 2 vital region (eye). */
void dgo_board_benson(dgo_board* b, int color, uint8_t* out /* [361] */) {
    Benson bn(B(b)->inner, color);
    for (int i = 0; i < 361; ++i) out[i] = bn.points[from_packed(i)];
}
int dgo_board_is_scorable(dgo_board* b) { return is_scorable(*B(b)); }
/* Territory per point as the game record counts it (score.rs:148-195, game_result.rs:45-93): 1 black, 2 white, 0 none */
void dgo_board_territory(dgo_board* b, uint8_t* out /* [361] */) { territory_status(*B(b), out); }
/* PolicyChecker::is_policy_candidate over 0..361.  kind 0 = StandardSearch (options.rs:53-57),
 * kind 1 = ScoringSearch (options.rs:109-138). */
void dgo_board_policy_candidates(dgo_board* b, int to_move, int kind, uint8_t* out /* [362] */) {
    const Board& board = *B(b);
    if (kind == 0) {
        for (int i = 0; i < 361; ++i) out[i] = (uint8_t)board.is_valid(to_move, from_packed(i));
        out[361] = 1;
        return;
    }
    Benson black(board.inner, BLACK), white(board.inner, WHITE);
    for (int i = 0; i < 361; ++i) {
        int p = from_packed(i);
        bool ok = !black.is_eye(p) && !white.is_eye(p);
        out[i] = (uint8_t)(ok && board.is_valid(to_move, p) && !is_simple_eye(board, to_move, p));
    }
    out[361] = 0;
}
int dgo_board_is_simple_eye(dgo_board* b, int color, int index) { return is_simple_eye(*B(b), color, from_packed(index)); }
long dgo_ladder_nodes(void) { return g_ladder_nodes; }
uint16_t dgo_f32_to_f16(float f) { return f32_to_f16(f); }

/* Replays `n` moves (color, packed index; index 361 = pass, skipped as in self_play.rs:442-451) and, for every
 * ply (position BEFORE the move, player to move = the move's colour), writes the legal mask and / or the V1
 * features (Identity symmetry, as dg/bench/feature.rs:29).  Returns the number of plies written, or -(ply+1)
 * when a move is illegal (dg_tests/tests/common/mod.rs:60 asserts legality).  Single thread. */
int dgo_replay(float komi, const uint8_t* colors, const uint16_t* moves, int n, uint16_t* features_out /* [n][11552] or NULL */,
               uint8_t* legal_out /* [n][361] or NULL */, uint64_t* hash_out /* [n] or NULL: hash AFTER the move */) {
    ensure_init();
    Board board(komi);
    for (int i = 0; i < n; ++i) {
        int color = colors[i];
        if (features_out) features_v1(board, color, 0, features_out + (size_t)i * 11552);
        if (legal_out)
            for (int j = 0; j < 361; ++j) legal_out[(size_t)i * 361 + j] = (uint8_t)board.is_valid(color, from_packed(j));
        if (moves[i] < 361) {
            int p = from_packed(moves[i]);
            if (!board.is_valid(color, p)) return -(i + 1);
            board.place(color, p);
        }
        if (hash_out) hash_out[i] = board.zobrist_hash;
    }
    return n;
}

}  // extern "C"
