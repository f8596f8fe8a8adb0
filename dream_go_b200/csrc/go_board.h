// go_board.h -- host-side Go rules, ladder reader and V1 feature extraction of the self-play engine.
//
// Product code (the host half of the hot path).  It computes what
// `dg_go::Board` / `features::V1` compute (reference: src/libdg_go/board.rs, board_fast.rs,
// utils/features.rs, utils/ladder.rs, utils/symmetry.rs) but is organised for the engine, not after the
// reference:
//
//   * points are the packed indices 19*y + x the network uses (no padded border; neighbour tables instead);
//   * every chain keeps its liberties as a 361-bit set, so "liberties if played" (24 of the 32 feature
//     planes) is an OR + popcount instead of a walk over the chain with a 420-entry seen-array;
//   * ladders are only read for the points that can start one (an adjacent enemy chain with exactly two
//     liberties / an adjacent own chain in atari) -- the reference clones the board and places a stone for
//     every legal point before finding that out;
//   * features are produced directly in the engine's compact H2D format (`dg_packed_position`: one
//     32-bit plane mask per point), 16x smaller than the fp16 NHWC tensor, with the symmetry applied on
//     the fly; the fp16 tensor is expanded on the GPU (kernels.cu: pack_compact_kernel).
//
// tests/test_go_parity.py checks all of it bit-exactly against the oracle restatement of the reference.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace dg {

constexpr int N_POINTS = 361;
constexpr int PASS = 361;
constexpr int BLACK = 1, WHITE = 2;
inline int opposite(int c) { return 3 - c; }

struct Bits {                       // 361-bit set over packed indices
    uint64_t w[6];
    void clear() { w[0] = w[1] = w[2] = w[3] = w[4] = w[5] = 0; }
    void set(int i) { w[i >> 6] |= 1ull << (i & 63); }
    void reset(int i) { w[i >> 6] &= ~(1ull << (i & 63)); }
    bool test(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    int count() const {
        return __builtin_popcountll(w[0]) + __builtin_popcountll(w[1]) + __builtin_popcountll(w[2]) +
               __builtin_popcountll(w[3]) + __builtin_popcountll(w[4]) + __builtin_popcountll(w[5]);
    }
    bool any() const { return (w[0] | w[1] | w[2] | w[3] | w[4] | w[5]) != 0; }
    void or_with(const Bits& o) { for (int i = 0; i < 6; ++i) w[i] |= o.w[i]; }
    int first() const {             // lowest set bit, -1 when empty
        for (int i = 0; i < 6; ++i) if (w[i]) return 64 * i + __builtin_ctzll(w[i]);
        return -1;
    }
    Bits operator&(const Bits& o) const { Bits r; for (int i = 0; i < 6; ++i) r.w[i] = w[i] & o.w[i]; return r; }
    Bits operator|(const Bits& o) const { Bits r; for (int i = 0; i < 6; ++i) r.w[i] = w[i] | o.w[i]; return r; }
    Bits andnot(const Bits& o) const { Bits r; for (int i = 0; i < 6; ++i) r.w[i] = w[i] & ~o.w[i]; return r; }   // this & ~o
    Bits shl(int k) const {         // k in 1..63
        Bits r;
        r.w[0] = w[0] << k;
        for (int i = 1; i < 6; ++i) r.w[i] = (w[i] << k) | (w[i - 1] >> (64 - k));
        return r;
    }
    Bits shr(int k) const {
        Bits r;
        for (int i = 0; i < 5; ++i) r.w[i] = (w[i] >> k) | (w[i + 1] << (64 - k));
        r.w[5] = w[5] >> k;
        return r;
    }
};

// Static geometry: neighbours in the reference's reading order East, South(-y), West, North(+y)
// (iter/adjacent_iter.rs:42-43), the 8 dihedral maps in the order of symmetry::ALL
// (utils/symmetry.rs:121-130) and the zobrist constants.
struct Tables {
    int16_t nbr[N_POINTS][4];       // -1 = off board
    uint8_t n_nbr[N_POINTS];
    int16_t nbr_list[N_POINTS][4];  // on-board neighbours only, same order
    Bits nbr_mask[N_POINTS];
    uint16_t sym[8][N_POINTS + 1];  // sym[t][i] = where point i goes under transform t; 361 -> 361
    int32_t sym32[8][368];          // the same as gather indices, padded to whole vectors of 8 (the padding reads entry 0)
    uint8_t sym_inverse[8];
    uint64_t zobrist[3][N_POINTS];
    Bits board, not_col0, not_col18;   // all 361 points / without column x = 0 / without column x = 18

    Tables() {
        board.clear(); not_col0.clear(); not_col18.clear();
        for (int p = 0; p < N_POINTS; ++p) {
            board.set(p);
            if (p % 19 != 0) not_col0.set(p);
            if (p % 19 != 18) not_col18.set(p);
        }
        static const int dx[4] = {1, 0, -1, 0}, dy[4] = {0, -1, 0, 1};
        for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
            int p = 19 * y + x, n = 0;
            nbr_mask[p].clear();
            for (int d = 0; d < 4; ++d) {
                int xx = x + dx[d], yy = y + dy[d];
                bool on = xx >= 0 && xx < 19 && yy >= 0 && yy < 19;
                nbr[p][d] = on ? (int16_t)(19 * yy + xx) : (int16_t)-1;
                if (on) { nbr_list[p][n++] = (int16_t)(19 * yy + xx); nbr_mask[p].set(19 * yy + xx); }
            }
            n_nbr[p] = (uint8_t)n;
            for (int k = n; k < 4; ++k) nbr_list[p][k] = -1;
        }
        for (int t = 0; t < 8; ++t) {
            for (int y = 0; y < 19; ++y) for (int x = 0; x < 19; ++x) {
                int cx = x - 9, cy = y - 9, tx, ty;
                switch (t) {
                    case 0: tx = cx; ty = cy; break;       // Identity
                    case 1: tx = -cx; ty = cy; break;      // FlipLR
                    case 2: tx = cx; ty = -cy; break;      // FlipUD
                    case 3: tx = cy; ty = cx; break;       // Transpose
                    case 4: tx = -cy; ty = -cx; break;     // TransposeAnti
                    case 5: tx = cy; ty = -cx; break;      // Rot90
                    case 6: tx = -cx; ty = -cy; break;     // Rot180
                    default: tx = -cy; ty = cx; break;     // Rot270
                }
                sym[t][19 * y + x] = (uint16_t)(19 * (ty + 9) + (tx + 9));
            }
            sym[t][PASS] = PASS;
            for (int i = 0; i < 368; ++i) sym32[t][i] = i <= PASS ? sym[t][i] : 0;
        }
        static const uint8_t inv[8] = {0, 1, 2, 3, 4, 7, 6, 5};     // symmetry.rs:78-89
        memcpy(sym_inverse, inv, 8);
        // The reference's constants are arbitrary random numbers (zobrist.rs:16-17); any table gives the
        // same rules.  splitmix64 stream, laid out over the reference's padded indices so that the
        // oracle (which uses the same generator by default) produces identical hashes.
        default_zobrist();
    }
    void default_zobrist() {
        uint64_t s = 0x6472656d2d676f21ull;
        for (int c = 0; c < 3; ++c)
            for (int i = 0; i < 420; ++i) {
                uint64_t z = (s += 0x9e3779b97f4a7c15ull);
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
                z ^= z >> 31;
                int col = i % 20, row = i / 20;
                if (col >= 1 && row >= 1 && row <= 19) zobrist[c][19 * (row - 1) + (col - 1)] = z;
            }
    }
    // A table in the reference's own layout, `zobrist::TABLE: [[u64; 420]; 3]` (zobrist.rs:18) indexed by the padded
    // vertex 20 * (y + 1) + (x + 1) (board_fast.rs): with the reference's constants the hashes are the reference's.
    void load_zobrist(const uint64_t* table /* [3][420] */) {
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < 19; ++y)
                for (int x = 0; x < 19; ++x) zobrist[c][19 * y + x] = table[c * 420 + 20 * (y + 1) + (x + 1)];
    }
};

inline Tables& mutable_tables() {
    static Tables t;
    return t;
}
inline const Tables& tables() { return mutable_tables(); }

// One board = stones + chains + the history the features and the super-ko rule need
// (board.rs:27-49: last 8 moves, last 16 whole-board hashes, komi, move count, last colour).
// Chains live in slots handed out from a free list, so copying a board (one copy per tree probe, one per
// ladder step) moves only the slots in use: ~5 KB mid-game instead of 19 KB.
struct Board {
    uint8_t color[N_POINTS + 3];
    uint16_t slot[N_POINTS];        // chain slot of the stone at a point
    uint16_t next[N_POINTS];        // circular list of the stones of a chain
    uint16_t free_list[N_POINTS];
    Bits stones[3];                 // [1] black, [2] white
    Bits visited;                   // point has held a stone at some time (board_fast.rs:446, board.rs:135)
    uint64_t hash;
    uint64_t hash_history[16];
    int16_t moves[8];               // most recent moves (ring)
    uint8_t moves_pos, hash_pos, last_played, pad_;
    uint16_t count;
    uint16_t n_slots, n_free;       // slots ever handed out / currently on the free list
    float komi;
    Bits libs[N_POINTS];            // liberties of the chain in a slot; only [0, n_slots) is meaningful
    struct ChainInfo {
        uint16_t n;                 // number of liberties, kept up to date by place(); 0 = the slot is free
        uint16_t head;              // one stone of the chain
    } chain[N_POINTS];              // only [0, n_slots) is meaningful

    Board() {}
    Board(const Board& o) { copy_from(o); }
    Board& operator=(const Board& o) { if (this != &o) copy_from(o); return *this; }
    void copy_from(const Board& o) {
        memcpy(static_cast<void*>(this), static_cast<const void*>(&o), offsetof(Board, libs) + o.n_slots * sizeof(Bits));
        memcpy(chain, o.chain, o.n_slots * sizeof(ChainInfo));
    }

    void init(float komi_) {
        memset(static_cast<void*>(this), 0, sizeof(*this));
        komi = komi_;
        for (int i = 0; i < 8; ++i) moves[i] = PASS;
    }
    int to_move() const { return last_played ? opposite(last_played) : BLACK; }        // board.rs:102-107
    int n_liberty(int p) const { return chain[slot[p]].n; }
    int recent_move(int i) const { return moves[(moves_pos + 7 - i) & 7]; }
    int alloc_slot() { return n_free ? free_list[--n_free] : n_slots++; }
    void free_slot(int s) { free_list[n_free++] = (uint16_t)s; chain[s].n = 0; }

    Bits empty_neighbours(int p) const {
        const Bits& m = tables().nbr_mask[p];
        Bits e;
        for (int i = 0; i < 6; ++i) e.w[i] = m.w[i] & ~(stones[1].w[i] | stones[2].w[i]);
        return e;
    }

    // Tromp-Taylor legality without super-ko (board_fast.rs:216-243).
    bool is_valid_fast(int c, int p) const {
        if (color[p]) return false;
        const Tables& T = tables();
        for (int k = 0; k < T.n_nbr[p]; ++k) {
            int q = T.nbr_list[p][k];
            if (!color[q]) return true;
            int nl = chain[slot[q]].n;
            if ((color[q] == c) == (nl >= 2)) return true;
        }
        return false;
    }

    // Hash change if `c` played `p` (board_fast.rs:406-424): the stone plus every captured chain.
    uint64_t place_hash(int c, int p) const {
        const Tables& T = tables();
        int opp = opposite(c);
        uint64_t h = T.zobrist[c][p];
        int seen[4], ns = 0;
        for (int k = 0; k < T.n_nbr[p]; ++k) {
            int q = T.nbr_list[p][k];
            if (color[q] != opp) continue;
            int sl = slot[q];
            if (chain[sl].n >= 2) continue;
            bool dup = false;
            for (int j = 0; j < ns; ++j) dup |= seen[j] == sl;
            if (dup) continue;
            seen[ns++] = sl;
            int s = q;
            do { h ^= T.zobrist[opp][s]; s = next[s]; } while (s != q);
        }
        return h;
    }

    // Positional super-ko over the last 16 positions (board.rs:132-141).
    bool is_ko(int c, int p) const {
        if (!visited.test(p)) return false;
        uint64_t h = hash ^ place_hash(c, p);
        for (int i = 0; i < 16; ++i) if (hash_history[i] == h) return true;
        return false;
    }
    bool is_valid(int c, int p) const { return is_valid_fast(c, p) && !is_ko(c, p); }     // board.rs:151-153

    void remove_chain(int at) {
        const Tables& T = tables();
        int own = color[at];
        free_slot(slot[at]);
        int s = at;
        do {
            int nx = next[s];
            color[s] = 0;
            stones[own].reset(s);
            hash ^= T.zobrist[own][s];
            for (int k = 0; k < T.n_nbr[s]; ++k) {
                int q = T.nbr_list[s][k];
                if (color[q] && color[q] != own && !libs[slot[q]].test(s)) { libs[slot[q]].set(s); ++chain[slot[q]].n; }
            }
            s = nx;
        } while (s != at);
    }

    // Plays without any legality check (board_fast.rs:434-475, board.rs:164-188).
    void place(int c, int p) {
        const Tables& T = tables();
        int opp = opposite(c);
        int mine = alloc_slot();
        color[p] = (uint8_t)c;
        stones[c].set(p);
        visited.set(p);
        slot[p] = (uint16_t)mine;
        next[p] = (uint16_t)p;
        libs[mine].clear();
        chain[mine].n = 0;
        chain[mine].head = (uint16_t)p;
        hash ^= T.zobrist[c][p];
        for (int k = 0; k < T.n_nbr[p]; ++k) {               // enemies lose a liberty; captures first so that
            int q = T.nbr_list[p][k];                        // the freed points count as liberties below
            if (color[q] != opp) continue;
            int sl = slot[q];
            if (libs[sl].test(p)) { libs[sl].reset(p); --chain[sl].n; }
            if (chain[sl].n == 0) remove_chain(q);
        }
        libs[mine] = empty_neighbours(p);
        for (int k = 0; k < T.n_nbr[p]; ++k) {               // merge with friends
            int q = T.nbr_list[p][k];
            if (color[q] != c) continue;
            int a = slot[p], b = slot[q];
            if (a == b) continue;
            libs[b].or_with(libs[a]);
            int s = p;
            do { slot[s] = (uint16_t)b; s = next[s]; } while (s != p);
            uint16_t t = next[p]; next[p] = next[q]; next[q] = t;
            free_slot(a);
        }
        libs[slot[p]].reset(p);
        chain[slot[p]].n = (uint16_t)libs[slot[p]].count();   // the one recount of a move: the chain the stone ended up in
        last_played = (uint8_t)c;
        count += 1;
        moves[moves_pos] = (int16_t)p;
        moves_pos = (moves_pos + 1) & 7;
        hash_history[hash_pos] = hash;
        hash_pos = (hash_pos + 1) & 15;
    }

    // Liberties the stone's chain would have after `c` plays the legal move `p` (board_fast.rs:484-539),
    // and whether the move is legal at all (fused: the feature loop needs both).  Returns -1 if illegal.
    int liberties_if(int c, int p) const {
        const Tables& T = tables();
        int n_empty = 0, nf = 0, nc = 0;
        int friends[4], captured[4], captured_at[4];
        bool ok = false;
        for (int k = 0; k < T.n_nbr[p]; ++k) {
            int q = T.nbr_list[p][k];
            if (!color[q]) { ++n_empty; continue; }
            int sl = slot[q];
            int n = chain[sl].n;
            if (color[q] == c) {
                ok |= n >= 2;
                bool dup = false;
                for (int j = 0; j < nf; ++j) dup |= friends[j] == sl;
                if (!dup) friends[nf++] = sl;
            } else if (n == 1) {
                bool dup = false;
                for (int j = 0; j < nc; ++j) dup |= captured[j] == sl;
                if (!dup) { captured[nc] = sl; captured_at[nc++] = q; }
            }
        }
        if (!(ok || n_empty || nc)) return -1;
        if (!nf && !nc) return n_empty;                     // a lone stone: its empty neighbours
        if (nf == 1 && !nc) {                               // joins one chain: its liberties minus `p` plus the new ones
            const Bits& fl = libs[friends[0]];
            int n = chain[friends[0]].n - 1;
            for (int k = 0; k < T.n_nbr[p]; ++k) {
                int q = T.nbr_list[p][k];
                if (!color[q] && !fl.test(q)) ++n;
            }
            return n;
        }
        Bits L = empty_neighbours(p);
        for (int j = 0; j < nf; ++j) L.or_with(libs[friends[j]]);
        L.reset(p);
        for (int j = 0; j < nc; ++j) {                      // rare: freed points next to the new chain
            int at = captured_at[j], s = at;
            do {
                bool touches = false;
                for (int k = 0; k < T.n_nbr[s] && !touches; ++k) {
                    int t = T.nbr_list[s][k];
                    if (t == p) touches = true;
                    else if (color[t] == c)
                        for (int f = 0; f < nf; ++f) touches |= friends[f] == slot[t];
                }
                if (touches) L.set(s);
                s = next[s];
            } while (s != at);
        }
        return L.count();
    }

    // liberties_if for both colours from one scan of the neighbourhood (the feature loop asks for both at every
    // empty point; most empty points touch no stone at all).
    void liberties_if_both(int c, int p, int* mine, int* theirs) const {
        const Tables& T = tables();
        int n_empty = 0, n_stone = 0;
        for (int k = 0; k < T.n_nbr[p]; ++k) {
            int q = T.nbr_list[p][k];
            if (color[q]) ++n_stone; else ++n_empty;
        }
        if (!n_stone) { *mine = *theirs = n_empty; return; }
        *mine = liberties_if(c, p);
        *theirs = liberties_if(opposite(c), p);
    }
};

// ---- ladder reader (utils/ladder.rs) ---------------------------------------------------------------------------
// The reference recurses with a cloned board per step; here the boards of one reading live in a per-thread
// arena indexed by depth (a ladder can run 100+ steps and a Board is too large for that many stack frames).

inline bool chain_can_capture(const Board& b, int at) {      // ladder.rs:33-41
    const Tables& T = tables();
    int own = b.color[at], s = at;
    do {
        for (int k = 0; k < T.n_nbr[s]; ++k) {
            int q = T.nbr_list[s][k];
            if (b.color[q] && b.color[q] != own && b.chain[b.slot[q]].n < 2) return true;
        }
        s = b.next[s];
    } while (s != at);
    return false;
}

struct LadderArena {
    static constexpr int MAX_DEPTH = 400;   // > 361 / 2 alternating placements; deeper readings answer "no ladder"
    Board* boards;
    LadderArena() : boards(static_cast<Board*>(::operator new(sizeof(Board) * (MAX_DEPTH + 2)))) {}
    ~LadderArena() { ::operator delete(boards); }
    LadderArena(const LadderArena&) = delete;
    LadderArena& operator=(const LadderArena&) = delete;
};
inline LadderArena& ladder_arena() {
    static thread_local LadderArena arena;
    return arena;
}

// arena.boards[slot] already holds the attacker's stone at `p`.  ladder.rs:53-119.  `depth` counts the attacker's
// stones of this reading (the cut-off below); the last attacker move of a step is played on the board itself
// instead of on a copy, so `slot` grows more slowly than `depth`.
inline bool ladder_capture_after_place(LadderArena& arena, int slot, int depth, int c, int p) {
    const Tables& T = tables();
    Board& board = arena.boards[slot];
    int opp = opposite(c);
    for (;;) {
        int run = -1, nl = 0;
        for (int k = 0; k < T.n_nbr[p] && run < 0; ++k) {
            int q = T.nbr_list[p][k];
            if (board.color[q] != opp) continue;
            int sl = board.slot[q];
            if (board.chain[sl].n >= 2 || chain_can_capture(board, q)) continue;
            int lib = board.libs[sl].first();                // in atari: its only liberty
            if (lib < 0) continue;
            nl = board.liberties_if(opp, lib);               // -1: the extension is not a legal move
            if (nl >= 0) run = lib;
        }
        if (run < 0) return false;
        if (nl < 2) return true;                             // the liberties after extending decide before the stone is placed
        if (nl >= 3) return false;
        board.place(opp, run);
        for (int k = 0; k < T.n_nbr[run]; ++k) {
            int q = T.nbr_list[run][k];
            if (board.color[q] == c && board.n_liberty(q) < 2) return false;
        }
        if (depth + 1 >= LadderArena::MAX_DEPTH) return false;
        int tries[4], nt = 0;
        for (int k = 0; k < T.n_nbr[run]; ++k) {
            int q = T.nbr_list[run][k];
            if (board.is_valid_fast(c, q)) tries[nt++] = q;
        }
        if (nt == 0) return false;
        for (int t = 0; t + 1 < nt; ++t) {
            Board& child = arena.boards[slot + 1];
            child.copy_from(board);
            child.place(c, tries[t]);
            if (ladder_capture_after_place(arena, slot + 1, depth + 1, c, tries[t])) return true;
        }
        p = tries[nt - 1];                                   // nobody looks at this board again: continue on it
        board.place(c, p);
        ++depth;
    }
}

// Where a ladder can start: the liberties of enemy chains with exactly two liberties (the move puts them in atari) and
// of own chains in atari (the move extends them) -- one pass over the chain slots, no stone is looked at.
inline void ladder_starts(const Board& b, int to_move, Bits& capture_at, Bits& escape_at) {
    capture_at.clear();
    escape_at.clear();
    for (int sl = 0; sl < b.n_slots; ++sl) {
        const int n = b.chain[sl].n;
        if (n == 0 || n > 2) continue;                       // a free slot / nothing starts here
        if (b.color[b.chain[sl].head] == to_move) { if (n < 2) escape_at.or_with(b.libs[sl]); }
        else if (n == 2) capture_at.or_with(b.libs[sl]);
    }
}

// The first step of ladder_capture_after_place read off the ORIGINAL board: four of five readings end there (the
// chain put in atari extends to three or more liberties), and a board copy costs more than the rest of such a reading.
// Exact whenever neither the attacker's stone at `p` nor the extension captures anything; otherwise, and when the
// ladder goes on, the caller reads it on a copy.  Returns 0 = no ladder, 1 = captured, 2 = read it on a copy.
inline int ladder_capture_first_step(const Board& b, int c, int p) {
    const Tables& T = tables();
    const int opp = opposite(c);
    int merged[4], nm = 0;                                   // own chains the stone at `p` joins
    Bits lm = b.empty_neighbours(p);                         // liberties of the joined chain
    for (int k = 0; k < T.n_nbr[p]; ++k) {
        int q = T.nbr_list[p][k];
        if (b.color[q] == opp) { if (b.chain[b.slot[q]].n < 2) return 2; }          // `p` captures: not handled here
        else if (b.color[q] == c) { merged[nm++] = b.slot[q]; lm.or_with(b.libs[b.slot[q]]); }
    }
    lm.reset(p);
    if (lm.count() < 2) return 0;        // every chain next to `p` could capture the attacker's own stones (chain_can_capture)
    auto in_merged = [&](int sl) { for (int j = 0; j < nm; ++j) if (merged[j] == sl) return true; return false; };
    for (int k = 0; k < T.n_nbr[p]; ++k) {
        int q = T.nbr_list[p][k];
        if (b.color[q] != opp) continue;
        const int sx = b.slot[q];
        if (b.chain[sx].n != 2) continue;                    // two liberties, `p` is one: in atari after the stone
        bool can_capture = false;                            // chain_can_capture on the board after the stone: the joined
        int s = q;                                           // chain has >= 2 liberties, every other chain is as it was
        do {
            for (int j = 0; j < T.n_nbr[s] && !can_capture; ++j) {
                int t = T.nbr_list[s][j];
                if (b.color[t] == c && !in_merged(b.slot[t]) && b.chain[b.slot[t]].n < 2) can_capture = true;
            }
            s = b.next[s];
        } while (s != q && !can_capture);
        if (can_capture) continue;
        Bits only = b.libs[sx];
        only.reset(p);
        const int lib = only.first();
        // liberties_if(opp, lib) on the board after the stone
        Bits l = b.empty_neighbours(lib);
        l.reset(p);
        bool legal = l.any();
        for (int j = 0; j < T.n_nbr[lib]; ++j) {
            int u = T.nbr_list[lib][j];
            if (u == p) continue;                            // the joined chain: >= 2 liberties, not captured
            if (b.color[u] == c) {
                if (!in_merged(b.slot[u]) && b.chain[b.slot[u]].n < 2) return 2;        // the extension captures: not handled here
            } else if (b.color[u] == opp) {
                const Bits& fl = b.libs[b.slot[u]];
                legal |= b.chain[b.slot[u]].n - (fl.test(p) ? 1 : 0) >= 2;
                l.or_with(fl);
            }
        }
        if (!legal) continue;                                // the extension would be suicide: look at the next chain
        l.reset(p);
        l.reset(lib);
        const int n = l.count();
        return n < 2 ? 1 : n >= 3 ? 0 : 2;
    }
    return 0;
}

inline bool is_ladder_capture(const Board& b, int c, int p) {                                     // ladder.rs:131-135
    // a ladder starts by putting an adjacent enemy chain in atari: it needs exactly two liberties now
    const Tables& T = tables();
    int opp = opposite(c);
    bool candidate = false;
    for (int k = 0; k < T.n_nbr[p]; ++k) {
        int q = T.nbr_list[p][k];
        if (b.color[q] == opp && b.chain[b.slot[q]].n == 2) candidate = true;
    }
    if (!candidate) return false;
    if (!b.color[p]) {
        const int first = ladder_capture_first_step(b, c, p);
        if (first != 2) return first == 1;
    }
    LadderArena& arena = ladder_arena();
    arena.boards[0].copy_from(b);
    arena.boards[0].place(c, p);
    return ladder_capture_after_place(arena, 0, 0, c, p);
}

inline bool is_ladder_escape(const Board& b, int c, int p) {                                      // ladder.rs:144-178
    const Tables& T = tables();
    bool in_atari = false;
    for (int k = 0; k < T.n_nbr[p]; ++k) {
        int q = T.nbr_list[p][k];
        if (b.color[q] == c && b.chain[b.slot[q]].n < 2) in_atari = true;
    }
    if (!in_atari) return false;
    if (b.liberties_if(c, p) != 2) return false;             // the callers only ask about legal moves
    LadderArena& arena = ladder_arena();
    Board& board = arena.boards[0];
    board.copy_from(b);
    board.place(c, p);
    int opp = opposite(c);
    for (int k = 0; k < T.n_nbr[p]; ++k) {
        int q = T.nbr_list[p][k];
        if (!board.is_valid_fast(opp, q)) continue;
        Board& child = arena.boards[1];
        child.copy_from(board);
        child.place(opp, q);
        if (ladder_capture_after_place(arena, 1, 1, opp, q)) return false;
    }
    return true;
}

// ---- unconditional life (utils/benson.rs, utils/flood_fill.rs) and what self-play builds on it --------------------
// Benson's algorithm on bitsets: a region is an empty-seeded connected component of the points not occupied by
// `color` (benson.rs:293-320), it is vital to a chain when every one of its points touches the chain
// (:188-208), a chain stays alive with two vital regions (:95-111) and a region stays while all the stones
// around it are alive (:115-131).
// All points adjacent to a point of `x` (the set itself is not included unless it is adjacent to itself).
inline Bits dilate(const Bits& x) {
    const Tables& T = tables();
    Bits r = x.shl(19) | x.shr(19) | (x.shl(1) & T.not_col0) | (x.shr(1) & T.not_col18);
    return r & T.board;
}

inline void benson(const Board& b, int color, Bits& alive, Bits& eyes) {
    const Tables& T = tables();
    alive.clear();
    eyes.clear();
    const Bits& own = b.stones[color];
    if (!own.any()) return;
    // A region with a point that touches no stone of `color` is vital to nobody and is dropped before anything
    // else happens (benson.rs:128-143): flood those regions away with whole-board bit operations first.  What is
    // left -- usually nothing before the endgame -- are the small enclosed regions the algorithm is about.
    const Bits other = T.board.andnot(own);                  // empty or enemy
    Bits dead = other.andnot(dilate(own));
    if (dead.any())
        for (;;) {
            Bits grow = (dilate(dead) & other).andnot(dead);
            if (!grow.any()) break;
            dead.or_with(grow);
        }
    const Bits rest = other.andnot(dead);
    Bits seeds = rest.andnot(b.stones[opposite(color)]);     // regions start from empty points (benson.rs:297-301)
    if (!seeds.any()) return;
    struct Region { Bits points, around; };
    struct Chain { Bits stones, touch; };
    Region regions[N_POINTS / 2 + 1];
    int nr = 0;
    Bits around_all;
    around_all.clear();
    while (seeds.any()) {                                    // what is left are small enclosed regions: flood them point by
        Region& r = regions[nr];                             // point (a whole-board shift per step costs more than they do)
        r.points.clear();
        r.around.clear();
        int stack[N_POINTS], top = 0;
        stack[top++] = seeds.first();
        r.points.set(stack[0]);
        while (top) {
            const int s = stack[--top];
            for (int k = 0; k < T.n_nbr[s]; ++k) {
                const int q = T.nbr_list[s][k];
                if (own.test(q)) r.around.set(q);            // never empty: every point of `rest` touches `own`
                else if (rest.test(q) && !r.points.test(q)) { r.points.set(q); stack[top++] = q; }
            }
        }
        around_all.or_with(r.around);
        seeds = seeds.andnot(r.points);
        ++nr;
    }
    if (nr < 2) return;                                      // a chain needs two vital regions (benson.rs:95-111)
    Chain chains[N_POINTS / 2 + 1];                          // only chains next to a region can ever qualify
    int nc = 0;
    {
        uint8_t slot_seen[N_POINTS];
        memset(slot_seen, 0, b.n_slots);
        Bits todo = around_all;
        while (todo.any()) {
            int p = todo.first();
            todo.reset(p);
            if (slot_seen[b.slot[p]]) continue;
            slot_seen[b.slot[p]] = 1;
            Chain& c = chains[nc++];
            c.stones.clear();
            c.touch.clear();
            int s = p;
            do { c.stones.set(s); c.touch.or_with(T.nbr_mask[s]); s = b.next[s]; } while (s != p);
        }
    }
    auto vital = [](const Region& r, const Chain& c) {       // every point of the region touches the chain (:188-208)
        uint64_t miss = 0;
        for (int i = 0; i < 6; ++i) miss |= r.points.w[i] & ~c.touch.w[i];
        return miss == 0;
    };
    {   // regions that are vital to nobody go first (benson.rs:128-143)
        int keep = 0;
        for (int i = 0; i < nr; ++i) {
            bool v = false;
            for (int j = 0; j < nc && !v; ++j) v = vital(regions[i], chains[j]);
            if (v) { if (keep != i) regions[keep] = regions[i]; ++keep; }
        }
        nr = keep;
    }
    for (;;) {
        bool changed = false;
        int keep = 0;
        for (int j = 0; j < nc; ++j) {
            int n = 0;
            for (int i = 0; i < nr; ++i) n += vital(regions[i], chains[j]);
            if (n >= 2) { if (keep != j) chains[keep] = chains[j]; ++keep; } else changed = true;
        }
        nc = keep;
        alive.clear();
        for (int j = 0; j < nc; ++j) alive.or_with(chains[j].stones);
        keep = 0;
        for (int i = 0; i < nr; ++i) {
            uint64_t miss = 0;
            for (int k = 0; k < 6; ++k) miss |= regions[i].around.w[k] & ~alive.w[k];
            if (!miss) { if (keep != i) regions[keep] = regions[i]; ++keep; } else changed = true;
        }
        nr = keep;
        if (!changed) break;
    }
    alive.clear();
    for (int j = 0; j < nc; ++j) alive.or_with(chains[j].stones);
    for (int i = 0; i < nr; ++i) eyes.or_with(regions[i].points);
}

// Score::is_scorable (utils/score.rs:97-110): every point is settled by unconditional life.
inline bool is_scorable(const Board& b) {
    Bits alive[3], eyes[3];
    benson(b, BLACK, alive[BLACK], eyes[BLACK]);
    benson(b, WHITE, alive[WHITE], eyes[WHITE]);
    for (int p = 0; p < N_POINTS; ++p) {
        int c = b.color[p];
        bool ok = c == 0 ? (eyes[BLACK].test(p) || eyes[WHITE].test(p)) : (alive[c].test(p) || eyes[opposite(c)].test(p));
        if (!ok) return false;
    }
    return true;
}

// Whose territory the game record counts each point as (utils/score.rs:148-195 with `finished` = the board itself,
// game_result.rs:45-93): stones belong to their owner unless they sit dead inside an unconditionally alive enemy eye;
// empty points belong to the colour whose unconditionally alive stones alone reach them.  1 black, 2 white, 0 none.
inline void territory_status(const Board& b, uint8_t out[N_POINTS]) {
    Bits alive[3], eyes[3];
    benson(b, BLACK, alive[BLACK], eyes[BLACK]);
    benson(b, WHITE, alive[WHITE], eyes[WHITE]);
    Bits reach[3];
    const Bits open = tables().board.andnot(alive[BLACK] | alive[WHITE]);   // the cleaned board: everything else is empty
    for (int c = BLACK; c <= WHITE; ++c) {
        reach[c] = alive[c];
        for (;;) {
            Bits grow = (dilate(reach[c]) & open).andnot(reach[c]);
            if (!grow.any()) break;
            reach[c].or_with(grow);
        }
    }
    for (int p = 0; p < N_POINTS; ++p) {
        int c = b.color[p];
        if (c) out[p] = (uint8_t)(alive[c].test(p) ? c : eyes[opposite(c)].test(p) ? opposite(c) : c);
        else out[p] = (uint8_t)(reach[BLACK].test(p) && !reach[WHITE].test(p) ? BLACK : reach[WHITE].test(p) && !reach[BLACK].test(p) ? WHITE : 0);
    }
}

// The own-eye heuristic of ScoringSearch (libdg_mcts/options.rs:180-214).
inline bool is_simple_eye(const Board& b, int color, int p) {
    const Tables& T = tables();
    // corner needs 2 of 2, edge 3 of 3, middle 4 of 4 orthogonal neighbours: all that exist
    for (int k = 0; k < T.n_nbr[p]; ++k) if (b.color[T.nbr_list[p][k]] != color) return false;
    int x = p % 19, y = p / 19, diag = 0;
    for (int dy = -1; dy <= 1; dy += 2) for (int dx = -1; dx <= 1; dx += 2) {
        int xx = x + dx, yy = y + dy;
        if (xx >= 0 && xx <= 18 && yy >= 0 && yy <= 18 && b.color[19 * yy + xx] == color) ++diag;
    }
    return diag >= (T.n_nbr[p] == 2 ? 1 : T.n_nbr[p] == 3 ? 2 : 3);
}

enum SearchKind { STANDARD_SEARCH = 0, SCORING_SEARCH = 1 };

// PolicyChecker::is_policy_candidate for all 362 moves (options.rs:53-57 and :109-138).  `legal` = Board::is_valid
// of `to_move` for the 361 points (as features_v1 returns it).
inline void policy_candidates(const Board& b, int to_move, int kind, const uint8_t* legal, uint8_t out[N_POINTS + 1]) {
    if (kind == STANDARD_SEARCH) {
        memcpy(out, legal, N_POINTS);
        out[PASS] = 1;
        return;
    }
    Bits alive, eyes_b, eyes_w;
    benson(b, BLACK, alive, eyes_b);
    benson(b, WHITE, alive, eyes_w);
    const Bits settled = eyes_b | eyes_w;                    // usually empty before the endgame
    if (settled.any()) {
        for (int p = 0; p < N_POINTS; ++p) out[p] = legal[p] && !settled.test(p);
    } else {
        memcpy(out, legal, N_POINTS);
    }
    // own eyes: only points whose on-board neighbours are all own stones can be one -- whole-board shifts find them
    const Tables& T = tables();
    Bits maybe_eye = T.board.andnot(dilate(T.board.andnot(b.stones[to_move])));
    while (maybe_eye.any()) {
        const int p = maybe_eye.first();
        maybe_eye.reset(p);
        if (out[p] && is_simple_eye(b, to_move, p)) out[p] = 0;
    }
    out[PASS] = 0;
}

// ---- V1 features in the compact format (utils/features.rs:154-250) ---------------------------------------------

inline uint16_t f32_to_f16_bits(float f) {                  // round to nearest even (fp16.rs:64-68)
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, e8 = (x >> 23) & 0xff, man = x & 0x7fffffu;
    if (e8 == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    int e = (int)e8 - 112;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - e;
        uint32_t h = man >> shift, rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)e << 10) | (man >> 13), rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

inline uint16_t komi_plane_bits(float komi) {               // features.rs:236
    float k = 0.5f + (0.5f * komi) / 7.5f;
    k = k > 1.0f ? 1.0f : k;
    k = k < 0.0f ? 0.0f : k;
    return f32_to_f16_bits(k);
}

// planes[sym(p)] bit c <=> feature (p, c) non-zero.  Also returns the legal-move mask (Board::is_valid,
// i.e. with super-ko) of `to_move` in `legal` when it is not null -- the prior construction
// (pool/policy_helper.rs:39-43) needs exactly the valid/ko bits this pass computes anyway.
inline void features_v1(const Board& b, int to_move, int symmetry, uint32_t planes[N_POINTS], uint16_t* k_bits,
                        uint8_t* legal /* [361] or null, identity orientation */) {
    const Tables& T = tables();
    const uint16_t* sym = T.sym[symmetry];
    const int opp = opposite(to_move);
    uint32_t global = to_move == BLACK ? 1u : 2u;
    // where a ladder can start: the liberties of enemy chains with exactly two liberties (the move puts them in
    // atari) / of own chains in atari (the move extends them) -- straight from the chains' liberty sets
    Bits capture_at, escape_at;
    ladder_starts(b, to_move, capture_at, escape_at);
    const Bits touched = dilate(b.stones[1] | b.stones[2]);  // points next to a stone
    uint32_t local[N_POINTS];
    bool any_ko = false;
    static const uint32_t ge_mask[7] = {0, 1, 3, 7, 15, 31, 63};   // ">= 1 .. >= n" thermometer code
    for (int p = 0; p < N_POINTS; ++p) {
        uint32_t m = 0;
        int col = b.color[p];
        if (col) {
            int n = b.chain[b.slot[p]].n;
            m = ge_mask[n > 6 ? 6 : n] << (col == to_move ? 5 : 17);
            if (legal) legal[p] = 0;
        } else {
            int mine, theirs;
            if (!touched.test(p)) mine = theirs = T.n_nbr[p];     // open point: its neighbours are its liberties
            else {
                mine = b.liberties_if(to_move, p);
                theirs = b.liberties_if(opp, p);
            }
            if (mine >= 0) m |= ge_mask[mine > 6 ? 6 : mine] << 11;
            if (theirs >= 0) m |= ge_mask[theirs > 6 ? 6 : theirs] << 23;
            bool ko = false;
            if (mine >= 0) {
                ko = b.is_ko(to_move, p);
                if (ko) { m |= 1u << 29; any_ko = true; }
                if (capture_at.test(p) && is_ladder_capture(b, to_move, p)) m |= 1u << 30;
                if (escape_at.test(p) && is_ladder_escape(b, to_move, p)) m |= 1u << 31;
            }
            if (legal) legal[p] = mine >= 0 && !ko;
        }
        local[p] = m;
    }
    int m0 = b.recent_move(0), m1 = b.recent_move(1);
    if (m0 != PASS) local[m0] |= 1u << 3;
    if (m1 != PASS) local[m1] |= 1u << 4;
    if (any_ko) global |= 4u;
    for (int p = 0; p < N_POINTS; ++p) planes[sym[p]] = local[p] | global;
    *k_bits = komi_plane_bits(b.komi);
}

// ---- raw position for the device feature kernel (csrc/features.cu) ----------------------------------------------------
// Everything the device needs to rebuild the planes: stones, visited bits, hashes, last moves -- plus the two ladder
// planes, which are read here (sequential search) for the points that can start a ladder and are legal, unless bit 3 of
// `symmetry` (DG_RAW_DEVICE_LADDERS) leaves them to the device's own reader (csrc/features.cu: namespace lad).
template <class Raw>
inline void raw_position(const Board& b, int to_move, int symmetry, Raw* out) {
    static_assert(sizeof(Bits) == 48, "Bits is 12 x u32");
    memcpy(out->black, b.stones[BLACK].w, 48);
    memcpy(out->white, b.stones[WHITE].w, 48);
    memcpy(out->visited, b.visited.w, 48);
    Bits capture_at, escape_at, capture, escape;
    capture.clear(); escape.clear();
    capture_at.clear(); escape_at.clear();
    if (!(symmetry & 8)) ladder_starts(b, to_move, capture_at, escape_at);       // bit 3 (DG_RAW_DEVICE_LADDERS): the device reads them
    Bits todo = capture_at | escape_at;
    while (todo.any()) {
        int p = todo.first();
        todo.reset(p);
        if (!b.is_valid_fast(to_move, p)) continue;
        if (capture_at.test(p) && is_ladder_capture(b, to_move, p)) capture.set(p);
        if (escape_at.test(p) && is_ladder_escape(b, to_move, p)) escape.set(p);
    }
    memcpy(out->ladder_capture, capture.w, 48);
    memcpy(out->ladder_escape, escape.w, 48);
    out->hash = b.hash;
    for (int i = 0; i < 16; ++i) out->hash_history[i] = b.hash_history[i];
    out->last_move[0] = (int16_t)b.recent_move(0);
    out->last_move[1] = (int16_t)b.recent_move(1);
    out->k_bits = komi_plane_bits(b.komi);
    out->to_move = (uint8_t)to_move;
    out->symmetry = (uint8_t)symmetry;
}

}  // namespace dg
