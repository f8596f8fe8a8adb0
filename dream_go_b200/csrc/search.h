// search.h -- Monte-Carlo tree search of the self-play engine (host side, product code).
//
// Computes what `dg_mcts::tree` / `dg_mcts::predict` compute (reference: src/libdg_mcts/tree.rs:1027-1510,
// lib.rs:83-200, time_control/mod.rs:47-97, choose.rs, dirichlet.rs, libdg_utils/config.rs:181-336, lcb.rs), with
// the data organised for the batched engine instead of a pool of racing worker threads:
//
//   * a node stores only its candidate moves (finite prior) as sorted (move, prior) pairs plus one small record
//     per VISITED child, field by field in parallel arrays that `select` scores eight children at a time (AVX2) --
//     the reference keeps a dense prior[368] and an 8-slot / 362-slot child table (tree.rs:540-700);
//   * a search advances in rounds: up to P probes are collected (virtual loss keeps them apart), their leaves are
//     evaluated as ONE batch together with the leaves of every other game on the device, then inserted in probe
//     order.  No locks, no atomics, and the result is a pure function of (position, weights, seed) -- the
//     reference's result depends on thread timing.  With P = 1 this is exactly the sequential algorithm.
//
// Arithmetic is fp32 in the reference's operation order (compiled with -ffp-contract=off) so that visit counts
// are bit-identical to the oracle restatement (tests/test_mcts_parity.py).
#pragma once

#include <immintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "go_board.h"

namespace dg {

constexpr int VLOSS_CNT = 32;                 // config.rs:187
constexpr int MIN_LCB_VISITS = 80;            // tree.rs:34
constexpr float NEG_INF = -std::numeric_limits<float>::infinity();

// Piecewise-linear schedules over the visit count of the CURRENT node (config.rs:181-192, 297-336).
struct Knot { int x; float y; };
inline float interpolate(const Knot* pts, int n, int x) {
    int i = 0;
    while (i < n && pts[i].x < x) ++i;
    if (i == n) return pts[n - 1].y;
    const Knot& x0 = pts[i == 0 ? 0 : i - 1];
    const Knot& x1 = pts[i];
    float a = x0.x >= x1.x ? 0.5f : (float)(x - x0.x) / (float)(x1.x - x0.x);
    return (1.0f - a) * x0.y + a * x1.y;
}
inline float uct_exp(int visits) {
    static const Knot k[5] = {{0, 0.77392f}, {800, 1.05439f}, {1600, 1.22798f}, {3200, 0.813532f}, {6400, 0.764326f}};
    return interpolate(k, 5, visits);
}
inline float fpu_reduce(int visits) {
    static const Knot k[5] = {{0, 0.631571f}, {800, 0.431547f}, {1600, 0.656083f}, {3200, 0.429231f}, {6400, 0.514494f}};
    return interpolate(k, 5, visits);
}
inline float lcb_critical_value(int visits) {
    static const Knot k[5] = {{0, 1.91753f}, {800, 1.86478f}, {1600, 1.86943f}, {3200, 2.20033f}, {6400, 1.78053f}};
    return interpolate(k, 5, visits);
}

struct Node;

// Index of `x` in a[0..n) or -1; `a` is readable up to the next multiple of 16 entries (unused entries hold 0xffff).
inline int find_u16(const uint16_t* a, int n, int x) {
    const __m256i key = _mm256_set1_epi16((short)x);
    for (int i = 0; i < n; i += 16) {
        uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi16(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + i)), key));
        if (m) { int k = i + (__builtin_ctz(m) >> 1); return k < n ? k : -1; }
    }
    return -1;
}

// permutation that moves the lanes selected by an 8-bit mask to the front, in order (AVX2 has no compress instruction)
struct CompressLut {
    alignas(32) int32_t idx[256][8];
    CompressLut() {
        for (int m = 0; m < 256; ++m) {
            int k = 0;
            for (int b = 0; b < 8; ++b) if (m >> b & 1) idx[m][k++] = b;
            for (; k < 8; ++k) idx[m][k] = 0;
        }
    }
};
inline const CompressLut& compress_lut() { static const CompressLut t; return t; }

// Candidate blocks all have the size of the largest (362 candidates), and a search frees and allocates thousands of them
// per move: each thread keeps the blocks it frees for its next allocations (glibc's own per-thread cache stops at 1 KiB).
constexpr size_t kCandBlock = (362 + 8) * 4 + 384 * 2 + 362 + 16;
struct CandBlockCache {
    std::vector<void*> blocks;
    ~CandBlockCache() { for (void* b : blocks) free(b); }
    void* get() {
        if (blocks.empty()) return malloc(kCandBlock);
        void* b = blocks.back();
        blocks.pop_back();
        return b;
    }
    void put(void* b) {
        if (!b) return;
        if (blocks.size() < 8192) blocks.push_back(b); else free(b);      // a thread that mostly frees keeps <= 21 MB
    }
};
inline CandBlockCache& cand_blocks() { static thread_local CandBlockCache c; return c; }

// A node keeps (a) its candidates and (b) one record per visited (or disqualified) child, in creation order.  Both are
// stored as parallel arrays in one allocation each: the root of a search has hundreds of children and `select` scores
// them eight at a time; records never move to another index, so a probe remembers the index instead of the move.
struct Node {
    uint8_t to_move;
    int16_t pass_count;
    float initial_value;
    int32_t total_count, vtotal_count;
    // Candidates (finite prior).  The first `sorted_n` are in DECREASING prior order (ties: increasing move); the
    // rest is unordered and gets selection-sorted only as far as a select() actually walks -- most nodes are
    // visited once or twice, so a full sort per node would cost more than the rest of the search.
    int n_cand = 0, sorted_n = 0;
    int first_untouched = 0;                   // every sorted candidate before this one has a child record
    float* cand_prior = nullptr;               // [n_cand]
    uint16_t* cand_move = nullptr;             // [n_cand rounded up to 16]
    uint8_t* cand_edge = nullptr;              // [n_cand] 1 = the candidate has a child record
    // child records
    int n_edges = 0, edge_cap = 0;
    float *e_prior = nullptr;                  // -inf = not (or no longer) a candidate
    float *e_value = nullptr, *e_value_s = nullptr;
    int32_t *e_count = nullptr, *e_vcount = nullptr;
    Node** e_child = nullptr;
    uint16_t* e_move = nullptr;
    uint8_t* e_expanding = nullptr;

    Node(int to_move_, float value, const float* prior /* [362] */)
        : to_move((uint8_t)to_move_), pass_count(0), initial_value(value), total_count(0), vtotal_count(0) {
        set_prior(prior);
    }
    ~Node() {
        for (int i = 0; i < n_edges; ++i) delete e_child[i];
        cand_blocks().put(cand_prior);
        free(e_child);
    }
    Node(const Node&) = delete;
    Node& operator=(const Node&) = delete;

    void set_prior(const float* prior) {       // lib.rs:170-182 replaces the prior of a re-used root
        // `prior` has 362 entries; the callers' buffers are not padded, so the last 2 are counted one by one
        const __m256 inf = _mm256_set1_ps(std::numeric_limits<float>::infinity());
        const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
        int n = 0;
        for (int i = 0; i < 360; i += 8)
            n += __builtin_popcount((unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_and_ps(_mm256_loadu_ps(prior + i), absmask), inf, _CMP_LT_OQ)));
        n += std::isfinite(prior[360]) + std::isfinite(prior[361]);
        // the compaction below stores whole vectors of 8 at the running position: 8 spare slots in both arrays
        const int n16 = (n + 8 + 15) & ~15;
        static_assert(kCandBlock >= (362 + 8) * 4 + ((362 + 8 + 15) & ~15) * 2 + 362 + 16, "a block holds the largest node");
        char* mem = cand_prior ? reinterpret_cast<char*>(cand_prior) : static_cast<char*>(cand_blocks().get());
        cand_prior = reinterpret_cast<float*>(mem);
        cand_move = reinterpret_cast<uint16_t*>(mem + (size_t)(n + 8) * 4);
        cand_edge = reinterpret_cast<uint8_t*>(mem + (size_t)(n + 8) * 4 + (size_t)n16 * 2);
        n_cand = n;
        n = 0;
        const CompressLut& lut = compress_lut();
        __m256i moves = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
        const __m256i eight = _mm256_set1_epi32(8);
        for (int i = 0; i < 360; i += 8) {         // keep the finite entries (and their indices), in order
            const __m256 x = _mm256_loadu_ps(prior + i);
            const unsigned keep = (unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_and_ps(x, absmask), inf, _CMP_LT_OQ));
            const __m256i order = _mm256_load_si256(reinterpret_cast<const __m256i*>(lut.idx[keep]));
            _mm256_storeu_ps(cand_prior + n, _mm256_permutevar8x32_ps(x, order));
            const __m256i m = _mm256_permutevar8x32_epi32(moves, order);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(cand_move + n), _mm_packus_epi32(_mm256_castsi256_si128(m), _mm256_extracti128_si256(m, 1)));
            n += __builtin_popcount(keep);
            moves = _mm256_add_epi32(moves, eight);
        }
        for (int i = 360; i < 362; ++i)
            if (std::isfinite(prior[i])) { cand_move[n] = (uint16_t)i; cand_prior[n] = prior[i]; ++n; }
        for (int i = n; i < n16; ++i) cand_move[i] = 0xffff;
        memset(cand_edge, 0, (size_t)n_cand);
        sorted_n = 0;
        first_untouched = 0;
        // The first visits of a node ask for its best few candidates one at a time, each time by a scan of arrays that have
        // gone cold in between: find the best four now, in one pass, while the arrays are in the cache.  (Candidates are in
        // increasing move order here, so among equal priors the earlier one ranks higher: strict comparisons.)
        if (n_cand > 4) {
            int top[4] = {-1, -1, -1, -1};
            auto offer = [&](int i) {
                const float x = cand_prior[i];
                if (top[3] >= 0 && !(x > cand_prior[top[3]])) return;
                int r = 3;
                while (r > 0 && (top[r - 1] < 0 || x > cand_prior[top[r - 1]])) { top[r] = top[r - 1]; --r; }
                top[r] = i;
            };
            for (int i = 0; i < 4; ++i) offer(i);
            int i = 4;
            for (; i + 8 <= n_cand; i += 8) {          // eight at a time against the fourth best so far; hits are rare
                unsigned hit = (unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(cand_prior + i), _mm256_set1_ps(cand_prior[top[3]]), _CMP_GT_OQ));
                for (; hit; hit &= hit - 1) offer(i + __builtin_ctz(hit));
            }
            for (; i < n_cand; ++i) offer(i);
            for (int r = 0; r < 4; ++r) {
                const int j = top[r];
                std::swap(cand_prior[r], cand_prior[j]);
                std::swap(cand_move[r], cand_move[j]);
                for (int t = r + 1; t < 4; ++t) if (top[t] == r) top[t] = j;     // what stood at r now stands at j
            }
            sorted_n = 4;
        }
        for (int k = 0; k < n_edges; ++k) {
            e_prior[k] = std::isfinite(prior[e_move[k]]) ? prior[e_move[k]] : NEG_INF;
            int c = cand_index(e_move[k]);
            if (c >= 0) cand_edge[c] = 1;
        }
    }
    void sort_up_to(int k) {                   // makes positions [0, k] final
        const int n = n_cand;
        while (sorted_n <= k && sorted_n < n) {
            int best = sorted_n;
            for (int i = sorted_n + 1; i < n; ++i)
                if (cand_prior[i] > cand_prior[best] || (cand_prior[i] == cand_prior[best] && cand_move[i] < cand_move[best])) best = i;
            std::swap(cand_move[sorted_n], cand_move[best]);
            std::swap(cand_prior[sorted_n], cand_prior[best]);
            std::swap(cand_edge[sorted_n], cand_edge[best]);
            ++sorted_n;
        }
    }
    int cand_index(int move) const { return find_u16(cand_move, n_cand, move); }
    int find(int move) const { return find_u16(e_move, n_edges, move); }       // index of the child record or -1
    void grow() {
        const int cap = edge_cap ? 2 * edge_cap : 16;                          // multiples of 16: find_u16 / the 8-wide scoring read whole blocks
        char* mem = static_cast<char*>(malloc((size_t)cap * (8 + 5 * 4 + 2 + 1)));
        Node** child = reinterpret_cast<Node**>(mem);
        float* prior = reinterpret_cast<float*>(mem + (size_t)cap * 8);
        float* value = prior + cap;
        float* value_s = value + cap;
        int32_t* count = reinterpret_cast<int32_t*>(value_s + cap);
        int32_t* vcount = count + cap;
        uint16_t* move = reinterpret_cast<uint16_t*>(vcount + cap);
        uint8_t* expanding = reinterpret_cast<uint8_t*>(move + cap);
        if (n_edges) {
            memcpy(child, e_child, (size_t)n_edges * 8);
            memcpy(prior, e_prior, (size_t)n_edges * 4);
            memcpy(value, e_value, (size_t)n_edges * 4);
            memcpy(value_s, e_value_s, (size_t)n_edges * 4);
            memcpy(count, e_count, (size_t)n_edges * 4);
            memcpy(vcount, e_vcount, (size_t)n_edges * 4);
            memcpy(move, e_move, (size_t)n_edges * 2);
            memcpy(expanding, e_expanding, (size_t)n_edges);
        }
        for (int i = n_edges; i < cap; ++i) {                                  // unused records never match and never score
            move[i] = 0xffff;
            prior[i] = NEG_INF;
            value[i] = 0.0f;
            count[i] = vcount[i] = 0;
        }
        free(e_child);
        e_child = child; e_prior = prior; e_value = value; e_value_s = value_s; e_count = count; e_vcount = vcount;
        e_move = move; e_expanding = expanding;
        edge_cap = cap;
    }
    int edge(int move) {                       // tree.rs:276-285: an absent child reads as (count 0, value = initial)
        int k = find(move);
        if (k >= 0) return k;
        if (n_edges == edge_cap) grow();
        int c = cand_index(move);
        if (c >= 0) cand_edge[c] = 1;
        k = n_edges++;
        e_move[k] = (uint16_t)move;
        e_expanding[k] = 0;
        e_prior[k] = c >= 0 ? cand_prior[c] : NEG_INF;
        e_count[k] = 0;
        e_vcount[k] = 0;
        e_value[k] = initial_value;
        e_value_s[k] = 0.0f;
        e_child[k] = nullptr;
        return k;
    }
    float prior_of(int move) const { int c = cand_index(move); return c >= 0 ? cand_prior[c] : NEG_INF; }
    int count_of(int move) const { int k = find(move); return k >= 0 ? e_count[k] : 0; }
    float value_of(int move) const { int k = find(move); return k >= 0 ? e_value[k] : initial_value; }

    void disqualify(int move) {                // tree.rs:1296-1301
        int k = edge(move);
        e_value[k] = NEG_INF;
        e_count[k] = 0;
    }
};

enum ProbeStatus { PROBE_FOUND = 0, PROBE_CONFLICT = 1, PROBE_NO_RESULT = 2 };
struct TraceEntry { Node* node; int move; int edge; };      // edge = index of the child record of `move` in `node`
typedef std::vector<TraceEntry> Trace;

// Node::select (tree.rs:1311-1385) and asm/argmax.rs:23-76: the reference scans 368 scores in blocks of 8; among
// equal maxima the LAST block wins and inside a block the FIRST lane.
inline ProbeStatus select(Node& node, bool apply_fpu, int* out_move, int* out_edge) {
    const int n = node.total_count + node.vtotal_count;
    const float sqrt_n = std::sqrt((float)(1 + n));
    const float u = uct_exp(n) * sqrt_n;
    float unvisited = node.initial_value;
    float reduce = 0.0f;
    if (apply_fpu) {
        reduce = fpu_reduce(n);
        float v = unvisited - reduce;
        unvisited = v > 0.0f ? v : 0.0f;       // _mm256_max_ps(v, 0): also maps -inf / NaN to 0
    }
    float best = NEG_INF;
    int best_move = -1;
    auto consider = [&](int move, float score) {   // equal scores: highest block of 8, then lowest index
        if (best_move < 0 || score > best ||
            (score == best && ((move >> 3) > (best_move >> 3) || ((move >> 3) == (best_move >> 3) && move < best_move)))) {
            best = score;
            best_move = move;
        }
    };
    // children with a record: their own value / visit count (a record whose move is not a candidate any more has
    // prior -inf in the reference and never wins).  score = value + prior * (u / (1 + count + vcount)), where an
    // unvisited child's value is reduced (first-play urgency) -- eight records at a time, IEEE operation for operation
    // what the scalar code computes (u / 1 == u, so the division needs no special case for unvisited children).
    {
        const __m256 vu = _mm256_set1_ps(u), vreduce = _mm256_set1_ps(reduce), zero = _mm256_setzero_ps();
        const __m256 ninf = _mm256_set1_ps(NEG_INF);
        const __m256i one = _mm256_set1_epi32(1);
        auto score8 = [&](int k0, __m256* live) {
            const __m256i total = _mm256_add_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(node.e_count + k0)),
                                                   _mm256_loadu_si256(reinterpret_cast<const __m256i*>(node.e_vcount + k0)));
            const __m256 bonus = _mm256_div_ps(vu, _mm256_cvtepi32_ps(_mm256_add_epi32(total, one)));
            __m256 value = _mm256_loadu_ps(node.e_value + k0);
            if (apply_fpu) {
                const __m256 fresh = _mm256_castsi256_ps(_mm256_cmpeq_epi32(total, _mm256_setzero_si256()));
                value = _mm256_blendv_ps(value, _mm256_max_ps(_mm256_sub_ps(value, vreduce), zero), fresh);
            }
            const __m256 prior = _mm256_loadu_ps(node.e_prior + k0);
            *live = _mm256_cmp_ps(prior, ninf, _CMP_GT_OQ);          // unused records (prior -inf) are never live
            return _mm256_add_ps(value, _mm256_mul_ps(prior, bonus));
        };
        bool done = false;
        if (node.n_edges > 16) {
            // Many records (the root): keep all scores, find the maximum with vector operations and apply the tie rule to
            // the records that reach it -- without NaNs the order (score, block of 8 descending, lane ascending) is total,
            // so that is what considering every record in turn gives.
            alignas(32) float sc[512];
            __m256 vmax = ninf;
            int nan = 0, any_live = 0;
            for (int k0 = 0; k0 < node.n_edges; k0 += 8) {
                __m256 live;
                const __m256 raw8 = score8(k0, &live);
                const __m256 s8 = _mm256_blendv_ps(ninf, raw8, live);
                nan |= _mm256_movemask_ps(_mm256_cmp_ps(s8, s8, _CMP_UNORD_Q));
                any_live |= _mm256_movemask_ps(live);
                _mm256_store_ps(sc + k0, s8);
                vmax = _mm256_max_ps(vmax, s8);
            }
            if (!nan) {
                done = true;
                if (any_live) {
                    alignas(32) float lanes[8];
                    _mm256_store_ps(lanes, vmax);
                    float m = lanes[0];
                    for (int j = 1; j < 8; ++j) m = lanes[j] > m ? lanes[j] : m;
                    const __m256 vm = _mm256_set1_ps(m);
                    for (int k0 = 0; k0 < node.n_edges; k0 += 8) {
                        const __m256 live = _mm256_cmp_ps(_mm256_loadu_ps(node.e_prior + k0), ninf, _CMP_GT_OQ);
                        int hit = _mm256_movemask_ps(_mm256_and_ps(_mm256_cmp_ps(_mm256_load_ps(sc + k0), vm, _CMP_EQ_OQ), live));
                        for (; hit; hit &= hit - 1) {
                            const int k = k0 + __builtin_ctz((unsigned)hit);
                            consider(node.e_move[k], sc[k]);
                        }
                    }
                }
            }
        }
        if (!done) {
            float score[8];
            for (int k0 = 0; k0 < node.n_edges; k0 += 8) {
                __m256 live8;
                const __m256 raw8 = score8(k0, &live8);
                _mm256_storeu_ps(score, raw8);
                const int live = _mm256_movemask_ps(live8);
                const int kn = std::min(8, node.n_edges - k0);
                for (int j = 0; j < kn; ++j)
                    if (live >> j & 1) consider(node.e_move[k0 + j], score[j]);
            }
        }
    }
    // untouched children all score `unvisited + prior * u`, which never increases along the prior-sorted list:
    // stop at the first one that falls strictly below the best score seen
    const int nc = node.n_cand;
    int i = node.first_untouched;
    while (i < node.sorted_n && node.cand_edge[i]) ++i;
    node.first_untouched = i;
    for (; i < nc; ++i) {
        node.sort_up_to(i);
        if (node.cand_edge[i]) continue;
        float score = unvisited + node.cand_prior[i] * u;
        if (best_move >= 0 && score < best) break;
        consider(node.cand_move[i], score);
    }
    if (best_move < 0 || !std::isfinite(best)) return PROBE_NO_RESULT;
    const int k = node.edge(best_move);
    bool was_expanding = node.e_expanding[k];
    node.e_expanding[k] = 1;
    if (was_expanding && !node.e_child[k]) return PROBE_CONFLICT;
    node.e_vcount[k] += VLOSS_CNT;
    node.vtotal_count += VLOSS_CNT;
    *out_move = best_move;
    *out_edge = k;
    return PROBE_FOUND;
}

inline void undo(const Trace& trace, bool undo_expanding) {    // tree.rs:1397-1409
    for (const TraceEntry& t : trace) {
        t.node->vtotal_count -= VLOSS_CNT;
        t.node->e_vcount[t.edge] -= VLOSS_CNT;
        if (undo_expanding && !t.node->e_child[t.edge]) t.node->e_expanding[t.edge] = 0;
    }
}

// tree::probe (tree.rs:1421-1471): walks down, playing the moves on `board`.
inline ProbeStatus probe(Node& root, Board& board, Trace& trace) {
    trace.clear();
    Node* current = &root;
    for (;;) {
        int move, k;
        ProbeStatus st = select(*current, !trace.empty(), &move, &k);
        if (st == PROBE_CONFLICT) { undo(trace, false); trace.clear(); return st; }
        if (st == PROBE_NO_RESULT) return st;
        trace.push_back(TraceEntry{current, move, k});
        if (move != PASS) board.place(current->to_move, move);
        else if (current->pass_count >= 1) break;
        Node* child = current->e_child[k];
        if (!child) break;
        current = child;
    }
    return PROBE_FOUND;
}

// tree::insert + UCT::update (tree.rs:1482-1510, 125-159)
inline void insert(const Trace& trace, int color, float value, const float* prior /* [362] */) {
    if (!trace.empty()) {
        const TraceEntry& last = trace.back();
        if (!last.node->e_child[last.edge]) {
            Node* next = new Node(color, value, prior);
            if (last.move == PASS) next->pass_count = (int16_t)(last.node->pass_count + 1);
            last.node->e_child[last.edge] = next;
        }
    }
    for (const TraceEntry& t : trace) {
        float v = color == t.node->to_move ? value : 1.0f - value;
        Node& nd = *t.node;
        const int k = t.edge;
        nd.total_count += 1;
        nd.vtotal_count -= VLOSS_CNT;
        float prev = nd.e_value[k], prev_s = nd.e_value_s[k];
        int prev_count = nd.e_count[k];
        nd.e_count[k] = prev_count + 1;
        float next = prev + (v - prev) / (float)(prev_count + 1);
        nd.e_value[k] = next;
        nd.e_value_s[k] = prev_s + (v - prev) * (v - next);
        nd.e_vcount[k] -= VLOSS_CNT;
    }
}

// time_control::is_done with RolloutLimit (time_control/mod.rs:47-97, rollout_limit.rs:35-45)
inline bool is_done(const Node& root, int limit) {
    if (root.total_count == 0) return false;
    if (root.total_count >= limit) return true;
    int remaining = limit - root.total_count;
    // min_promote_rollouts (:47-74): visits the runner-up needs to catch up = largest count - second largest
    // (whichever of two tied leaders the reference's argmax picks, the difference is the same)
    int c1 = 0, c2 = 0;
    for (int k = 0; k < root.n_edges; ++k) {
        const int c = root.e_count[k];
        if (c > c1) { c2 = c1; c1 = c; }
        else if (c > c2) c2 = c;
    }
    int min_promote = c1 > c2 ? c1 - c2 : 0;
    return min_promote > remaining;
}

// ---- choosing the move ------------------------------------------------------------------------------------------

inline float normal_lcb(float p_hat, float p_std, int n, int m) {      // libdg_utils/lcb.rs:28-36
    if (n <= 0) return 0.0f;
    float z = lcb_critical_value(m);
    return p_hat - z * p_std / std::sqrt((float)n);
}

// tree.rs:1524-1560; returns <0, 0, >0 like Ordering
inline int compare_children(const Node& node, int a, int b) {
    const int ea = node.find(a), eb = node.find(b);
    int ac = ea >= 0 ? node.e_count[ea] : 0, bc = eb >= 0 ? node.e_count[eb] : 0;
    auto cmp = [](float x, float y) { return x < y ? -1 : x > y ? 1 : 0; };
    if (ac >= MIN_LCB_VISITS && bc >= MIN_LCB_VISITS) {
        float as = std::sqrt(node.e_value_s[ea] / ((float)ac + 1e-5f)), bs = std::sqrt(node.e_value_s[eb] / ((float)bc + 1e-5f));
        float al = normal_lcb(node.e_value[ea], as, ac, node.total_count), bl = normal_lcb(node.e_value[eb], bs, bc, node.total_count);
        if (al != bl) return cmp(al, bl);
    }
    if (ac != bc) return ac < bc ? -1 : 1;
    float ap = node.prior_of(a), bp = node.prior_of(b);
    if (ap != bp) return cmp(ap, bp);
    return cmp(node.value_of(a), node.value_of(b));
}

// choose.rs:62-99 (+ percentile :25-48).  items are visit counts (or a policy); returns the index or -1.
inline int choose(const std::vector<double>& items, double cutoff_percentile, double temperature, double at) {
    double total = 0.0;
    for (double x : items) if (std::isfinite(x)) total += x;
    std::vector<int> order(items.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return items[a] < items[b]; });   // only values matter below
    double max_value = total * (1.0 - cutoff_percentile), so_far = 0.0, threshold = 0.0;
    bool found = false;
    for (size_t k = order.size(); k-- > 0;) {
        so_far += items[order[k]];
        if (so_far >= max_value) { threshold = items[order[k]]; found = true; break; }
    }
    if (!found) return -1;
    double cum_total = 0.0;
    std::vector<double> cum(items.size(), std::numeric_limits<double>::quiet_NaN());
    for (size_t i = 0; i < items.size(); ++i)
        if (items[i] >= threshold) { cum_total += std::pow(items[i] / so_far, temperature); cum[i] = cum_total; }
    double target = at * cum_total;
    for (size_t i = 0; i < items.size(); ++i) if (cum[i] >= target) return (int)i;
    return -1;
}

// Node::best (tree.rs:1232-1283).  `at` = the uniform random number the stochastic branch draws.
inline int best(const Node& node, float temperature, double at, float* value_out) {
    int pick;
    if (temperature <= 9e-2f) {
        // children.nonzero() (tree.rs:871-890) walks the 8-slot table in insertion order while the node is
        // sparse and the dense table in index order once a 9th child exists; it only matters for exact ties
        std::vector<int> visited;
        for (int k = 0; k < node.n_edges; ++k) if (node.e_count[k] != 0) visited.push_back(node.e_move[k]);
        if (node.n_edges > 8) std::sort(visited.begin(), visited.end());
        pick = PASS;
        bool first = true;
        for (int i : visited) {                // Iterator::max_by keeps the LAST of equal maxima
            if (first || compare_children(node, pick, i) <= 0) { pick = i; first = false; }
        }
        *value_out = node.value_of(pick);
        return pick;
    }
    std::vector<double> visits(362);
    for (int i = 0; i < 362; ++i) visits[i] = (double)node.count_of(i);
    pick = choose(visits, 0.5, 1.0 / (double)temperature, at);
    if (pick < 0) { *value_out = node.initial_value; return PASS; }
    *value_out = node.value_of(pick);
    return pick;
}

// Node::forward (tree.rs:1198-1225): detaches and returns the sub-tree of `move` (nullptr if there is none);
// the rest of `node` is destroyed.
inline Node* forward(Node* node, int move) {
    Node* next = nullptr;
    if (int k = node->find(move); k >= 0) { next = node->e_child[k]; node->e_child[k] = nullptr; }
    if (!next && move == PASS) {
        float prior[362];
        for (int i = 0; i < 362; ++i) prior[i] = 0.0f;
        next = new Node(opposite(node->to_move), 0.5f, prior);
        next->pass_count = (int16_t)(node->pass_count + 1);
    }
    delete node;
    return next;
}

// Node::softmax (tree.rs:1275-1289): visit distribution of the root.
inline void visit_distribution(const Node& node, float out[362]) {
    float total = 0.0f;
    for (int i = 0; i < 362; ++i) out[i] = 0.0f;
    std::vector<int> visited;
    for (int k = 0; k < node.n_edges; ++k) if (node.e_count[k] != 0) visited.push_back(node.e_move[k]);
    std::sort(visited.begin(), visited.end());
    for (int i : visited) total += (float)node.count_of(i);
    for (int i : visited) out[i] = (float)node.count_of(i) / total;
}

// ---- root prior: average over the 8 symmetries (lib.rs:83-133) and Dirichlet noise (dirichlet.rs:40-76) ----------

// dirichlet::add_ex with the normalised sample `eta` supplied by the caller (eta[i] = g_i / sum g over the finite entries)
inline void mix_noise(float* x /* [362] */, const float* eta, float beta) {
    for (int i = 0; i < 362; ++i)
        if (std::isfinite(x[i])) x[i] = (1.0f - beta) * x[i] + beta * eta[i];
}

// ---- randomness (injected everywhere; the reference uses thread_rng) ------------------------------------------------
struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed = 1) { reseed(seed); }
    void reseed(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            uint64_t z = (seed += 0x9e3779b97f4a7c15ull);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {                          // xoshiro256**
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }       // [0, 1)
    int below(int n) { return (int)(uniform() * n); }
    double normal() {
        for (;;) {                             // Marsaglia polar method
            double u = 2.0 * uniform() - 1.0, v = 2.0 * uniform() - 1.0, q = u * u + v * v;
            if (q > 0.0 && q < 1.0) return u * std::sqrt(-2.0 * std::log(q) / q);
        }
    }
    double gamma(double shape) {               // Marsaglia & Tsang; shape < 1 boosted by U^(1/shape) (as rand_distr does)
        if (shape < 1.0) {
            double u = uniform();
            while (u <= 0.0) u = uniform();
            return gamma(shape + 1.0) * std::pow(u, 1.0 / shape);
        }
        double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = uniform();
            if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
            if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
        }
    }
    // eta for mix_noise: Gamma(shape) samples over the finite entries of x, normalised (dirichlet.rs:48-70)
    void dirichlet(const float* x, double shape, float* eta) {
        double g[362], sum;
        int count;
        do {
            sum = 0.0;
            count = 0;
            for (int i = 0; i < 362; ++i) {
                g[i] = 0.0;
                if (std::isfinite(x[i])) { g[i] = gamma(shape); sum += g[i]; ++count; }
            }
        } while (count != 0 && !(sum > std::numeric_limits<double>::min()));
        for (int i = 0; i < 362; ++i) eta[i] = count ? (float)(g[i] / sum) : 0.0f;
    }
};

}  // namespace dg
