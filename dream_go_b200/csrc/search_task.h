// search_task.h -- one tree search as a resumable task that EMITS leaf positions and ABSORBS their evaluations.
//
// `dg_mcts::predict` (src/libdg_mcts/lib.rs:145-200) blocks a game thread while a pool of workers probes its tree
// (pool/worker_thread.rs:60-177).  Here a search never blocks: the self-play driver asks every game for its next
// leaves, evaluates all of them as one device batch, and hands the results back.  Phases:
//
//   ROOT     8 positions = the root under all 8 symmetries (full_forward, lib.rs:83-133)
//   PROBING  up to `probes_per_round` leaves per round (tree::probe -> Event::predict, pool/event.rs:46-59)
//   DONE     move chosen (Node::best, lib.rs:190-194)
#pragma once

#include <immintrin.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/dg_engine.h"
#include "search.h"

namespace dg {

inline float f16_to_f32(uint16_t h) { return _cvtsh_ss(h); }     // exact (F16C); the build targets x86-64-v3

// Transposition table of network evaluations: `NnPredictor::{fetch, cache}` + `LruCache` (predictors/nn.rs:29-82,
// lru_cache.rs).  Key = (zobrist hash, colour to move); the value is kept in IDENTITY orientation
// (`cache` stores with_transform(response, symmetry.inverse()), `fetch` returns with_transform(entry, symmetry),
// predictor.rs:30-44), so one evaluation answers the position under every symmetry.  `get` makes the entry the most
// recent one, `insert` of an existing key does nothing, the least recent entry is dropped beyond `capacity`.
// The reference keeps ONE table of 200,000 entries for the whole process behind one mutex (nn.rs:48-50).  Here a table
// has `stripes` independent LRU lists, each behind its own lock (a key lives in the stripe its hash selects): one stripe
// is exactly the reference's table; many stripes are the process-wide table that all games and worker threads share
// without queueing on one lock (eviction is then least-recent within a stripe).  A per-game single-stripe table keeps the
// games a function of the seed; a shared one, like the reference's, makes them depend on the other games' timing.
class PredictionCache {
  public:
    struct Entry {
        uint64_t hash;
        uint16_t value;
        uint8_t to_move;
        uint16_t policy[362];
        int32_t prev, next;
    };
    explicit PredictionCache(size_t capacity, int stripes = 1) {
        size_t n = (size_t)(stripes < 1 ? 1 : stripes);
        if (capacity > 0 && capacity < n) n = capacity;
        for (size_t i = 0; i < n; ++i) {
            stripes_.emplace_back(new Stripe());
            stripes_.back()->capacity = capacity / n;
        }
    }
    size_t size() const {
        size_t n = 0;
        for (const auto& st : stripes_) { std::lock_guard<std::mutex> g(st->m); n += st->map.size(); }
        return n;
    }
    std::atomic<long> hits{0}, misses{0};

    // Copies the entry out (the table may drop it the moment the lock is released); true on a hit.
    bool get(uint64_t hash, int to_move, Entry* out) {
        Stripe& st = stripe(hash);
        std::lock_guard<std::mutex> g(st.m);
        auto it = st.map.find(Key{hash, (uint8_t)to_move});
        if (it == st.map.end()) { misses.fetch_add(1, std::memory_order_relaxed); return false; }
        hits.fetch_add(1, std::memory_order_relaxed);
        st.detach(it->second);
        st.attach(it->second);
        *out = st.pool[it->second];
        return true;
    }
    // `policy` is in orientation `symmetry`; it is stored un-transformed
    void insert(uint64_t hash, int to_move, int symmetry, uint16_t value, const uint16_t* policy) {
        Stripe& st = stripe(hash);
        if (st.capacity == 0) return;
        Key key{hash, (uint8_t)to_move};
        std::lock_guard<std::mutex> g(st.m);
        if (st.map.count(key)) return;
        int32_t idx;
        if (!st.free.empty()) { idx = st.free.back(); st.free.pop_back(); }
        else { idx = (int32_t)st.pool.size(); st.pool.emplace_back(); }
        Entry& e = st.pool[idx];
        e.hash = hash;
        e.to_move = (uint8_t)to_move;
        e.value = value;
        const uint16_t* inv = tables().sym[tables().sym_inverse[symmetry]];
        for (int i = 0; i < N_POINTS; ++i) e.policy[inv[i]] = policy[i];     // with_transform(response, symmetry.inverse())
        e.policy[PASS] = policy[PASS];
        st.attach(idx);
        st.map.emplace(key, idx);
        if (st.map.size() > st.capacity) {
            int32_t t = st.tail;
            st.detach(t);
            st.map.erase(Key{st.pool[t].hash, st.pool[t].to_move});
            st.free.push_back(t);
        }
    }

  private:
    struct Key {
        uint64_t hash;
        uint8_t to_move;
        bool operator==(const Key& o) const { return hash == o.hash && to_move == o.to_move; }
    };
    struct KeyHash { size_t operator()(const Key& k) const { return (size_t)(k.hash * 0x9e3779b97f4a7c15ull) ^ k.to_move; } };
    struct Stripe {
        mutable std::mutex m;
        size_t capacity = 0;
        std::vector<Entry> pool;
        std::vector<int32_t> free;
        std::unordered_map<Key, int32_t, KeyHash> map;
        int32_t head = -1, tail = -1;
        void attach(int32_t i) {
            pool[i].prev = -1;
            pool[i].next = head;
            if (head >= 0) pool[head].prev = i;
            head = i;
            if (tail < 0) tail = i;
        }
        void detach(int32_t i) {
            Entry& e = pool[i];
            if (e.prev >= 0) pool[e.prev].next = e.next;
            if (e.next >= 0) pool[e.next].prev = e.prev;
            if (head == i) head = e.next;
            if (tail == i) tail = e.prev;
        }
    };
    Stripe& stripe(uint64_t hash) {
        return stripes_.size() == 1 ? *stripes_[0] : *stripes_[(size_t)((hash * 0xd6e8feb86659fd93ull) >> 32) % stripes_.size()];
    }
    std::vector<std::unique_ptr<Stripe>> stripes_;
};

// What create_initial_policy (pool/policy_helper.rs:28-75) derives from the board alone; computed when the leaf is
// emitted (the feature pass already produced the legal mask), applied when its evaluation arrives.
struct PriorPlan {
    uint8_t candidate[368];                    // after symmetry elimination; [362, 368) stays 0 (whole vectors of 8 below)
    uint16_t rep[N_POINTS + 1];                // orbit representative of each point (`indices`, :54-72)
    bool folded;                               // the board has a symmetry: some rep[p] != p

    void build(const Board& b, int to_move, int search_kind, const uint8_t* legal) {
        const Tables& T = tables();
        policy_candidates(b, to_move, search_kind, legal, candidate);
        memset(candidate + N_POINTS + 1, 0, 368 - (N_POINTS + 1));
        int syms[8], ns = 1;
        syms[0] = 0;                               // the identity
        for (int t = 1; t < 8; ++t) {
            bool same = true;
            const uint16_t* m = T.sym[t];
            for (int p = 0; p < N_POINTS && same; ++p) same = b.color[p] == b.color[m[p]];
            if (same) syms[ns++] = t;
        }
        folded = ns > 1;
        if (!folded) return;                       // every point is its own representative: apply() does not look at `rep`
        for (int p = 0; p < N_POINTS; ++p) {
            int best = p;
            for (int k = 0; k < ns; ++k) { int q = T.sym[syms[k]][p]; if (q < best) best = q; }
            rep[p] = (uint16_t)best;
            if (best != p) candidate[p] = 0;
        }
        rep[PASS] = PASS;
    }

    // add_valid_candidates + normalize_policy (policy_helper.rs:87-134; lane order of asm/sum_finite.rs:23-57)
    void apply(const uint16_t* policy /* [362] fp16, orientation `symmetry` */, int symmetry, float sum_to, float* prior /* [368] */) const {
        const Tables& T = tables();
        if (!folded) {
            // every point is its own representative: prior[q] = 0 + policy[sym(q)] for the candidates, -inf elsewhere
            // i = fwd[q]  <=>  q = inverse(symmetry)[i]; fp16 -> fp32 for the whole vector first, then one gather per 8 points
            alignas(32) float p32[368];
            for (int i = 0; i < 360; i += 8)
                _mm256_store_ps(p32 + i, _mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(policy + i))));
            p32[360] = f16_to_f32(policy[360]);
            p32[361] = f16_to_f32(policy[361]);
            const int32_t* fwd = T.sym32[symmetry];
            const __m256 zero = _mm256_setzero_ps(), ninf = _mm256_set1_ps(NEG_INF);
            for (int q = 0; q < 368; q += 8) {
                const __m256 x = _mm256_i32gather_ps(p32, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(fwd + q)), 4);
                const __m256i c8 = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(candidate + q)));
                const __m256 keep = _mm256_castsi256_ps(_mm256_cmpgt_epi32(c8, _mm256_setzero_si256()));
                _mm256_storeu_ps(prior + q, _mm256_blendv_ps(ninf, _mm256_add_ps(zero, x), keep));
            }
        } else {
            for (int i = 0; i < 368; ++i) prior[i] = NEG_INF;
            for (int p = 0; p <= N_POINTS; ++p) if (candidate[p]) prior[p] = 0.0f;
            prior[PASS] += f16_to_f32(policy[PASS]);
            const uint16_t* inv = T.sym[T.sym_inverse[symmetry]];
            for (int i = 0; i < N_POINTS; ++i) prior[rep[inv[i]]] += f16_to_f32(policy[i]);
        }
        // eight interleaved partial sums over the finite entries, as the reference's AVX code keeps them; a masked-out
        // entry adds +0.0, which leaves a partial sum (never -0.0) unchanged
        const __m256 inf = _mm256_set1_ps(std::numeric_limits<float>::infinity());
        const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
        __m256 acc = _mm256_setzero_ps();
        int finite = 0;
        for (int i = 0; i < 368; i += 8) {
            const __m256 x = _mm256_loadu_ps(prior + i);
            const __m256 ok = _mm256_cmp_ps(_mm256_and_ps(x, absmask), inf, _CMP_LT_OQ);     // finite: not inf, not NaN
            acc = _mm256_add_ps(acc, _mm256_and_ps(x, ok));
            finite += __builtin_popcount((unsigned)_mm256_movemask_ps(ok));
        }
        float lane[8];
        _mm256_storeu_ps(lane, acc);
        float sum = ((lane[0] + lane[1]) + (lane[2] + lane[3])) + ((lane[4] + lane[5]) + (lane[6] + lane[7]));
        if (sum < 1e-6f) {
            for (int i = 0; i < 368; ++i) if (std::isfinite(prior[i])) prior[i] = sum_to / (float)finite;
        } else {
            const __m256 recip = _mm256_set1_ps(1.0f / (sum / sum_to));
            for (int i = 0; i < 368; i += 8) _mm256_storeu_ps(prior + i, _mm256_mul_ps(_mm256_loadu_ps(prior + i), recip));
        }
    }
};

// DG_SELFPLAY_TRACE: cycles spent in the phases of a probe (summed over all threads; relaxed atomics, off by default)
struct PhaseClock {
    std::atomic<uint64_t> copy{0}, probe{0}, extract{0}, plan{0}, apply{0}, insert{0}, leaves{0};
    bool on = false;
};
inline PhaseClock& phase_clock() { static PhaseClock c; return c; }
struct PhaseTimer {
    uint64_t t;
    bool on;
    PhaseTimer() : t(0), on(phase_clock().on) { if (on) t = __rdtsc(); }
    void lap(std::atomic<uint64_t>& into) { if (on) { uint64_t n = __rdtsc(); into.fetch_add(n - t, std::memory_order_relaxed); t = n; } }
};

struct SearchOptions {
    int search_kind = STANDARD_SEARCH;         // which PolicyChecker (options.rs)
    bool deterministic = false;                // SearchOptions::deterministic()
    int num_rollout = 800;                     // RolloutLimit
    int probes_per_round = 1;
    float dirichlet_beta = 0.25f;              // config.rs:165-166 (self-play)
    float dirichlet_shape = 0.03f;             // lib.rs:164
    float temperature = 0.8f;                  // config.rs:171-172 (self-play)
    const float* noise = nullptr;              // injected eta[362] (tests); else drawn from the task's Rng
    const uint8_t* leaf_symmetries = nullptr;  // injected sequence (tests); else drawn from the task's Rng
    int n_leaf_symmetries = 0;
    double choose_at = -1.0;                   // injected uniform number for the stochastic move choice
    bool policy_only = false;                  // no tree: play from the averaged policy (self_play.rs:360-396)
    PredictionCache* cache = nullptr;          // transposition table (predictors/nn.rs:29-82); null = none
    bool device_ladders = false;               // raw positions: the device reads the ladders (DG_RAW_DEVICE_LADDERS)
};

class SearchTask {
  public:
    enum Phase { ROOT, PROBING, DONE };

    SearchTask() {}
    ~SearchTask() { delete root_; }
    SearchTask(const SearchTask&) = delete;
    SearchTask& operator=(const SearchTask&) = delete;

    // `starting_tree` (may be null) is consumed, as in lib.rs:150.
    void start(const Board& board, int color, const SearchOptions& opt, Node* starting_tree, uint64_t seed) {
        delete root_;
        root_ = starting_tree;
        board_ = board;
        color_ = color;
        opt_ = opt;
        rng_.reseed(seed);
        phase_ = ROOT;
        pending_.clear();
        n_pending_ = 0;
        leaf_counter_ = 0;
        evals_ = 0;
        cache_inserts_ = 0;
        value_ = 0.5f;
        index_ = PASS;
    }

    Phase phase() const { return phase_; }
    bool done() const { return phase_ == DONE; }
    float value() const { return value_; }
    int index() const { return index_; }
    long evals() const { return evals_; }
    long cache_inserts() const { return cache_inserts_; }
    const Node* root() const { return root_; }
    Node* take_root() { Node* r = root_; root_ = nullptr; return r; }
    bool policy_only() const { return opt_.policy_only; }
    const float* root_policy() const { return root_policy_; }

    // Appends this round's leaf positions and returns how many; 0 with done() == true ends the search.
    // Exactly one of `packed` / `raw` is given: compact feature planes computed here on the host, or raw positions
    // whose planes and legal moves the device derives (csrc/features.cu) -- then absorb() receives the legal masks.
    int emit(std::vector<dg_packed_position>* packed, std::vector<dg_raw_position>* raw = nullptr) {
        if (phase_ == DONE) return 0;
        if (phase_ == ROOT && opt_.cache) {
            // full_forward (lib.rs:97-111): a cached evaluation answers all 8 symmetries without the network
            PredictionCache::Entry entry;
            const PredictionCache::Entry* hit = opt_.cache->get(board_.hash, color_, &entry) ? &entry : nullptr;
            if (hit) opt_.cache->hits += 7; else opt_.cache->misses += 7;      // the reference asks once per symmetry
            if (hit) {
                uint8_t legal[N_POINTS];
                for (int p = 0; p < N_POINTS; ++p) legal[p] = (uint8_t)board_.is_valid(color_, p);
                root_plan_.build(board_, color_, opt_.search_kind, legal);
                uint16_t value[8], policy[8 * 362];
                const Tables& T = tables();
                for (int t = 0; t < 8; ++t) {              // Prediction::with_transform(entry, t)
                    value[t] = hit->value;
                    for (int i = 0; i < N_POINTS; ++i) policy[t * 362 + T.sym[t][i]] = hit->policy[i];
                    policy[t * 362 + PASS] = hit->policy[PASS];
                }
                n_pending_ = 0;
                root_cached_ = true;
                absorb(value, policy, nullptr);
                root_cached_ = false;
                if (phase_ == DONE) return 0;
            }
        }
        if (phase_ == ROOT) {
            if (raw) {
                size_t at = raw->size();
                raw->resize(at + 8);
                const int extra = (opt_.search_kind << 4) | (opt_.device_ladders ? 8 : 0);
                raw_position(board_, color_, extra, &(*raw)[at]);
                for (int t = 1; t < 8; ++t) { (*raw)[at + t] = (*raw)[at]; (*raw)[at + t].symmetry = (uint8_t)(t | extra); }
            } else {
                uint8_t legal[N_POINTS];
                size_t at = packed->size();
                packed->resize(at + 8);
                for (int t = 0; t < 8; ++t) {
                    dg_packed_position& pos = (*packed)[at + t];
                    features_v1(board_, color_, t, pos.planes, &pos.k_bits, t == 0 ? legal : nullptr);
                    pos.reserved = 0;
                }
                root_plan_.build(board_, color_, opt_.search_kind, legal);
            }
            n_pending_ = 8;
            return 8;
        }
        int emitted = 0, hits_this_round = 0;
        if (pending_.size() < (size_t)opt_.probes_per_round) pending_.resize(opt_.probes_per_round);
        while (emitted < opt_.probes_per_round) {
            if (is_done(*root_, opt_.num_rollout)) break;
            Pending& p = pending_[emitted];
            PhaseTimer pt;
            p.board.copy_from(board_);
            pt.lap(phase_clock().copy);
            ProbeStatus st = probe(*root_, p.board, p.trace);
            pt.lap(phase_clock().probe);
            if (st == PROBE_CONFLICT) break;                 // somebody (an earlier probe of this round) is expanding it
            if (st == PROBE_NO_RESULT) break;
            p.to_move = opposite(p.trace.back().node->to_move);
            p.symmetry = next_leaf_symmetry();
            if (opt_.cache) {                                // Event::predict (pool/event.rs:50-52): fetch before extracting
                PredictionCache::Entry entry;
                if (const PredictionCache::Entry* hit = opt_.cache->get(p.board.hash, p.to_move, &entry) ? &entry : nullptr) {
                    uint8_t legal[N_POINTS];
                    for (int q = 0; q < N_POINTS; ++q) legal[q] = (uint8_t)p.board.is_valid(p.to_move, q);
                    p.plan.build(p.board, p.to_move, opt_.search_kind, legal);
                    float prior[368];
                    p.plan.apply(hit->policy, 0, 1.0f, prior);   // the entry is in identity orientation
                    insert(p.trace, p.to_move, 0.5f * f16_to_f32(hit->value) + 0.5f, prior);
                    ++cache_inserts_;
                    if (++hits_this_round > 4 * opt_.probes_per_round) break;   // leave the round eventually
                    continue;
                }
            }
            if (raw) {
                raw->emplace_back();
                raw_position(p.board, p.to_move, p.symmetry | (opt_.search_kind << 4) | (opt_.device_ladders ? 8 : 0), &raw->back());
            } else {
                uint8_t legal[N_POINTS];
                packed->emplace_back();
                dg_packed_position& pos = packed->back();
                features_v1(p.board, p.to_move, p.symmetry, pos.planes, &pos.k_bits, legal);
                pos.reserved = 0;
                p.plan.build(p.board, p.to_move, opt_.search_kind, legal);
            }
            pt.lap(phase_clock().extract);
            ++emitted;
        }
        n_pending_ = emitted;
        if (emitted == 0) {
            if (hits_this_round > 0 && !is_done(*root_, opt_.num_rollout)) return emit(packed, raw);   // only cache hits: go on
            finish();
        }
        return emitted;
    }
    int emit(std::vector<dg_packed_position>& out) { return emit(&out, nullptr); }

    // Evaluations of the leaves of the last emit(), in the same order; `legal` ([n][361], identity orientation) is the
    // device's legal-move mask in raw mode, null otherwise.
    // `priors` ([n][368], optional): ready-to-insert priors of the leaves computed on the device (dg_engine_forward_raw_prior);
    // the root evaluation always goes through the host (8 symmetries averaged with weight 0.125 each).
    void absorb(const uint16_t* value, const uint16_t* policy /* [n][362] */, const uint8_t* legal = nullptr,
                const float* priors = nullptr) {
        evals_ += n_pending_;
        if (phase_ == ROOT) {
            if (legal) root_plan_.build(board_, color_, opt_.search_kind, legal);
            if (opt_.cache && !root_cached_)                 // lib.rs:124-131: every new response is offered to the table
                for (int t = 0; t < 8; ++t) opt_.cache->insert(board_.hash, color_, t, value[t], policy + (size_t)t * 362);
            float prior[368], acc[368];
            for (int i = 0; i < 368; ++i) acc[i] = NEG_INF;
            for (int p = 0; p <= N_POINTS; ++p) if (root_plan_.candidate[p]) acc[p] = 0.0f;
            float v = 0.0f;
            for (int t = 0; t < 8; ++t) {                     // lib.rs:112-127
                root_plan_.apply(policy + (size_t)t * 362, t, 0.125f, prior);
                v += (0.5f * f16_to_f32(value[t]) + 0.5f) * 0.125f;
                for (int i = 0; i < 362; ++i) acc[i] += prior[i];
            }
            memcpy(root_policy_, acc, sizeof(root_policy_));
            if (opt_.policy_only) {                          // self_play.rs:360-377
                std::vector<double> items(362);
                for (int i = 0; i < 362; ++i) items[i] = (double)acc[i];
                double at = opt_.choose_at >= 0.0 ? opt_.choose_at : rng_.uniform();
                int pick = choose(items, 0.5, 1.0 / (double)opt_.temperature, at);
                index_ = pick < 0 ? PASS : pick;
                value_ = v;
                phase_ = DONE;
                n_pending_ = 0;
                return;
            }
            if (!opt_.deterministic) {                       // lib.rs:163-165
                float eta[362];
                const float* noise = opt_.noise;
                if (!noise) { rng_.dirichlet(acc, opt_.dirichlet_shape, eta); noise = eta; }
                mix_noise(acc, noise, opt_.dirichlet_beta);
            }
            if (root_) root_->set_prior(acc);                // lib.rs:170-182
            else root_ = new Node(color_, v, acc);
            root_value_ = v;
            phase_ = PROBING;
            n_pending_ = 0;
            return;
        }
        float prior[368];
        for (int k = 0; k < n_pending_; ++k) {               // pool/worker_thread.rs:88-98
            Pending& p = pending_[k];
            float winrate = 0.5f * f16_to_f32(value[k]) + 0.5f;
            PhaseTimer pt;
            if (priors) {
                insert(p.trace, p.to_move, winrate, priors + (size_t)k * 368);
            } else {
                if (legal) p.plan.build(p.board, p.to_move, opt_.search_kind, legal + (size_t)k * N_POINTS);
                pt.lap(phase_clock().plan);
                p.plan.apply(policy + (size_t)k * 362, p.symmetry, 1.0f, prior);
                pt.lap(phase_clock().apply);
                insert(p.trace, p.to_move, winrate, prior);
            }
            pt.lap(phase_clock().insert);
            if (pt.on) phase_clock().leaves.fetch_add(1, std::memory_order_relaxed);
            if (opt_.cache) opt_.cache->insert(p.board.hash, p.to_move, p.symmetry, value[k], policy + (size_t)k * 362);   // worker_thread.rs:96
        }
        n_pending_ = 0;
    }

  private:
    struct Pending {
        Trace trace;
        Board board;
        PriorPlan plan;
        int to_move, symmetry;
    };

    int next_leaf_symmetry() {                               // pool/event.rs:47: one random symmetry per leaf
        if (opt_.leaf_symmetries && opt_.n_leaf_symmetries > 0) return opt_.leaf_symmetries[leaf_counter_++ % opt_.n_leaf_symmetries] & 7;
        ++leaf_counter_;
        return rng_.below(8);
    }

    void finish() {                                          // lib.rs:190-194
        float t = (!opt_.deterministic && board_.count < 8) ? opt_.temperature : 0.0f;
        double at = opt_.choose_at >= 0.0 ? opt_.choose_at : (t > 9e-2f ? rng_.uniform() : 0.0);
        index_ = best(*root_, t, at, &value_);
        phase_ = DONE;
    }

    Node* root_ = nullptr;
    Board board_;
    int color_ = BLACK;
    SearchOptions opt_;
    Rng rng_;
    Phase phase_ = DONE;
    PriorPlan root_plan_;
    std::vector<Pending> pending_;
    int n_pending_ = 0;
    long leaf_counter_ = 0, evals_ = 0, cache_inserts_ = 0;
    bool root_cached_ = false;
    float value_ = 0.5f, root_value_ = 0.5f;
    float root_policy_[362];
    int index_ = PASS;
};

}  // namespace dg
