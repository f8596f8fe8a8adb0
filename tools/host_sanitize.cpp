// AddressSanitizer / UBSan (and ThreadSanitizer: -fsanitize=thread) run of the host half (go_api.cpp + search_api.cpp) with the
// RandomPredictor: self-play with per-game transposition tables, with --ex-it, policy-only play, with ONE striped table shared
// by all games and threads, the queue-driven driver (dg_selfplay_run_engine) on a CPU stand-in for the leaf-batch queue, and a
// 600-ply game through the board API.
//   g++ -O1 -g -fsanitize=address,undefined -march=x86-64-v3 -ffp-contract=off -std=c++17 -Idream_go_b200/csrc \
//       tools/host_sanitize.cpp dream_go_b200/csrc/search_api.cpp dream_go_b200/csrc/go_api.cpp -o /tmp/host_sanitize -lpthread && /tmp/host_sanitize
#include <cstdio>
#include <vector>
#include "../include/dg_mcts.h"
extern "C" int32_t dg_engine_forward_packed(dg_engine*, const dg_packed_position*, int32_t, uint16_t*, uint16_t*) { return -1; }
extern "C" int32_t dg_engine_forward_raw(dg_engine*, const dg_raw_position*, int32_t, uint16_t*, uint16_t*, uint8_t*) { return -1; }
extern "C" int32_t dg_engine_forward_raw_prior(dg_engine*, const dg_raw_position*, int32_t, uint16_t*, uint16_t*, uint8_t*, float*) { return -1; }
// ---- a functional CPU stand-in for the engine's leaf-batch queue (include/dg_engine.h), so that the queue-driven driver
// (dg_selfplay_run_engine: polling scheduler, round-robin over engines, deadline handling) runs under the sanitizers without a
// device.  A "device" thread per batch evaluates the raw positions: legal moves from the stones and hashes alone (chains by
// flood fill, captures, super-ko through the zobrist table -- what csrc/features.cu does), value / policy = a hash of the stones.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include "../dream_go_b200/csrc/go_board.h"
#include "../dream_go_b200/csrc/search.h"
struct dg_engine { int max_batch = 256, workspaces = 4; std::atomic<int> taken{0}; };
struct dg_leaf_batch {
    dg_engine* e;
    std::vector<dg_raw_position> slots;
    std::vector<uint16_t> value, policy;
    std::vector<uint8_t> legal;
    std::vector<float> prior;
    std::atomic<int> fill{0}, committed{0}, ready{1};
    int submitted = 0;
    bool want_prior = false;
    std::thread worker;
    ~dg_leaf_batch() { if (worker.joinable()) worker.join(); }
};
static void mock_legal(const dg_raw_position& r, uint8_t* legal) {
    using namespace dg;
    const Tables& T = tables();
    auto bit = [](const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; };
    int col[361], lab[361], nlib[361];
    for (int p = 0; p < 361; ++p) { col[p] = bit(r.black, p) ? 1 : bit(r.white, p) ? 2 : 0; lab[p] = -1; }
    std::vector<std::vector<int>> chains;
    for (int p = 0; p < 361; ++p) {
        if (!col[p] || lab[p] >= 0) continue;
        std::vector<int> st{p}, ch;
        lab[p] = (int)chains.size();
        while (!st.empty()) {
            int s = st.back(); st.pop_back(); ch.push_back(s);
            for (int k = 0; k < T.n_nbr[s]; ++k) { int q = T.nbr_list[s][k]; if (col[q] == col[p] && lab[q] < 0) { lab[q] = lab[p]; st.push_back(q); } }
        }
        bool seen[361] = {false}; int n = 0;
        for (int s : ch) for (int k = 0; k < T.n_nbr[s]; ++k) { int q = T.nbr_list[s][k]; if (!col[q] && !seen[q]) { seen[q] = true; ++n; } }
        nlib[chains.size()] = n;
        chains.push_back(ch);
    }
    const int c = r.to_move, opp = 3 - c;
    for (int p = 0; p < 361; ++p) {
        legal[p] = 0;
        if (col[p]) continue;
        bool ok = false;
        uint64_t h = r.hash ^ T.zobrist[c][p];
        int cap[4], nc = 0;
        for (int k = 0; k < T.n_nbr[p]; ++k) {
            int q = T.nbr_list[p][k];
            if (!col[q]) { ok = true; continue; }
            const int n = nlib[lab[q]];
            if ((col[q] == c) == (n >= 2)) ok = true;
            if (col[q] == opp && n < 2) {
                bool dup = false;
                for (int j = 0; j < nc; ++j) dup |= cap[j] == lab[q];
                if (!dup) { cap[nc++] = lab[q]; for (int s : chains[lab[q]]) h ^= T.zobrist[opp][s]; }
            }
        }
        if (!ok) continue;
        bool ko = false;
        if (bit(r.visited, p)) for (int i = 0; i < 16; ++i) ko |= r.hash_history[i] == h;
        legal[p] = !ko;
    }
}
static void mock_evaluate(dg_leaf_batch* b) {
    for (int i = 0; i < b->submitted; ++i) {
        const dg_raw_position& r = b->slots[i];
        mock_legal(r, &b->legal[(size_t)i * 361]);
        uint64_t h = 0xcbf29ce484222325ull ^ r.to_move ^ ((uint64_t)(r.symmetry & 7) << 8);
        for (int w = 0; w < 12; ++w) { h ^= r.black[w]; h *= 0x100000001b3ull; h ^= r.white[w]; h *= 0x100000001b3ull; }
        dg::Rng rng(h);
        b->value[i] = dg::f32_to_f16_bits((float)(2.0 * rng.uniform() - 1.0));
        float x[362], total = 0.f;
        for (int k = 0; k < 362; ++k) { x[k] = (float)rng.uniform(); total += x[k]; }
        for (int k = 0; k < 362; ++k) b->policy[(size_t)i * 362 + k] = dg::f32_to_f16_bits(x[k] / total);
        if (b->want_prior) {
            // the priors the device would return (dg_engine_forward_raw_prior): the stones put back on a board (black first, then
            // white: no chain of a legal position ever runs out of liberties on the way), then the host's own prior construction
            dg_board* board = dg_board_new(7.5f);
            for (int p = 0; p < 361; ++p) if ((r.black[p >> 5] >> (p & 31)) & 1u) dg_board_place(board, dg::BLACK, p);
            for (int p = 0; p < 361; ++p) if ((r.white[p >> 5] >> (p & 31)) & 1u) dg_board_place(board, dg::WHITE, p);
            dg_board_prior(board, r.to_move, r.symmetry >> 4, &b->legal[(size_t)i * 361], &b->policy[(size_t)i * 362], r.symmetry & 7, 1.0f,
                           &b->prior[(size_t)i * 368]);
            dg_board_free(board);
        }
    }
    b->ready.store(1, std::memory_order_release);
}
extern "C" int32_t dg_engine_max_batch(dg_engine* e) { return e ? e->max_batch : 0; }
extern "C" int32_t dg_engine_num_workspaces(dg_engine* e) { return e ? e->workspaces : 0; }
extern "C" int32_t dg_engine_batch_acquire(dg_engine* e, dg_leaf_batch** out) {
    if (!e || !out) return -5;
    dg_leaf_batch* b = new dg_leaf_batch();
    b->e = e;
    b->slots.resize(e->max_batch); b->value.resize(e->max_batch); b->policy.resize((size_t)e->max_batch * 362); b->legal.resize((size_t)e->max_batch * 361);
    b->prior.resize((size_t)e->max_batch * 368);
    e->taken++;
    *out = b;
    return 0;
}
extern "C" void dg_engine_batch_release(dg_leaf_batch* b) { if (b) { b->e->taken--; delete b; } }
extern "C" int32_t dg_leaf_batch_push(dg_leaf_batch* b, const dg_raw_position* pos, int32_t n) {
    int at = b->fill.load();
    do { if (at < 0 || at + n > b->e->max_batch) return -1; } while (!b->fill.compare_exchange_weak(at, at + n));
    memcpy(&b->slots[at], pos, sizeof(dg_raw_position) * (size_t)n);
    b->committed.fetch_add(n, std::memory_order_release);
    return at;
}
extern "C" int32_t dg_leaf_batch_submit(dg_leaf_batch* b, uint32_t outputs) {
    const int n = b->fill.exchange(INT32_MIN);
    if (n <= 0) return -5;
    while (b->committed.load(std::memory_order_acquire) < n) {}
    if (b->worker.joinable()) b->worker.join();
    b->submitted = n;
    b->want_prior = (outputs & DG_LEAF_PRIOR) != 0;
    b->ready.store(0, std::memory_order_release);
    b->worker = std::thread(mock_evaluate, b);
    return 0;
}
extern "C" int32_t dg_leaf_batch_ready(dg_leaf_batch* b) { return b->ready.load(std::memory_order_acquire); }
extern "C" void dg_leaf_batch_reset(dg_leaf_batch* b) { b->committed.store(0); b->fill.store(0, std::memory_order_release); }
extern "C" const uint16_t* dg_leaf_batch_value(const dg_leaf_batch* b) { return b->value.data(); }
extern "C" const uint16_t* dg_leaf_batch_policy(const dg_leaf_batch* b) { return b->policy.data(); }
extern "C" const uint8_t* dg_leaf_batch_legal(const dg_leaf_batch* b) { return b->legal.data(); }
extern "C" const float* dg_leaf_batch_prior(const dg_leaf_batch* b) { return b->prior.data(); }
// `host_sanitize bench [seconds] [games] [threads] [flags]`: the queue-driven driver on the stand-in for a fixed time, with
// DG_SELFPLAY_TRACE=1 the worker threads' cycles per leaf by phase (games run from the opening into the middle game)
static int bench(int argc, char** argv) {
    dg_engine e; dg_engine* one[1] = {&e};
    dg_selfplay_config c{}; c.num_games = 100000; c.num_rollout = 800; c.probes_per_round = 8; c.seed = 20261017; c.dirichlet_noise = 0.25f; c.temperature = 0.8f;
    c.max_seconds = argc > 2 ? atof(argv[2]) : 10.0; c.num_parallel = argc > 3 ? atoi(argv[3]) : 16; c.num_threads = argc > 4 ? atoi(argv[4]) : 1;
    const uint32_t flags = argc > 5 ? (uint32_t)atoi(argv[5]) : 0u;
    dg_selfplay_stats s{};
    int rc = dg_selfplay_run_engine(one, 1, flags, &c, &s, nullptr, 0);
    printf("bench rc %d games %ld moves %ld evals %ld seconds %.2f evals/s %.0f moves/s %.1f\n", rc, (long)s.games_finished, (long)s.moves, (long)s.evals, s.seconds,
           s.evals / s.seconds, s.moves / s.seconds);
    return rc;
}
// `host_sanitize fuzz [cases] [seed]`: random configurations of the queue-driven driver on one or two stand-in engines -- every
// run ends, releases its leaf batches, and plays the games of the plainest schedule (one engine, one worker, host priors)
static int fuzz(int argc, char** argv) {
    const int cases = argc > 2 ? atoi(argv[2]) : 40;
    dg::Rng rng(argc > 3 ? (uint64_t)atoll(argv[3]) : 1);
    auto pick = [&](std::initializer_list<int> xs) { return *(xs.begin() + rng.below((int)xs.size())); };
    int bad = 0, refused = 0;
    for (int k = 0; k < cases; ++k) {
        dg_selfplay_config c{};
        c.num_games = pick({1, 2, 5, 9}); c.num_parallel = pick({1, 2, 5, 12, 40}); c.num_rollout = pick({1, 2, 20, 70}); c.probes_per_round = pick({1, 2, 4, 8, 12});
        c.max_plies = pick({1, 6, 20, 50}); c.seed = 1 + (uint64_t)rng.below(1 << 30); c.dirichlet_noise = 0.25f; c.temperature = 0.8f;
        c.ex_it = rng.below(4) == 0; c.num_ex_it_rollout = pick({5, 40}); c.cache_capacity = pick({0, 0, 9, 400});
        dg_selfplay_stats want{}, got{};
        dg_engine r0; dg_engine* plain[1] = {&r0};
        dg_selfplay_config base = c; base.num_threads = 1; base.num_groups = 0;
        const int rc0 = dg_selfplay_run_engine(plain, 1, 0u, &base, &want, nullptr, 0);
        dg_engine e0, e1; dg_engine* two[2] = {&e0, &e1};
        c.num_threads = pick({1, 2, 3, 6}); c.num_groups = pick({0, 1, 2, 3, 4});
        const int n_engines = 1 + rng.below(2);
        const uint32_t flags = (uint32_t)pick({0, 1, 4, 5});
        const int rc = dg_selfplay_run_engine(two, n_engines, flags, &c, &got, nullptr, 0);
        const bool same = rc == 0 && rc0 == 0 && got.digest == want.digest && got.moves == want.moves && got.games_finished == c.num_games;
        const bool both_refused = rc == -5 || rc0 == -5;          // games per group x leaves per round beyond one leaf batch: DG_ERR_INVALID_ARGUMENT
        refused += both_refused;
        const int held = r0.taken.load() + e0.taken.load() + e1.taken.load();
        if ((!same && !both_refused) || held != 0) {
            ++bad;
            printf("fuzz case %d FAILED: rc %d / %d games %ld of %d moves %ld / %ld digest %llx / %llx held %d (parallel %d rollout %d probes %d plies %d threads %d groups %d engines %d flags %u ex_it %d cache %d seed %llu)\n",
                   k, rc, rc0, (long)got.games_finished, c.num_games, (long)got.moves, (long)want.moves, (unsigned long long)got.digest, (unsigned long long)want.digest, held,
                   c.num_parallel, c.num_rollout, c.probes_per_round, c.max_plies, c.num_threads, c.num_groups, n_engines, flags, c.ex_it, c.cache_capacity, (unsigned long long)c.seed);
        }
    }
    printf("fuzz: %d cases, %d refused (do not fit a leaf batch), %d failures\n", cases, refused, bad);
    return bad ? 1 : 0;
}
int main(int argc, char** argv){
  if (argc > 1 && !strcmp(argv[1], "bench")) return bench(argc, argv);
  if (argc > 1 && !strcmp(argv[1], "fuzz")) return fuzz(argc, argv);
  for (int variant = 0; variant < 4; ++variant) {
    dg_selfplay_config c{}; c.num_games=5; c.num_parallel=3; c.num_rollout= variant==2 ? 1 : 60; c.probes_per_round=4; c.max_plies=30; c.num_threads=3; c.dirichlet_noise=0.25f; c.temperature=0.8f; c.seed=3+variant;
    c.ex_it = variant==1; c.num_ex_it_rollout=80; c.cache_capacity = variant==0 ? 64 : variant==3 ? 20000 : 0; c.cache_shared = variant==3 ? 8 : 0; c.num_groups = variant % 3 + 1;
    dg_selfplay_stats s{};
    std::vector<char> sgf(1<<20);
    int rc=dg_selfplay_run(dg_random_predict,nullptr,&c,&s,sgf.data(),sgf.size());
    printf("variant %d rc %d games %ld moves %ld evals %ld hits %ld\n",variant,rc,(long)s.games_finished,(long)s.moves,(long)s.evals,(long)s.cache_hits);
  }
  // the queue-driven driver on two stand-in engines: same games whatever the number of workers, groups and engines
  {
    uint64_t digests[6];
    for (int variant = 0; variant < 6; ++variant) {
      dg_engine e0, e1;
      dg_engine* engines[2] = {&e0, &e1};
      dg_selfplay_config c{}; c.num_games=7; c.num_parallel=5; c.num_rollout=50; c.probes_per_round=4; c.max_plies=24; c.dirichlet_noise=0.25f; c.temperature=0.8f; c.seed=11;
      c.num_threads = variant == 0 ? 1 : 4; c.num_groups = variant == 2 ? 1 : 2;
      dg_selfplay_stats s{};
      std::vector<char> sgf(1<<20);
      // variants 4, 5: the priors from the "device" (always / placed by the driver from the workers' load)
      const uint32_t flags = variant == 4 ? DG_SELFPLAY_DEVICE_PRIORS : variant == 5 ? DG_SELFPLAY_AUTO_PRIORS | DG_SELFPLAY_DEVICE_PRIORS : 0u;
      int rc = dg_selfplay_run_engine(engines, variant == 3 ? 1 : 2, flags, &c, &s, sgf.data(), sgf.size());
      digests[variant] = s.digest;
      printf("queue variant %d rc %d games %ld moves %ld evals %ld batches %ld leaf batches still held %d\n", variant, rc, (long)s.games_finished,
             (long)s.moves, (long)s.evals, (long)s.rounds, e0.taken.load() + e1.taken.load());
      if (rc || s.games_finished != 7 || digests[variant] != digests[0]) { printf("FAILED\n"); return 1; }
    }
    // the process-wide transposition table under the queue-driven driver (the parameters of
    // tests/test_selfplay_gpu.py::test_self_play_on_engine_with_one_shared_table): every game ends, later games reuse earlier evaluations
    {
      long evals[2], moves[2], hits[2];
      for (int shared = 0; shared < 2; ++shared) {
        dg_engine e; dg_engine* one[1] = {&e};
        dg_selfplay_config c{}; c.num_games=8; c.num_parallel=2; c.num_rollout=24; c.probes_per_round=2; c.max_plies=8; c.num_threads=2; c.seed=3;
        c.dirichlet_noise=0.25f; c.temperature=0.8f; c.cache_capacity = shared ? 200000 : 0; c.cache_shared = shared ? 16 : 0;
        dg_selfplay_stats s{};
        std::vector<char> sgf(1<<20);
        int rc = dg_selfplay_run_engine(one, 1, 0u, &c, &s, sgf.data(), sgf.size());
        evals[shared] = (long)s.evals; moves[shared] = (long)s.moves; hits[shared] = (long)s.cache_hits;
        printf("queue shared-table %d rc %d games %ld moves %ld evals %ld hits %ld held %d\n", shared, rc, (long)s.games_finished, moves[shared], evals[shared],
               hits[shared], e.taken.load());
        if (rc || s.games_finished != 8 || e.taken.load() != 0) { printf("FAILED\n"); return 1; }
      }
      if (moves[0] != 64 || moves[1] != 64 || hits[1] <= 0 || evals[1] >= evals[0]) { printf("FAILED\n"); return 1; }
    }
    dg_engine e0; dg_engine* engines[1] = {&e0};                   // a deadline in the middle of the run
    dg_selfplay_config c{}; c.num_games=100000; c.num_parallel=16; c.num_rollout=200; c.probes_per_round=8; c.num_threads=4; c.seed=5; c.max_seconds=0.5;
    c.dirichlet_noise=0.25f; c.temperature=0.8f;
    dg_selfplay_stats s{};
    int rc = dg_selfplay_run_engine(engines, 1, 0u, &c, &s, nullptr, 0);
    printf("queue deadline run rc %d evals %ld seconds %.2f held %d\n", rc, (long)s.evals, s.seconds, e0.taken.load());
    if (rc || s.evals <= 0 || e0.taken.load() != 0) { printf("FAILED\n"); return 1; }
  }
  // board API: a game with captures, features, priors
  dg_board* b = dg_board_new(7.5f);
  int color = 1; unsigned long long rng = 12345;
  for (int ply = 0; ply < 600; ++ply) {
    uint8_t legal[361]; dg_packed_position pos; dg_raw_position raw;
    dg_board_features_packed(b, color, ply % 8, &pos, legal);
    dg_board_raw_position(b, color, ply % 8, &raw);
    { uint8_t again[361]; mock_legal(raw, again); if (memcmp(again, legal, 361)) { printf("stand-in legal mask differs at ply %d\n", ply); return 1; } }
    std::vector<int> cand; for (int p = 0; p < 361; ++p) if (legal[p]) cand.push_back(p);
    if (cand.empty()) break;
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    uint16_t policy[362]; for (int i = 0; i < 362; ++i) policy[i] = 0x1c00; float prior[368];
    dg_board_prior(b, color, ply & 1, legal, policy, ply % 8, 1.0f, prior);
    uint8_t terr[361]; dg_board_territory(b, terr); dg_board_is_scorable(b);
    dg_board_place(b, color, cand[(rng >> 33) % cand.size()]); color = 3 - color;
  }
  dg_board_free(b);
  printf("board api ok\n");
}
