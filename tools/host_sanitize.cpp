// AddressSanitizer / UBSan (and ThreadSanitizer: -fsanitize=thread) run of the host half (go_api.cpp + search_api.cpp) with the
// RandomPredictor: self-play with per-game transposition tables, with --ex-it, policy-only play, with ONE striped table shared
// by all games and threads, and a 600-ply game through the board API.
//   g++ -O1 -g -fsanitize=address,undefined -march=x86-64-v3 -ffp-contract=off -std=c++17 -Idream_go_b200/csrc \
//       tools/host_sanitize.cpp dream_go_b200/csrc/search_api.cpp dream_go_b200/csrc/go_api.cpp -o /tmp/host_sanitize -lpthread && /tmp/host_sanitize
#include <cstdio>
#include <vector>
#include "../include/dg_mcts.h"
extern "C" int32_t dg_engine_forward_packed(dg_engine*, const dg_packed_position*, int32_t, uint16_t*, uint16_t*) { return -1; }
extern "C" int32_t dg_engine_forward_raw(dg_engine*, const dg_raw_position*, int32_t, uint16_t*, uint16_t*, uint8_t*) { return -1; }
extern "C" int32_t dg_engine_forward_raw_prior(dg_engine*, const dg_raw_position*, int32_t, uint16_t*, uint16_t*, uint8_t*, float*) { return -1; }
extern "C" int32_t dg_engine_max_batch(dg_engine*) { return 0; }
extern "C" int32_t dg_engine_num_workspaces(dg_engine*) { return 0; }
extern "C" int32_t dg_engine_batch_acquire(dg_engine*, dg_leaf_batch**) { return -1; }
extern "C" void dg_engine_batch_release(dg_leaf_batch*) {}
extern "C" int32_t dg_leaf_batch_push(dg_leaf_batch*, const dg_raw_position*, int32_t) { return -1; }
extern "C" int32_t dg_leaf_batch_submit(dg_leaf_batch*, uint32_t) { return -1; }
extern "C" int32_t dg_leaf_batch_ready(dg_leaf_batch*) { return -1; }
extern "C" void dg_leaf_batch_reset(dg_leaf_batch*) {}
extern "C" const uint16_t* dg_leaf_batch_value(const dg_leaf_batch*) { return nullptr; }
extern "C" const uint16_t* dg_leaf_batch_policy(const dg_leaf_batch*) { return nullptr; }
extern "C" const uint8_t* dg_leaf_batch_legal(const dg_leaf_batch*) { return nullptr; }
extern "C" const float* dg_leaf_batch_prior(const dg_leaf_batch*) { return nullptr; }
int main(){
  for (int variant = 0; variant < 4; ++variant) {
    dg_selfplay_config c{}; c.num_games=5; c.num_parallel=3; c.num_rollout= variant==2 ? 1 : 60; c.probes_per_round=4; c.max_plies=30; c.num_threads=3; c.dirichlet_noise=0.25f; c.temperature=0.8f; c.seed=3+variant;
    c.ex_it = variant==1; c.num_ex_it_rollout=80; c.cache_capacity = variant==0 ? 64 : variant==3 ? 20000 : 0; c.cache_shared = variant==3 ? 8 : 0; c.num_groups = variant % 3 + 1;
    dg_selfplay_stats s{};
    std::vector<char> sgf(1<<20);
    int rc=dg_selfplay_run(dg_random_predict,nullptr,&c,&s,sgf.data(),sgf.size());
    printf("variant %d rc %d games %ld moves %ld evals %ld hits %ld\n",variant,rc,(long)s.games_finished,(long)s.moves,(long)s.evals,(long)s.cache_hits);
  }
  // board API: a game with captures, features, priors
  dg_board* b = dg_board_new(7.5f);
  int color = 1; unsigned long long rng = 12345;
  for (int ply = 0; ply < 600; ++ply) {
    uint8_t legal[361]; dg_packed_position pos; dg_raw_position raw;
    dg_board_features_packed(b, color, ply % 8, &pos, legal);
    dg_board_raw_position(b, color, ply % 8, &raw);
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
