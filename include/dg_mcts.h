/*
 * dg_mcts.h -- C ABI of the tree search and the self-play driver (host side of the self-play hot path).
 *
 * Replaces, for this path, the Rust crate API of `dg_mcts`:
 *   `predict(pool, options, time_strategy, starting_tree, board, color)`     src/libdg_mcts/lib.rs:145-200
 *   `tree::Node::{forward, disqualify, best, softmax}`                        src/libdg_mcts/tree.rs:1198-1301
 *   `trait Predictor::predict(features, batch)`                               src/libdg_mcts/predictor.rs:59-92
 *   `self_play(network, num_games, ex_it)`                                    src/libdg_mcts/self_play.rs:423-591
 * Exported by the same `libdg_engine.so` as dg_engine.h / dg_go.h.
 */
#ifndef DG_MCTS_H
#define DG_MCTS_H

#include <stdint.h>
#include "dg_engine.h"
#include "dg_go.h"

#ifdef __cplusplus
extern "C" {
#endif

/* `Predictor::predict` (predictor.rs:91) over the compact position format: evaluates `n` positions, writes n fp16
 * values (post-tanh) and n*362 fp16 policies (post-softmax).  Returns 0 on success.  Called from one thread at a
 * time per search / per self-play driver. */
typedef int32_t (*dg_predict_fn)(void* ctx, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy);

/* The predictor that is the product: ctx = dg_engine*, evaluation through dg_engine_forward_packed. */
int32_t dg_engine_predict(void* engine, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy);

/* The same over raw positions: the predictor derives the feature planes itself and also returns the legal-move mask
 * of every position ([n][361], `Board::is_valid(to_move, .)`), which the search needs for the prior. */
typedef int32_t (*dg_predict_raw_fn)(void* ctx, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy,
                                     uint8_t* legal);
/* ctx = dg_engine*, evaluation through dg_engine_forward_raw (feature planes and legal moves computed on the device). */
int32_t dg_engine_predict_raw(void* engine, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy,
                              uint8_t* legal);

/* ... and with the leaves' ready-to-insert priors ([n][368] floats; dg_engine_forward_raw_prior): the host then only walks
 * the tree.  policy and legal are still returned (the root evaluation and the transposition table use them). */
typedef int32_t (*dg_predict_prior_fn)(void* ctx, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy,
                                       uint8_t* legal, float* prior);
int32_t dg_engine_predict_prior(void* engine, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy,
                                uint8_t* legal, float* prior);

/* `RandomPredictor` (predictors/random.rs:30-59) as a deterministic function of the position; ctx = NULL or a
 * uint64_t* salt.  No device involved: it measures the host half of self-play alone. */
int32_t dg_random_predict(void* ctx, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy);
/* The same with a PEAKED policy (softmax(sharpness * u), ctx = float* sharpness or NULL for 12): a stand-in for a trained
 * network's narrow, deep searches when measuring the host side (transposition-table hit rates, tree depth). */
int32_t dg_peaked_predict(void* ctx, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy);

struct dg_cache;
typedef struct dg_search_options {
    int32_t  search;              /* DG_STANDARD_SEARCH / DG_SCORING_SEARCH (options.rs:66-81, 141-162)                 */
    int32_t  deterministic;       /* SearchOptions::deterministic(): no Dirichlet noise, greedy move choice             */
    int32_t  num_rollout;         /* RolloutLimit::new(n) (time_control/rollout_limit.rs)                               */
    int32_t  probes_per_round;    /* leaves of ONE tree evaluated together (1 = the sequential algorithm)               */
    float    dirichlet_noise;     /* DIRICHLET_NOISE, 0.25 in self-play (config.rs:165-166)                             */
    float    temperature;         /* TEMPERATURE, 0.8 in self-play, used for the first 8 plies (lib.rs:190-194)         */
    uint64_t seed;                /* all randomness of the search derives from it (the reference uses thread_rng)       */
    const float*   noise;         /* optional: the normalised Dirichlet sample eta[362] to mix in (tests)              */
    const uint8_t* leaf_symmetries; /* optional: symmetry of the k-th leaf = leaf_symmetries[k % n] (tests)            */
    int32_t  n_leaf_symmetries;
    double   choose_at;           /* optional: the uniform number of the stochastic move choice; < 0 = draw it          */
    struct dg_cache* cache;       /* optional: transposition table of evaluations (NnPredictor::fetch / cache)          */
    int32_t  device_ladders;      /* raw-position predictors: leave the ladder planes to the device (DG_RAW_DEVICE_LADDERS) */
} dg_search_options;

/* Transposition table `LruCache<(zobrist hash, to_move), Prediction>` (src/libdg_mcts/predictors/nn.rs:29-82,
 * lru_cache.rs): entries are kept in identity orientation and answer every symmetry (predictor.rs:30-44). */
typedef struct dg_cache dg_cache;
dg_cache* dg_cache_new(int32_t capacity);                 /* one LRU list behind one lock: the reference's table (200,000 entries) */
/* The same table split into `stripes` LRU lists with a lock each (<= 0: 64): many searches on many threads share it without
 * queueing on one mutex; eviction is least-recent within a stripe.  Every dg_cache is safe to use from several threads. */
dg_cache* dg_cache_new_shared(int32_t capacity, int32_t stripes);
void      dg_cache_free(dg_cache* cache);
void      dg_cache_stats(const dg_cache* cache, int64_t* hits, int64_t* misses, int64_t* size);

typedef struct dg_tree dg_tree;   /* `tree::Node` */

/* `dg_mcts::predict`: searches `board` for `color`.  `starting_tree` (may be NULL) is consumed.  Outputs: the value
 * and index (0..361, 361 = pass) of the chosen move, and the searched tree (caller frees or forwards it).
 * evals_out (optional) = positions evaluated.  Returns 0, the predictor's error, or DG_ERR_INVALID_ARGUMENT for a null
 * predictor / options / board, a colour other than 1 / 2 or an unknown search kind (the starting tree is consumed either way). */
int32_t  dg_mcts_predict(dg_predict_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                         const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                         int64_t* evals_out);
/* The same search with a raw-position predictor: identical trees, the host skips the feature planes. */
int32_t  dg_mcts_predict_raw(dg_predict_raw_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                             const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                             int64_t* evals_out);
int32_t  dg_mcts_predict_prior(dg_predict_prior_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                               const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                               int64_t* evals_out);
void     dg_tree_free(dg_tree* tree);
dg_tree* dg_tree_forward(dg_tree* tree, int32_t index);            /* Node::forward (tree.rs:1198-1225); consumes `tree`; NULL when there
                                                                      is no sub-tree (or index is outside 0..361) */
void     dg_tree_disqualify(dg_tree* tree, int32_t index);         /* Node::disqualify (tree.rs:1296-1301); indices outside 0..361 are ignored */
int32_t  dg_tree_total_count(const dg_tree* tree);
int32_t  dg_tree_to_move(const dg_tree* tree);
float    dg_tree_initial_value(const dg_tree* tree);
/* per child 0..361: visit count, mean value (initial value if unvisited), prior (-inf = not a candidate) */
void     dg_tree_children(const dg_tree* tree, int32_t* count, float* value, float* prior);
int64_t  dg_tree_num_nodes(const dg_tree* tree);

/* ---- self-play (self_play.rs:423-591) ------------------------------------------------------------------------------ */
typedef struct dg_selfplay_config {
    int32_t  num_games;           /* games to play in total (`--self-play N`)                                           */
    int32_t  num_parallel;        /* games in flight (`--num-games`, config.rs NUM_GAMES)                               */
    int32_t  num_rollout;         /* `--num-rollout` (config.rs NUM_ROLLOUT); per move 800*max(0.1, 4w(1-w)), :234-241   */
    int32_t  probes_per_round;    /* leaves per tree per device batch                                                   */
    int32_t  max_plies;           /* 722 in the reference (self_play.rs:437); smaller values bound benchmark runs       */
    int32_t  num_threads;         /* host threads for probing / feature extraction (<= 0: all cores)                    */
    int32_t  ex_it;               /* `--ex-it`: 5 % of the eligible positions get a full search (self_play.rs:287-319)  */
    int32_t  num_ex_it_rollout;   /* `--num-ex-it-rollout`                                                              */
    float    dirichlet_noise, temperature;
    uint64_t seed;
    double   max_seconds;         /* stop starting new rounds after this much wall time (<= 0: no limit)                */
    int32_t  cache_capacity;      /* entries of each game's transposition table (0 = none); see dg_cache                */
    int32_t  num_groups;          /* groups of games, each either on the host or on the device (1..8); 0 = chosen from num_parallel * probes_per_round */
    int32_t  cache_shared;        /* 0: one table of cache_capacity entries per game (games stay a function of the seed);
                                     n > 0: ONE table of cache_capacity entries in n lock stripes shared by all games, as the
                                     reference's process-wide LRU (predictors/nn.rs:48-50) -- games then depend on each other's timing */
} dg_selfplay_config;

typedef struct dg_selfplay_stats {
    int64_t games_finished, moves, evals, rounds, searches;
    double  seconds, eval_seconds;        /* wall time total / spent inside the predictor                              */
    double  mean_batch;                   /* positions per predictor call                                              */
    uint64_t digest;                      /* order-independent hash of every finished game's move list (determinism)   */
    int64_t cache_hits;                   /* leaf / root evaluations answered by the transposition tables              */
} dg_selfplay_stats;

/* Plays `num_games` games, `num_parallel` at a time, all sharing one predictor; one SGF record per finished game is
 * appended to `sgf_out` (NUL-terminated, truncated at sgf_capacity; may be NULL).  Returns 0 or the predictor's error. */
int32_t  dg_selfplay_run(dg_predict_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                         char* sgf_out, int64_t sgf_capacity);

/* The same with a raw-position predictor (dg_engine_predict_raw): same games for the same seed. */
int32_t  dg_selfplay_run_raw(dg_predict_raw_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                             char* sgf_out, int64_t sgf_capacity);

int32_t  dg_selfplay_run_prior(dg_predict_prior_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                               char* sgf_out, int64_t sgf_capacity);

/* The product path: self-play on `n_engines` engines of ONE process (one per device, as `Device::all()` + the round-robin
 * of predictors/nn.rs:84-92), the host threads shared by all of them.  Every group of games owns a leaf batch of its
 * engine (dg_engine.h): the worker that advances the group's last game pushes the leaves and submits them (one graph
 * launch), any worker notices the completion flag and goes on with the group's games -- no thread blocks in a predictor
 * call (pool/worker_thread.rs:88-99 does).  Each engine needs num_workspaces >= its number of groups (config->num_groups
 * per engine, 2..8) and max_batch >= games per group x max(8, probes_per_round).  Same games as the three calls above for the
 * same seed.  flags: DG_SELFPLAY_DEVICE_PRIORS = the leaves' priors are built on the device (dg_engine_forward_raw_prior). */
#define DG_SELFPLAY_DEVICE_PRIORS  0x1u
#define DG_SELFPLAY_DEVICE_LADDERS 0x2u   /* the ladder planes are read on the device too: the host only walks the trees */
#define DG_SELFPLAY_AUTO_PRIORS    0x4u   /* the driver decides batch by batch from how busy its worker threads are (windows of 64
                                             batches): above 85 % the host is the limit and the priors move to the device, below 60 % back
                                             to the host; DG_SELFPLAY_DEVICE_PRIORS is then only the start value.  Same games either way
                                             (the priors are bit-identical). */
int32_t  dg_selfplay_run_engine(dg_engine* const* engines, int32_t n_engines, uint32_t flags, const dg_selfplay_config* config,
                                dg_selfplay_stats* stats, char* sgf_out, int64_t sgf_capacity);

#ifdef __cplusplus
}
#endif
#endif /* DG_MCTS_H */
