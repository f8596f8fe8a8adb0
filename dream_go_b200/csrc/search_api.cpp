// search_api.cpp -- C ABI of include/dg_mcts.h: blocking single search (dg_mcts_predict) and the self-play driver.
//
// The driver is the engine-side replacement of `self_play` + `Pool` + `Batcher` (src/libdg_mcts/self_play.rs:423-500,
// pool/pool.rs, pool/batch.rs): the games in flight are split into a few groups.  A group is either on the host -- any free
// worker thread advances the next of its games (insert the evaluations, probe the tree, extract the next leaves), and the
// worker that finishes the last one gathers the leaves and submits them -- or on the device, where the group's device
// thread sits in the blocking predictor call.  Every game owns its random stream, so the games played are a function of
// the seed alone, whatever the number of threads and groups.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include <sys/prctl.h>
#include <time.h>

#include "../../include/dg_mcts.h"
#include "search_task.h"

using namespace dg;

static inline Node* N(dg_tree* t) { return reinterpret_cast<Node*>(t); }
static inline const Node* N(const dg_tree* t) { return reinterpret_cast<const Node*>(t); }

// What the reference's types make unrepresentable (`Color`, the `SearchOptions` implementations): a caller of the C ABI can
// hand over any integer.
static bool valid_search_call(const dg_search_options* o, const dg_board* board, int32_t color) {
    return o && board && (color == BLACK || color == WHITE) && (o->search == STANDARD_SEARCH || o->search == SCORING_SEARCH) &&
           (o->n_leaf_symmetries <= 0 || o->leaf_symmetries != nullptr);
}

static SearchOptions convert(const dg_search_options* o) {
    SearchOptions s;
    s.search_kind = o->search;
    s.deterministic = o->deterministic != 0;
    s.num_rollout = o->num_rollout;
    s.probes_per_round = o->probes_per_round > 0 ? o->probes_per_round : 1;
    s.dirichlet_beta = o->dirichlet_noise;
    s.temperature = o->temperature;
    s.noise = o->noise;
    s.leaf_symmetries = o->leaf_symmetries;
    s.n_leaf_symmetries = o->n_leaf_symmetries;
    s.choose_at = o->choose_at;
    s.cache = reinterpret_cast<PredictionCache*>(o->cache);
    s.device_ladders = o->device_ladders != 0;
    return s;
}

static int64_t count_nodes(const Node* n) {
    int64_t c = 1;
    for (int k = 0; k < n->n_edges; ++k) if (n->e_child[k]) c += count_nodes(n->e_child[k]);
    return c;
}

// ---- SGF record (self_play.rs:187-214, game_result.rs:23-43) ---------------------------------------------------------
static const char B85[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!#$%&()*+-;<=>?@^_`{|}~";

static void b85_encode_f16(const float* x, int n, std::string& out) {      // libdg_utils/b85.rs:141-165
    for (int i = 0; i + 1 < n; i += 2) {
        uint16_t a = f32_to_f16_bits(x[i]), b = f32_to_f16_bits(x[i + 1]);
        a = (uint16_t)((a << 8) | (a >> 8));
        b = (uint16_t)((b << 8) | (b >> 8));
        uint64_t acc = ((uint64_t)a << 16) | b;
        out.push_back(B85[acc / 52200625]);
        out.push_back(B85[(acc / 614125) % 85]);
        out.push_back(B85[(acc / 7225) % 85]);
        out.push_back(B85[(acc / 85) % 85]);
        out.push_back(B85[acc % 85]);
    }
}

static void sgf_point(int index, std::string& out) {                        // utils/sgf.rs:36-43 (CGoban)
    if (index >= N_POINTS) return;
    out.push_back((char)('a' + index % 19));
    out.push_back((char)('a' + index / 19));
}

namespace {

struct Player {                                                  // self_play.rs:217-241
    float winrate = 0.5f;
    Node* root = nullptr;
    int color = BLACK;
    int num_rollout(int max_rollout) const {
        float m = 4.0f * winrate * (1.0f - winrate);
        m = m < 0.1f ? 0.1f : m;
        return (int)(m * (float)max_rollout);
    }
    void update(float value) { winrate -= 0.2f * (winrate - value); }        // MovingAverage, MOMENTUM = 0.2
};

struct Game {
    enum Mode { IDLE, SEARCH, EX_IT };
    int64_t id = -1;
    Board board;
    Player players[2];                                           // players[0] is to move (self_play.rs:433-457)
    int pass_count = 0;
    bool active = false;
    Mode mode = IDLE;
    bool allow_pass = false;
    Rng rng;
    SearchTask task;
    std::unique_ptr<PredictionCache> cache;                      // this game's own table (cache_shared == 0)
    long cache_hits_before = 0;
    std::string sgf;
    std::vector<uint16_t> moves;
    std::vector<dg_packed_position> batch;                       // this round's leaves (host feature planes) ...
    std::vector<dg_raw_position> raw_batch;                      // ... or raw positions (planes derived on the device)
    int n_emitted = 0;
    size_t batch_offset = 0;
    // result of the main search, kept while an ex-it search replaces the recorded statistics
    float value = 0.5f;
    int index = PASS;
    Node* tree = nullptr;
    int64_t evals = 0, searches = 0, n_moves = 0;

    ~Game() { clear_trees(); }
    void clear_trees() {
        delete players[0].root; players[0].root = nullptr;
        delete players[1].root; players[1].root = nullptr;
        delete tree; tree = nullptr;
    }
};

float random_komi(Rng& rng) {                                    // lib.rs:210-224
    float v = (float)rng.uniform();
    if (v < 0.4f) return 7.5f;
    if (v < 0.8f) return 6.5f;
    if (v < 0.9f) return 0.5f;
    return (float)(rng.below(16) - 8) + 0.5f;
}

struct Driver {
    dg_selfplay_config cfg;
    std::vector<Game> games;
    int64_t started = 0, finished = 0;
    std::string sgf_all;
    uint64_t digest = 0;
    int64_t total_moves = 0, total_evals = 0, total_searches = 0, total_cache_hits = 0;
    PredictionCache* shared_cache = nullptr;                     // one table for every game (cfg.cache_shared), as predictors/nn.rs:48-50
    bool device_ladders = false;                                 // DG_SELFPLAY_DEVICE_LADDERS

    void start_game(Game& g) {
        g.clear_trees();
        g.id = started++;
        g.rng.reseed(cfg.seed * 0x9e3779b97f4a7c15ull + (uint64_t)g.id * 0xd1342543de82ef95ull + 1);
        g.board.init(random_komi(g.rng));
        g.players[0] = Player();
        g.players[1] = Player();
        g.players[0].color = BLACK;
        g.players[1].color = WHITE;
        g.pass_count = 0;
        g.active = true;
        g.mode = Game::IDLE;
        g.sgf.clear();
        g.moves.clear();
        g.cache.reset(cfg.cache_capacity > 0 && !shared_cache ? new PredictionCache((size_t)cfg.cache_capacity) : nullptr);
    }

    void finish_game(Game& g) {                                  // game_result.rs:23-93 (Ended)
        uint8_t status[N_POINTS];
        territory_status(g.board, status);
        int black = 0, white = 0;
        std::string tb, tw;                                      // get_territory_as_sgf (:45-66)
        for (int p = 0; p < N_POINTS; ++p) {
            if (status[p] == WHITE) { ++white; tw += "["; sgf_point(p, tw); tw += "]"; }
            else if (status[p] == BLACK) { ++black; tb += "["; sgf_point(p, tb); tb += "]"; }
        }
        float w = (float)white + g.board.komi, b = (float)black;  // get_winner_as_sgf (:78-93)
        char head[160], res[32];
        if (b > w) snprintf(res, sizeof(res), "B+%.1f", b - w);
        else if (w > b) snprintf(res, sizeof(res), "W+%.1f", w - b);
        else snprintf(res, sizeof(res), "0");
        snprintf(head, sizeof(head), "(;GM[1]FF[4]SZ[19]RU[Chinese]KM[%.1f]RE[%s]", g.board.komi, res);
        sgf_all += head;
        sgf_all += g.sgf;
        if (!tb.empty()) { sgf_all += "TB"; sgf_all += tb; }
        if (!tw.empty()) { sgf_all += "TW"; sgf_all += tw; }
        sgf_all += ")\n";
        uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)g.id;
        for (uint16_t m : g.moves) { h ^= m; h *= 0x100000001b3ull; }
        digest += h * 0x9e3779b97f4a7c15ull;
        total_moves += g.n_moves;
        total_evals += g.evals;
        total_searches += g.searches;
        if (g.cache) { total_cache_hits += g.cache->hits.load(); g.cache.reset(); }
        g.n_moves = g.evals = g.searches = 0;
        g.active = false;
        g.clear_trees();
        ++finished;
    }

    void begin_search(Game& g, bool ex_it) {                     // Player::predict / predict_aux (self_play.rs:243-276, 321-358)
        Player& p = g.players[0];
        SearchOptions opt;
        opt.search_kind = g.allow_pass ? STANDARD_SEARCH : SCORING_SEARCH;
        opt.deterministic = !g.allow_pass;                       // ScoringSearch::deterministic() (options.rs:160-162)
        opt.probes_per_round = cfg.probes_per_round > 0 ? cfg.probes_per_round : 1;
        opt.dirichlet_beta = cfg.dirichlet_noise;
        opt.temperature = cfg.temperature;
        opt.num_rollout = ex_it ? cfg.num_ex_it_rollout : p.num_rollout(cfg.num_rollout);
        opt.policy_only = !ex_it && opt.num_rollout <= 1;
        opt.cache = shared_cache ? shared_cache : g.cache.get();
        opt.device_ladders = device_ladders;
        Node* tree = p.root;
        p.root = nullptr;
        if (tree && !g.allow_pass) tree->disqualify(PASS);
        g.task.start(g.board, p.color, opt, tree, g.rng.next());
        g.mode = ex_it ? Game::EX_IT : Game::SEARCH;
        ++g.searches;
    }

    void record(Game& g, int color, int index, bool with_value, float value, const Node* tree, int rollouts, const float* softmax) {
        std::string& s = g.sgf;                                  // Display for Played (self_play.rs:187-214), without C[]
        s += color == BLACK ? ";B[" : ";W[";
        sgf_point(index, s);
        s += "]";
        if (tree) {
            // Node::prior() = argmax_f32 over the dense prior (tree.rs:1266-1270, asm/argmax.rs tie rule)
            int arg = PASS;
            float bestp = NEG_INF;
            for (int i = 0; i < tree->n_cand; ++i) {
                int m = tree->cand_move[i];
                float p = tree->cand_prior[i];
                if (p > bestp || (p == bestp && ((m >> 3) > (arg >> 3) || ((m >> 3) == (arg >> 3) && m < arg)))) { bestp = p; arg = m; }
            }
            if (arg != PASS) { s += "TR["; sgf_point(arg, s); s += "]"; }
        }
        if (rollouts > 1 && softmax) {
            char tv[32];
            snprintf(tv, sizeof(tv), "TV[%d]P[", rollouts);
            s += tv;
            b85_encode_f16(softmax, 362, s);
            s += "]";
        }
        if (with_value) {
            char v[32];
            snprintf(v, sizeof(v), "V[%.4f]", color == BLACK ? 2.0f * value - 1.0f : -2.0f * value + 1.0f);
            s += v;
        }
    }

};

}  // namespace

extern "C" {

int32_t dg_engine_predict(void* engine, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy) {
    dg_engine* e = static_cast<dg_engine*>(engine);
    const int32_t chunk = dg_engine_max_batch(e);
    if (chunk <= 0) return DG_ERR_INVALID_ARGUMENT;
    for (int32_t at = 0; at < n; at += chunk) {
        int32_t m = n - at < chunk ? n - at : chunk;
        int32_t rc = dg_engine_forward_packed(e, positions + at, m, value + at, policy + (size_t)at * 362);
        if (rc) return rc;
    }
    return DG_OK;
}

// `RandomPredictor` (predictors/random.rs:30-59) made a function of the position: value uniform in (-1, 1), policy
// 362 uniform numbers normalised to 1.  Host-only: measures the search / feature path without a device.
int32_t dg_random_predict(void* ctx, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy) {
    uint64_t salt = ctx ? *static_cast<const uint64_t*>(ctx) : 0;
    for (int i = 0; i < n; ++i) {
        uint64_t h = 0xcbf29ce484222325ull ^ salt;
        for (int p = 0; p < N_POINTS; ++p) { h ^= positions[i].planes[p]; h *= 0x100000001b3ull; }
        Rng rng(h ^ positions[i].k_bits);
        value[i] = f32_to_f16_bits((float)(2.0 * rng.uniform() - 1.0));
        float x[362], total = 0.0f;
        for (int k = 0; k < 362; ++k) { x[k] = (float)rng.uniform(); total += x[k]; }
        float recip = 1.0f / total;
        for (int k = 0; k < 362; ++k) policy[(size_t)i * 362 + k] = f32_to_f16_bits(x[k] * recip);
    }
    return DG_OK;
}

// A peaked stand-in for a trained network (host only, measurement): policy = softmax(sharpness * u) with u uniform per
// point and position, so a handful of moves carry the mass and searches go deep instead of wide.  ctx = float* sharpness
// (NULL: 12).
int32_t dg_peaked_predict(void* ctx, const dg_packed_position* positions, int32_t n, uint16_t* value, uint16_t* policy) {
    const float sharp = ctx ? *static_cast<const float*>(ctx) : 12.0f;
    for (int i = 0; i < n; ++i) {
        uint64_t h = 0xcbf29ce484222325ull;
        for (int p = 0; p < N_POINTS; ++p) { h ^= positions[i].planes[p]; h *= 0x100000001b3ull; }
        Rng rng(h ^ positions[i].k_bits);
        value[i] = f32_to_f16_bits((float)(0.6 * rng.uniform() - 0.3));
        float x[362], total = 0.0f;
        for (int k = 0; k < 362; ++k) { x[k] = std::exp(sharp * ((float)rng.uniform() - 1.0f)); total += x[k]; }
        float recip = 1.0f / total;
        for (int k = 0; k < 362; ++k) policy[(size_t)i * 362 + k] = f32_to_f16_bits(x[k] * recip);
    }
    return DG_OK;
}

int32_t dg_mcts_predict(dg_predict_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                        const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                        int64_t* evals_out) {
    if (!predictor || !valid_search_call(options, board, color)) { delete N(starting_tree); return DG_ERR_INVALID_ARGUMENT; }
    SearchTask task;
    task.start(*reinterpret_cast<const Board*>(board), color, convert(options), N(starting_tree), options->seed);
    std::vector<dg_packed_position> batch;
    std::vector<uint16_t> value, policy;
    for (;;) {
        batch.clear();
        int n = task.emit(batch);
        if (n == 0) break;
        value.resize(n);
        policy.resize((size_t)n * 362);
        int32_t rc = predictor(ctx, batch.data(), n, value.data(), policy.data());
        if (rc) return rc;
        task.absorb(value.data(), policy.data());
    }
    if (value_out) *value_out = task.value();
    if (index_out) *index_out = task.index();
    if (evals_out) *evals_out = task.evals();
    Node* root = task.take_root();
    if (tree_out) *tree_out = reinterpret_cast<dg_tree*>(root);
    else delete root;
    return DG_OK;
}

dg_cache* dg_cache_new(int32_t capacity) { return reinterpret_cast<dg_cache*>(new PredictionCache(capacity > 0 ? (size_t)capacity : 0)); }
dg_cache* dg_cache_new_shared(int32_t capacity, int32_t stripes) {
    return reinterpret_cast<dg_cache*>(new PredictionCache(capacity > 0 ? (size_t)capacity : 0, stripes > 0 ? stripes : 64));
}
void dg_cache_free(dg_cache* cache) { delete reinterpret_cast<PredictionCache*>(cache); }
void dg_cache_stats(const dg_cache* cache, int64_t* hits, int64_t* misses, int64_t* size) {
    const PredictionCache* c = reinterpret_cast<const PredictionCache*>(cache);
    if (hits) *hits = c->hits.load();
    if (misses) *misses = c->misses.load();
    if (size) *size = (int64_t)c->size();
}

void dg_tree_free(dg_tree* tree) { delete N(tree); }
dg_tree* dg_tree_forward(dg_tree* tree, int32_t index) {           // an index outside 0..361 has no sub-tree: the tree is consumed
    if (!tree) return nullptr;
    if (index < 0 || index > PASS) { delete N(tree); return nullptr; }
    return reinterpret_cast<dg_tree*>(forward(N(tree), index));
}
void dg_tree_disqualify(dg_tree* tree, int32_t index) { if (tree && index >= 0 && index <= PASS) N(tree)->disqualify(index); }
int32_t dg_tree_total_count(const dg_tree* tree) { return N(tree)->total_count; }
int32_t dg_tree_to_move(const dg_tree* tree) { return N(tree)->to_move; }
float dg_tree_initial_value(const dg_tree* tree) { return N(tree)->initial_value; }
void dg_tree_children(const dg_tree* tree, int32_t* count, float* value, float* prior) {
    const Node* n = N(tree);
    for (int i = 0; i < 362; ++i) {
        if (count) count[i] = 0;
        if (value) value[i] = n->initial_value;
        if (prior) prior[i] = NEG_INF;
    }
    for (int k = 0; k < n->n_edges; ++k) {
        if (count) count[n->e_move[k]] = n->e_count[k];
        if (value) value[n->e_move[k]] = n->e_value[k];
    }
    if (prior) for (int i = 0; i < n->n_cand; ++i) prior[n->cand_move[i]] = n->cand_prior[i];
}
int64_t dg_tree_num_nodes(const dg_tree* tree) { return tree ? count_nodes(N(tree)) : 0; }

// `engines` != nullptr: the product path -- every group of games owns a leaf batch (include/dg_engine.h) of one of the
// engines (group g -> engine g mod n_engines, the round-robin of predictors/nn.rs:87-89): the worker that advanced the
// group's last game pushes the leaves and submits them with one graph launch, and whichever worker looks for work next
// notices the completion flag -- no device thread, no blocking call, no condition variable on that path.
static int32_t selfplay_impl(dg_predict_fn predictor, dg_predict_raw_fn raw_predictor, dg_predict_prior_fn prior_predictor, void* ctx,
                             dg_engine* const* engines, int32_t n_engines, uint32_t engine_flags,
                             const dg_selfplay_config* config, dg_selfplay_stats* stats, char* sgf_out, int64_t sgf_capacity) {
    const bool engine_mode = engines != nullptr;
    if ((!predictor && !raw_predictor && !prior_predictor && !engine_mode) || (engine_mode && n_engines <= 0) || !config ||
        config->num_games <= 0 || config->num_parallel <= 0)
        return DG_ERR_INVALID_ARGUMENT;
    const bool raw_mode = raw_predictor != nullptr || prior_predictor != nullptr || engine_mode;
    const bool prior_mode = prior_predictor != nullptr || (engine_mode && (engine_flags & DG_SELFPLAY_DEVICE_PRIORS));   // (the start value under DG_SELFPLAY_AUTO_PRIORS)
    Driver d;
    d.cfg = *config;
    d.device_ladders = engine_mode && (engine_flags & DG_SELFPLAY_DEVICE_LADDERS);
    std::unique_ptr<PredictionCache> process_table;
    if (config->cache_capacity > 0 && config->cache_shared) {
        process_table.reset(new PredictionCache((size_t)config->cache_capacity, config->cache_shared));
        d.shared_cache = process_table.get();
    }
    if (d.cfg.max_plies <= 0 || d.cfg.max_plies > 722) d.cfg.max_plies = 722;
    int hw = (int)std::thread::hardware_concurrency();
    int n_threads = d.cfg.num_threads > 0 ? d.cfg.num_threads : (hw > 0 ? hw : 1);
    int n_slots = std::min(config->num_parallel, config->num_games);
    const int n_workers_total = std::max(1, std::min(n_threads, n_slots));
    d.games = std::vector<Game>(n_slots);
    for (Game& g : d.games) d.start_game(g);

    // games alternate in groups: while the device evaluates one group's leaves the host works on the others.  More
    // groups hide more host time, fewer groups make larger device batches; aim at >= 128-256 leaves per batch
    int want_groups = config->num_groups;
    if (const char* env = getenv("DG_SELFPLAY_GROUPS")) want_groups = atoi(env);
    const int round_leaves = std::max(8, config->probes_per_round);            // a game emits its probes or the 8 root symmetries
    const int per_engine_slots = engine_mode ? (n_slots + n_engines - 1) / n_engines : n_slots;
    if (want_groups <= 0) {
        const int leaves = per_engine_slots * (config->probes_per_round > 0 ? config->probes_per_round : 1);
        want_groups = std::max(2, std::min(4, leaves / 256));
    }
    if (want_groups > 8) want_groups = 8;
    if (engine_mode) {                        // a group owns a workspace of its engine and its leaves must fit one leaf batch
        int cap = dg_engine_max_batch(engines[0]), max_groups = 8;
        for (int i = 0; i < n_engines; ++i) {
            cap = std::min(cap, (int)dg_engine_max_batch(engines[i]));
            max_groups = std::min(max_groups, (int)dg_engine_num_workspaces(engines[i]));
        }
        if (want_groups > max_groups) want_groups = max_groups;
        while (want_groups < max_groups && ((per_engine_slots + want_groups - 1) / want_groups) * round_leaves > cap) ++want_groups;
        if (want_groups < 1 || ((per_engine_slots + want_groups - 1) / want_groups) * round_leaves > cap) return DG_ERR_INVALID_ARGUMENT;
        want_groups *= n_engines;
    }
    const int n_groups = std::min(want_groups, n_slots);
    // A group is either on the host (its games are advanced one by one by whichever worker is free; the worker that
    // finishes the last one gathers the leaves and submits them) or on the device (its device thread is inside the
    // blocking predictor call / its leaf batch is in flight).  Nobody waits for a particular group: workers take the next
    // game of ANY group that has results, so the host stays busy while at least one group is back and the device while at
    // least one is submitted.
    struct Group {
        std::vector<int> slots;
        std::vector<dg_packed_position> batch;
        std::vector<dg_raw_position> raw_batch;
        std::vector<uint16_t> value, policy;
        std::vector<uint8_t> legal;
        std::vector<float> prior;
        size_t n_leaves = 0;
        size_t size() const { return n_leaves; }
        // where the evaluations of the round are read from: the vectors above, or the pinned outputs of the leaf batch
        const uint16_t *r_value = nullptr, *r_policy = nullptr;
        const uint8_t* r_legal = nullptr;
        const float* r_prior = nullptr;
        dg_leaf_batch* lb = nullptr;                              // engine mode
        bool with_prior = false;                                  // this round's batch also carries the leaves' priors
        int64_t t_submit = 0;
        enum State { HOST, SUBMITTED, RETIRED } state = HOST;     // guarded by sched_m
        size_t next = 0, done = 0;                                // games handed out / advanced in this round (sched_m)
        bool absorb = false;                                      // the round starts from results of the device
        std::thread device_thread;
        std::condition_variable cv;                               // wakes the device thread (with sched_m)
    };
    std::vector<std::unique_ptr<Group>> group_store;
    for (int gi = 0; gi < n_groups; ++gi) group_store.emplace_back(new Group());
    auto groups = [&](int gi) -> Group& { return *group_store[gi]; };
    for (int i = 0; i < n_slots; ++i) groups(i % n_groups).slots.push_back(i);
    if (engine_mode) {
        for (int gi = 0; gi < n_groups; ++gi) {
            int32_t arc = dg_engine_batch_acquire(engines[gi % n_engines], &groups(gi).lb);
            if (arc) {
                for (int k = 0; k < gi; ++k) dg_engine_batch_release(groups(k).lb);
                return arc;
            }
        }
    }

    auto t_start = std::chrono::steady_clock::now();
    auto seconds = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    std::atomic<int64_t> eval_ns{0};
    int64_t rounds = 0, positions = 0;
    int32_t rc = DG_OK;

    // self_play_one's loop body after `predict` returned (self_play.rs:439-457); true when the game has ended.
    auto after_move = [&](Game& g) {
        Player& me = g.players[0];
        int index = g.index;
        g.moves.push_back((uint16_t)index);
        ++g.n_moves;
        g.mode = Game::IDLE;
        if (index == PASS) {
            g.pass_count += 1;
            if (g.pass_count >= 2 && is_scorable(g.board)) { g.active = false; g.n_emitted = -1; return true; }
        } else {
            g.pass_count = 0;
            g.board.place(me.color, index);
        }
        Player& other = g.players[1];
        if (other.root) other.root = forward(other.root, index);
        std::swap(g.players[0], g.players[1]);
        return false;
    };

    // Advances one game until it has leaves for the device (returns with g.batch filled) or has nothing to do.
    auto advance = [&](Game& g, const uint16_t* value, const uint16_t* policy, const uint8_t* legal, const float* prior) {
        if (!g.active) return;
        if (g.n_emitted > 0) {
            g.task.absorb(value + g.batch_offset, policy + g.batch_offset * 362, legal ? legal + g.batch_offset * 361 : nullptr,
                          prior ? prior + g.batch_offset * 368 : nullptr);
            g.evals += g.n_emitted;
            g.n_emitted = 0;
        }
        g.batch.clear();
        g.raw_batch.clear();
        for (;;) {
            if (g.mode == Game::IDLE) {
                if ((int)g.board.count >= d.cfg.max_plies) { g.mode = Game::IDLE; g.active = false; g.n_emitted = -1; return; }   // ended: 722 plies
                g.allow_pass = is_scorable(g.board);
                d.begin_search(g, false);
            }
            int n = raw_mode ? g.task.emit(nullptr, &g.raw_batch) : g.task.emit(&g.batch, nullptr);
            if (n > 0) { g.n_emitted = n; return; }
            // the search is done
            Player& me = g.players[0];
            if (g.mode == Game::SEARCH) {
                g.value = g.task.value();
                g.index = g.task.index();
                delete g.tree;
                g.tree = g.task.take_root();
                bool policy_only = g.task.policy_only();
                if (!policy_only && !std::isfinite(g.value)) {  // self_play.rs:333-336: pass, forget the tree
                    delete g.tree;
                    g.tree = nullptr;
                    g.index = PASS;
                    d.record(g, me.color, PASS, false, 0.0f, nullptr, 0, nullptr);
                    if (after_move(g)) return;
                    continue;
                }
                // is_good_candidate (self_play.rs:287-291): value within [-0.8, 0.8] (a winrate: only the upper bound
                // bites) and, only then, a 5 % draw
                if (d.cfg.ex_it && g.value >= -0.80f && g.value <= 0.80f && (float)g.rng.uniform() < 0.05f) {
                    d.begin_search(g, true);
                    continue;
                }
                if (policy_only) {
                    d.record(g, me.color, g.index, true, g.value, nullptr, 1, nullptr);
                } else {
                    float softmax[362];
                    visit_distribution(*g.tree, softmax);
                    d.record(g, me.color, g.index, true, g.value, g.tree, g.tree->total_count, softmax);
                }
            } else {                                             // EX_IT: statistics from the second search, move from the first
                Node* deep = g.task.take_root();
                float softmax[362];
                visit_distribution(*deep, softmax);
                d.record(g, me.color, g.index, true, g.task.value(), deep, deep->total_count, softmax);
                delete deep;
            }
            me.update(g.value);
            if (g.tree) { me.root = forward(g.tree, g.index); g.tree = nullptr; }
            if (after_move(g)) return;
        }
    };

    // DG_SELFPLAY_TRACE=1: where the time goes (stderr, at the end)
    const bool trace_driver = getenv("DG_SELFPLAY_TRACE") != nullptr;
    phase_clock().on = trace_driver;
    auto now_ns = [] { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::atomic<int64_t> ns_worker_idle{0}, ns_worker_busy{0}, ns_worker_cpu{0}, ns_device_cpu{0};
    auto thread_cpu_ns = [] { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return (int64_t)ts.tv_sec * 1000000000 + ts.tv_nsec; };
    int64_t ns_serial = 0;                    // guarded by gather_m
    int calls_in_flight = 0;                  // wall time during which no predictor call was in flight (sched_m)
    int64_t idle_since = 0, ns_idle = 0;
    // DG_SELFPLAY_AUTO_PRIORS: where the leaves' priors are built follows which side is scarce.  The signal is how busy the
    // worker threads are (time inside advance() / finalize() over threads x wall time, per window of 64 batches): above 85 %
    // the host is the limit and the priors move to the device (that takes a fifth of the host's work per leaf away); below
    // 60 % they move back, because on the device they only lengthen the batches; a decision stands for at least four
    // windows.  (Whether the device runs out of work is NOT the signal: with two groups of games it does so 10 % of the time
    // for structural reasons, and device priors then cost 4 %.)  The priors are bit-identical either way, so the games do not
    // depend on the switch.
    const bool auto_priors = engine_mode && (engine_flags & DG_SELFPLAY_AUTO_PRIORS) != 0;
    bool priors_on_device = prior_mode;
    std::atomic<int64_t> busy_ns{0};
    int64_t win_start = 0, win_idle = 0, win_busy = 0, prior_switches = 0, prior_batches = 0;
    int win_batches = 0, win_hold = 0;

    std::mutex sched_m;                       // group states, task cursors, rc, stop
    std::condition_variable sched_cv;         // workers: "a group came back from the device" / "stop"
    std::mutex gather_m;                      // the Driver's shared state (game ids, records, totals) and the counters below
    int live_groups = n_groups;
    bool stop = false, quit = false;
    auto out_of_time = [&] { return d.cfg.max_seconds > 0 && seconds() >= d.cfg.max_seconds; };
    auto retire = [&](Group& grp) {           // sched_m held
        grp.state = Group::RETIRED;
        if (--live_groups == 0) stop = true;
    };

    // The worker that advanced the last game of a round: finished games are replaced, leaves are gathered in slot order,
    // the batch goes to the group's device thread.
    auto finalize = [&](Group& grp) {
        std::unique_lock<std::mutex> gl(gather_m);
        const int64_t t_begin = trace_driver ? now_ns() : 0;
        grp.batch.clear();
        grp.raw_batch.clear();
        grp.n_leaves = 0;
        if (grp.lb) dg_leaf_batch_reset(grp.lb);                 // every game of the group has consumed its results
        int32_t push_rc = DG_OK;
        for (int s : grp.slots) {
            Game& g = d.games[s];
            while (g.n_emitted == -1) {                          // (a game answered entirely by the transposition table ends at once)
                g.n_emitted = 0;
                g.active = true;
                d.finish_game(g);
                if (d.started < d.cfg.num_games && !out_of_time()) {
                    d.start_game(g);
                    advance(g, nullptr, nullptr, nullptr, nullptr);
                }
            }
            if (g.active && g.n_emitted > 0) {
                g.batch_offset = grp.size();
                if (grp.lb) {                                    // straight into the pinned input array of the leaf batch
                    if (dg_leaf_batch_push(grp.lb, g.raw_batch.data(), (int32_t)g.raw_batch.size()) != (int32_t)g.batch_offset) push_rc = DG_ERR_INVALID_ARGUMENT;
                } else {
                    grp.batch.insert(grp.batch.end(), g.batch.begin(), g.batch.end());
                    grp.raw_batch.insert(grp.raw_batch.end(), g.raw_batch.begin(), g.raw_batch.end());
                }
                grp.n_leaves += g.batch.size() + g.raw_batch.size();
            }
        }
        const bool empty = grp.size() == 0;
        if (!empty) {
            if (!grp.lb) {
                grp.value.resize(grp.size());
                grp.policy.resize(grp.size() * 362);
                if (raw_mode) grp.legal.resize(grp.size() * 361);
                if (prior_mode) grp.prior.resize(grp.size() * 368);
                grp.r_value = grp.value.data(); grp.r_policy = grp.policy.data();
                grp.r_legal = grp.legal.data(); grp.r_prior = grp.prior.data();
            }
            ++rounds;
            positions += (int64_t)grp.size();
        }
        if (trace_driver) ns_serial += now_ns() - t_begin;
        gl.unlock();
        if (!empty && grp.lb) {
            // one graph launch: H2D, planes + legal moves, tower, heads, (priors), D2H, completion flag
            bool with_prior;
            { std::lock_guard<std::mutex> lk(sched_m); with_prior = priors_on_device; }
            int32_t r = push_rc ? push_rc : dg_leaf_batch_submit(grp.lb, with_prior ? DG_LEAF_PRIOR : 0u);
            std::lock_guard<std::mutex> lk(sched_m);
            if (r != DG_OK) {
                if (rc == DG_OK) rc = r;
                retire(grp);
                stop = true;
            } else {
                grp.with_prior = with_prior;
                grp.t_submit = now_ns();
                if (with_prior) ++prior_batches;
                if (calls_in_flight++ == 0 && idle_since) {
                    ns_idle += grp.t_submit - idle_since;
                    win_idle += grp.t_submit - idle_since;
                }
                if (auto_priors) {
                    if (win_start == 0) win_start = grp.t_submit;
                    if (++win_batches >= 64) {
                        const int64_t busy = busy_ns.load(std::memory_order_relaxed);
                        const double load = (double)(busy - win_busy) / ((double)n_workers_total * (double)std::max<int64_t>(1, grp.t_submit - win_start));
                        const bool want = load > 0.85 ? true : load < 0.60 ? false : priors_on_device;
                        if (win_hold > 0) --win_hold;
                        else if (want != priors_on_device) { priors_on_device = want; ++prior_switches; win_hold = 4; }
                        win_start = grp.t_submit;
                        win_busy = busy;
                        win_idle = 0;
                        win_batches = 0;
                    }
                }
                grp.state = Group::SUBMITTED;
            }
            if (r != DG_OK) sched_cv.notify_all();
            return;
        }
        {
            std::lock_guard<std::mutex> lk(sched_m);
            if (empty) retire(grp); else grp.state = Group::SUBMITTED;
        }
        if (empty) sched_cv.notify_all(); else grp.cv.notify_one();
    };

    // engine mode, sched_m held: the leaf batch of a submitted group has completed (or failed)
    auto complete = [&](Group& grp, int32_t ready) {
        const int64_t t = now_ns();
        eval_ns += t - grp.t_submit;
        if (--calls_in_flight == 0) idle_since = t;
        if (ready < 0) {
            if (rc == DG_OK) rc = ready;
            retire(grp);
            stop = true;
        } else if (out_of_time()) {
            retire(grp);                                       // the work in flight at the deadline is dropped
        } else {
            grp.r_value = dg_leaf_batch_value(grp.lb); grp.r_policy = dg_leaf_batch_policy(grp.lb);
            grp.r_legal = dg_leaf_batch_legal(grp.lb); grp.r_prior = dg_leaf_batch_prior(grp.lb);
            grp.absorb = true;
            grp.next = grp.done = 0;
            grp.state = Group::HOST;
        }
    };

    auto device_loop = [&](Group& grp) {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(sched_m);
                grp.cv.wait(lk, [&] { return grp.state == Group::SUBMITTED || quit; });
                if (quit) { if (trace_driver) ns_device_cpu += thread_cpu_ns(); return; }
                if (trace_driver && calls_in_flight++ == 0 && idle_since) ns_idle += now_ns() - idle_since;
            }
            auto t0 = std::chrono::steady_clock::now();
            int32_t r = prior_predictor
                ? prior_predictor(ctx, grp.raw_batch.data(), (int32_t)grp.raw_batch.size(), grp.value.data(), grp.policy.data(), grp.legal.data(),
                                  grp.prior.data())
                : raw_predictor
                ? raw_predictor(ctx, grp.raw_batch.data(), (int32_t)grp.raw_batch.size(), grp.value.data(), grp.policy.data(), grp.legal.data())
                : predictor(ctx, grp.batch.data(), (int32_t)grp.batch.size(), grp.value.data(), grp.policy.data());
            eval_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> lk(sched_m);
                if (trace_driver && --calls_in_flight == 0) idle_since = now_ns();
                if (r != DG_OK) {
                    if (rc == DG_OK) rc = r;
                    retire(grp);
                    stop = true;
                } else if (out_of_time()) {
                    retire(grp);                                   // the work in flight at the deadline is dropped
                } else {
                    grp.absorb = true;
                    grp.next = grp.done = 0;
                    grp.state = Group::HOST;
                }
            }
            sched_cv.notify_all();
        }
    };

    auto worker_loop = [&] {
        int64_t t_mark = trace_driver ? now_ns() : 0;
        prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);               // the short naps below are meant to be short
        for (;;) {
            Group* grp = nullptr;
            size_t i = 0;
            {
                std::unique_lock<std::mutex> lk(sched_m);
                for (int idle_polls = 0;; ++idle_polls) {
                    if (stop) { if (trace_driver) ns_worker_cpu += thread_cpu_ns(); return; }
                    if (engine_mode)                              // has a leaf batch come back?  (a read of pinned memory each)
                        for (int gi = 0; gi < n_groups; ++gi)
                            if (groups(gi).state == Group::SUBMITTED)
                                if (int32_t ready = dg_leaf_batch_ready(groups(gi).lb)) complete(groups(gi), ready);
                    if (stop) continue;
                    for (int gi = 0; gi < n_groups && !grp; ++gi)
                        if (groups(gi).state == Group::HOST && groups(gi).next < groups(gi).slots.size()) { grp = &groups(gi); i = grp->next++; }
                    if (grp) break;
                    if (trace_driver) { const int64_t t = now_ns(); ns_worker_busy += t - t_mark; t_mark = t; }
                    if (engine_mode) {
                        // nothing to do until a batch returns: poll its flag -- a few pauses first, then naps that leave
                        // the core to a sibling thread
                        lk.unlock();
                        if (idle_polls < 64) { for (int k = 0; k < 32; ++k) __builtin_ia32_pause(); }
                        else { timespec nap{0, 5000}; nanosleep(&nap, nullptr); }
                        lk.lock();
                    } else {
                        sched_cv.wait(lk);
                    }
                    if (trace_driver) { const int64_t t = now_ns(); ns_worker_idle += t - t_mark; t_mark = t; }
                }
            }
            const bool absorb = grp->absorb;
            const int64_t t_work = auto_priors ? now_ns() : 0;
            advance(d.games[grp->slots[i]], absorb ? grp->r_value : nullptr, absorb ? grp->r_policy : nullptr,
                    absorb && raw_mode ? grp->r_legal : nullptr,
                    absorb && (engine_mode ? grp->with_prior : prior_mode) ? grp->r_prior : nullptr);
            bool last;
            { std::lock_guard<std::mutex> lk(sched_m); last = ++grp->done == grp->slots.size(); }
            if (auto_priors) busy_ns.fetch_add(now_ns() - t_work, std::memory_order_relaxed);
            if (last) finalize(*grp);
        }
    };

    if (!engine_mode)
        for (int gi = 0; gi < n_groups; ++gi) groups(gi).device_thread = std::thread(device_loop, std::ref(groups(gi)));
    {
        std::vector<std::thread> workers;
        const int n_workers = n_workers_total;
        for (int i = 1; i < n_workers; ++i) workers.emplace_back(worker_loop);
        const int64_t cpu_before = trace_driver ? thread_cpu_ns() : 0;   // the calling thread has a history
        worker_loop();
        if (trace_driver) ns_worker_cpu -= cpu_before;
        sched_cv.notify_all();
        for (auto& t : workers) t.join();
    }
    {   // device threads still inside a call (deadline, error) come back first
        { std::lock_guard<std::mutex> lk(sched_m); quit = true; }
        for (int gi = 0; gi < n_groups; ++gi) {
            if (engine_mode) dg_engine_batch_release(groups(gi).lb);          // waits for a batch still in flight (deadline, error)
            else { groups(gi).cv.notify_all(); groups(gi).device_thread.join(); }
        }
    }

    if (trace_driver)
        fprintf(stderr, "[dg_selfplay] %d groups, %d threads, %lld rounds: workers busy %.3f s / idle %.3f s (sum over threads), serial gather %.3f s, "
                "predictor calls %.3f s (sum over device threads), no call in flight %.3f s, CPU time: workers %.3f s, device threads %.3f s, wall %.3f s\n",
                n_groups, n_threads, (long long)rounds, ns_worker_busy.load() * 1e-9, ns_worker_idle.load() * 1e-9, ns_serial * 1e-9,
                (double)eval_ns.load() * 1e-9, ns_idle * 1e-9, ns_worker_cpu.load() * 1e-9, ns_device_cpu.load() * 1e-9, seconds());
    if (trace_driver && engine_mode)
        fprintf(stderr, "[dg_selfplay] priors on the device for %lld of %lld batches, %lld switches%s\n", (long long)prior_batches, (long long)rounds,
                (long long)prior_switches, auto_priors ? " (auto)" : "");
    if (trace_driver) {
        PhaseClock& c = phase_clock();
        const double n = (double)std::max<uint64_t>(1, c.leaves.load());
        fprintf(stderr, "[dg_selfplay] cycles per leaf: board copy %.0f, probe %.0f, extract %.0f, prior plan %.0f, prior apply %.0f, insert %.0f\n",
                c.copy.load() / n, c.probe.load() / n, c.extract.load() / n, c.plan.load() / n, c.apply.load() / n, c.insert.load() / n);
    }
    // account for the games that were cut off by max_seconds
    for (Game& g : d.games) {
        d.total_moves += g.n_moves; d.total_evals += g.evals; d.total_searches += g.searches;
        if (g.cache) d.total_cache_hits += g.cache->hits.load();
    }
    if (d.shared_cache) d.total_cache_hits += d.shared_cache->hits.load();
    if (stats) {
        stats->games_finished = d.finished;
        stats->moves = d.total_moves;
        stats->evals = positions;
        stats->rounds = rounds;
        stats->searches = d.total_searches;
        stats->seconds = seconds();
        stats->eval_seconds = (double)eval_ns.load() * 1e-9;
        stats->mean_batch = rounds ? (double)positions / (double)rounds : 0.0;
        stats->digest = d.digest;
        stats->cache_hits = d.total_cache_hits;
    }
    if (sgf_out && sgf_capacity > 0) {
        size_t n = std::min<size_t>(d.sgf_all.size(), (size_t)sgf_capacity - 1);
        memcpy(sgf_out, d.sgf_all.data(), n);
        sgf_out[n] = 0;
    }
    return rc;
}

int32_t dg_selfplay_run(dg_predict_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                        char* sgf_out, int64_t sgf_capacity) {
    return selfplay_impl(predictor, nullptr, nullptr, ctx, nullptr, 0, 0u, config, stats, sgf_out, sgf_capacity);
}

int32_t dg_selfplay_run_raw(dg_predict_raw_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                            char* sgf_out, int64_t sgf_capacity) {
    return selfplay_impl(nullptr, predictor, nullptr, ctx, nullptr, 0, 0u, config, stats, sgf_out, sgf_capacity);
}

int32_t dg_selfplay_run_prior(dg_predict_prior_fn predictor, void* ctx, const dg_selfplay_config* config, dg_selfplay_stats* stats,
                              char* sgf_out, int64_t sgf_capacity) {
    return selfplay_impl(nullptr, nullptr, predictor, ctx, nullptr, 0, 0u, config, stats, sgf_out, sgf_capacity);
}

int32_t dg_selfplay_run_engine(dg_engine* const* engines, int32_t n_engines, uint32_t flags, const dg_selfplay_config* config,
                               dg_selfplay_stats* stats, char* sgf_out, int64_t sgf_capacity) {
    if (!engines || n_engines <= 0) return DG_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_engines; ++i) if (!engines[i]) return DG_ERR_INVALID_ARGUMENT;
    return selfplay_impl(nullptr, nullptr, nullptr, nullptr, engines, n_engines, flags, config, stats, sgf_out, sgf_capacity);
}

int32_t dg_engine_predict_prior(void* engine, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy, uint8_t* legal,
                                float* prior) {
    dg_engine* e = static_cast<dg_engine*>(engine);
    const int32_t chunk = dg_engine_max_batch(e);
    if (chunk <= 0) return DG_ERR_INVALID_ARGUMENT;
    for (int32_t at = 0; at < n; at += chunk) {
        int32_t m = n - at < chunk ? n - at : chunk;
        int32_t rc = dg_engine_forward_raw_prior(e, positions + at, m, value + at, policy + (size_t)at * 362, legal + (size_t)at * 361,
                                                 prior + (size_t)at * 368);
        if (rc) return rc;
    }
    return DG_OK;
}

int32_t dg_mcts_predict_prior(dg_predict_prior_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                              const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                              int64_t* evals_out) {
    if (!predictor || !valid_search_call(options, board, color)) { delete N(starting_tree); return DG_ERR_INVALID_ARGUMENT; }
    SearchTask task;
    task.start(*reinterpret_cast<const Board*>(board), color, convert(options), N(starting_tree), options->seed);
    std::vector<dg_raw_position> batch;
    std::vector<uint16_t> value, policy;
    std::vector<uint8_t> legal;
    std::vector<float> prior;
    for (;;) {
        batch.clear();
        int n = task.emit(nullptr, &batch);
        if (n == 0) break;
        value.resize(n);
        policy.resize((size_t)n * 362);
        legal.resize((size_t)n * 361);
        prior.resize((size_t)n * 368);
        int32_t rc = predictor(ctx, batch.data(), n, value.data(), policy.data(), legal.data(), prior.data());
        if (rc) return rc;
        task.absorb(value.data(), policy.data(), legal.data(), prior.data());
    }
    if (value_out) *value_out = task.value();
    if (index_out) *index_out = task.index();
    if (evals_out) *evals_out = task.evals();
    Node* root = task.take_root();
    if (tree_out) *tree_out = reinterpret_cast<dg_tree*>(root);
    else delete root;
    return DG_OK;
}

int32_t dg_engine_predict_raw(void* engine, const dg_raw_position* positions, int32_t n, uint16_t* value, uint16_t* policy, uint8_t* legal) {
    dg_engine* e = static_cast<dg_engine*>(engine);
    const int32_t chunk = dg_engine_max_batch(e);
    if (chunk <= 0) return DG_ERR_INVALID_ARGUMENT;
    for (int32_t at = 0; at < n; at += chunk) {
        int32_t m = n - at < chunk ? n - at : chunk;
        int32_t rc = dg_engine_forward_raw(e, positions + at, m, value + at, policy + (size_t)at * 362, legal + (size_t)at * 361);
        if (rc) return rc;
    }
    return DG_OK;
}

int32_t dg_mcts_predict_raw(dg_predict_raw_fn predictor, void* ctx, const dg_search_options* options, dg_tree* starting_tree,
                            const dg_board* board, int32_t color, float* value_out, int32_t* index_out, dg_tree** tree_out,
                            int64_t* evals_out) {
    if (!predictor || !valid_search_call(options, board, color)) { delete N(starting_tree); return DG_ERR_INVALID_ARGUMENT; }
    SearchTask task;
    task.start(*reinterpret_cast<const Board*>(board), color, convert(options), N(starting_tree), options->seed);
    std::vector<dg_raw_position> batch;
    std::vector<uint16_t> value, policy;
    std::vector<uint8_t> legal;
    for (;;) {
        batch.clear();
        int n = task.emit(nullptr, &batch);
        if (n == 0) break;
        value.resize(n);
        policy.resize((size_t)n * 362);
        legal.resize((size_t)n * 361);
        int32_t rc = predictor(ctx, batch.data(), n, value.data(), policy.data(), legal.data());
        if (rc) return rc;
        task.absorb(value.data(), policy.data(), legal.data());
    }
    if (value_out) *value_out = task.value();
    if (index_out) *index_out = task.index();
    if (evals_out) *evals_out = task.evals();
    Node* root = task.take_root();
    if (tree_out) *tree_out = reinterpret_cast<dg_tree*>(root);
    else delete root;
    return DG_OK;
}

}  // extern "C"
