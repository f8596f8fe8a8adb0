/*
 * dg_engine.h -- C ABI of the B200 (sm_100a) self-play inference engine.
 *
 * This is the drop-in boundary for dream-go's neural-network hot path.  The
 * reference boundary is a Rust crate API (`src/libdg_nn/lib.rs:33-36`:
 * `Network`, `Workspace`, `WorkspaceGuard`, `forward`, `OutputMap`, `Error`)
 * whose only native edge is `libdg_cuda`'s `extern "C"` bindings to
 * cudart/cuDNN (the files under `src/libdg_cuda/cudnn/`).  A maintainer replaces the body
 * of `libdg_nn` with ~100 lines of `extern "C"` calls into this library (the
 * stub is shown in INTEGRATION.md); nothing above `dg_nn::forward` changes.
 *
 * Plain pointers and sizes only -- no torch, no C++ types.  All functions are
 * thread-safe on one engine unless noted.  One engine owns one CUDA device.
 *
 * Status codes: 0 = OK, negatives map onto the reference's
 * `enum Error { CuDNN(Status), Cuda(Error), MalformedWeights, MissingWeights }`
 * (`src/libdg_nn/error.rs:19-24`).
 */
#ifndef DG_ENGINE_H
#define DG_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_OK                      0
#define DG_ERR_CUDA               -1  /* Error::Cuda(..)   -- a CUDA runtime/driver call failed */
#define DG_ERR_KERNEL             -2  /* Error::CuDNN(..)  -- the compute path rejected the request (there is no cuDNN here) */
#define DG_ERR_MALFORMED_WEIGHTS  -3  /* Error::MalformedWeights */
#define DG_ERR_MISSING_WEIGHTS    -4  /* Error::MissingWeights   */
#define DG_ERR_INVALID_ARGUMENT   -5  /* reference: debug_assert / panic (graph.rs:124-125, nn.rs:85) */

/* Network geometry (src/libdg_go/utils/features.rs:88-96, layers/common.rs:22-25). */
#define DG_NUM_FEATURES   32
#define DG_NUM_POINTS     361
#define DG_POLICY_SIZE    362
#define DG_FEATURE_SIZE   (DG_NUM_POINTS * DG_NUM_FEATURES)   /* 11,552 fp16 per position */

/* dg_engine_config.flags */
#define DG_FLAG_DEBUG_DIRECT_CONV  0x1u  /* tests only: run every convolution on the slow one-thread-per-output
                                            cross-check kernel instead of the tcgen05 kernel */

#define DG_FLAG_NO_PDL             0x2u  /* debug: launch the layers without programmatic dependent launch */
#define DG_FLAG_LAYERWISE          0x4u  /* debug: one launch per convolution instead of the persistent tower kernel */
#define DG_FLAG_NO_ROTATE          0x8u  /* debug: do not rotate the unit -> CTA-pair assignment between layers */
#define DG_FLAG_TOWER_LATE_A      0x40u  /* debug (A/B): the tower kernel requests a layer's first activation windows after its filter slabs */
#define DG_FLAG_SEPARATE_HEAD_CONV 0x80u /* debug (A/B): the head convolution runs as its own launch instead of as the tower kernel's last layer */
#define DG_FLAG_NO_GRAPH          0x20u  /* debug: leaf batches are enqueued call by call instead of as one captured graph */
#define DG_FLAG_BLOCKING_SYNC     0x10u  /* blocking calls poll the stream between 20 us naps instead of spinning in the driver
                                            (self-play: every core is busy searching and a spinning waiter steals one) */

typedef struct dg_engine dg_engine;

/* Replaces the implicit configuration of `Network::new()` + `Builder::get_workspace`
 * (src/libdg_nn/network.rs:92-124, graph.rs:50-74) and the `SOFTMAX_TEMPERATURE`
 * global (src/libdg_utils/config.rs:176-177, baked in at policy_head.rs:46). */
typedef struct dg_engine_config {
    int32_t  device;               /* CUDA ordinal; the reference picks it with cudaSetDevice (predictors/nn.rs:87-89) */
    int32_t  max_batch;            /* largest batch a single forward may carry (>= 1) */
    float    softmax_temperature;  /* <= 0 -> 0.709888 */
    int32_t  num_workspaces;       /* forwards that may be in flight concurrently; <= 0 -> 2 (nn.rs:64-67) */
    uint32_t flags;
} dg_engine_config;

/* One named tensor of the weight file, as `loader.rs:36-100` would decode it:
 * dtype is the file's "t" field ("f2", "f4", "i4", "i1"), data the decoded
 * little-endian elements.  Names are the reference's JSON keys
 * ("01_upsample/conv_1:0", ... -- SURVEY.md Appendix A). */
typedef struct dg_tensor_view {
    const char* name;
    const char* dtype;
    const void* data;
    uint64_t    nbytes;
} dg_tensor_view;

/* Lossless compact form of one V1 feature tensor (features.rs:154-250): every
 * plane is binary except planes 0/1 which are the constant `k`, so a position
 * is 361 32-bit plane masks + k.  bit c of planes[p] set <=> feature (p, c) is
 * non-zero; bits 0 and 1 mean "value k_bits", all others mean 1.0. */
typedef struct dg_packed_position {
    uint32_t planes[DG_NUM_POINTS];
    uint16_t k_bits;               /* fp16 bits of k = clamp(0.5 + 0.5*komi/7.5, 0, 1) (features.rs:236) */
    uint16_t reserved;
} dg_packed_position;

/* The smallest description of a position from which the DEVICE computes the V1 feature planes and the legal moves
 * (csrc/features.cu): stones, which points have ever held a stone (the super-ko rule only looks at those,
 * board.rs:135), the zobrist hash of the position and of the last 16 positions (board.rs:132-141; the keys are the
 * engine's own, csrc/go_board.h), the last two moves (361 = none) and the two ladder planes -- read on the host (a ladder is
 * a sequential search) or, with DG_RAW_DEVICE_LADDERS, left to the device.  Bit p of a mask = point p = 19*y + x.  `dg_board_raw_position` (dg_go.h) fills it. */
typedef struct dg_raw_position {
    uint32_t black[12], white[12], visited[12], ladder_capture[12], ladder_escape[12];
    uint64_t hash;
    uint64_t hash_history[16];
    int16_t  last_move[2];
    uint16_t k_bits;               /* fp16 bits of k (features.rs:236) */
    uint8_t  to_move;              /* 1 black, 2 white (anything else: DG_ERR_INVALID_ARGUMENT from the calls that take positions) */
    uint8_t  symmetry;             /* bits 0-2: orientation the planes are produced in (symmetry::ALL order);
                                      bit 3: DG_RAW_DEVICE_LADDERS -- the two ladder masks above are not filled in, the device
                                      reads the ladders itself (one warp per reading, utils/ladder.rs:53-179);
                                      bits 4-7: search options of the position (DG_STANDARD_SEARCH / DG_SCORING_SEARCH) */
} dg_raw_position;                 /* 384 bytes */
#define DG_RAW_DEVICE_LADDERS 0x08

/* ---- devices ------------------------------------------------------------------------------- */

/* `Device::len()` / `Device::all()` (src/libdg_cuda/devices.rs:55-64) restricted to what this engine runs on: the number
 * of CUDA devices if every one of them is an sm_100 part, else the number of leading sm_100 devices; 0 without a driver
 * or device (the reference's `Device::all().expect(..)` then panics, predictors/nn.rs:65).  One process may hold one
 * engine per device and call them from any thread. */
int32_t dg_device_count(void);
/* `cudaGetDevice` / `Device::set_current` (devices.rs:70-73) of the calling thread: the reference's NnPredictor picks a
 * device per batch with set_current and then asks the network for a workspace (predictors/nn.rs:87-93); a shim maps the
 * calling thread's current device to its engine with these.  -1 / DG_ERR_CUDA on failure. */
int32_t dg_current_device(void);
int32_t dg_set_current_device(int32_t device);

/* ---- lifetime ------------------------------------------------------------------------------ */

/* `Network::new()` minus the file search.  Fails with DG_ERR_CUDA when the device is missing
 * or is not an sm_100 part -- there is no CPU or library fallback. */
int32_t dg_engine_create(const dg_engine_config* config, dg_engine** out);
void    dg_engine_destroy(dg_engine* engine);

/* `loader::load(path)` (src/libdg_nn/loader.rs:102-116): JSON + base85 weight file.
 * Missing/unreadable/empty file -> DG_ERR_MISSING_WEIGHTS, undecodable -> DG_ERR_MALFORMED_WEIGHTS.
 * Weights are uploaded (and re-laid-out for the tensor cores) eagerly, replacing the lazy
 * per-tensor upload of src/libdg_nn/tensor.rs:123-141.  Not thread-safe against forwards. */
int32_t dg_engine_load_weights_json(dg_engine* engine, const char* path);
/* Same, from decoded tensors (tests and embedders that already hold the tensors). */
int32_t dg_engine_load_weights_raw(dg_engine* engine, const dg_tensor_view* tensors, int32_t count);

/* ---- the hot path -------------------------------------------------------------------------- */

/* `dg_nn::forward(&mut Workspace, &[f16]) -> OutputMap<f16>` (src/libdg_nn/graph.rs:123-158)
 * fused with `Network::get_workspace(batch)` (network.rs:132-143): blocking, host buffers.
 *   features   : batch * 11,552 fp16 bit patterns, NHWC (index 32*(19y+x)+c)
 *   value_out  : batch fp16 (post-tanh)            policy_out : batch * 362 fp16 (post-softmax)
 * Caller-owned buffers are not retained after return.  Buffers obtained from
 * dg_engine_alloc_host are DMA'd directly; any other pointer is staged. */
int32_t dg_engine_forward_f16(dg_engine* engine, const uint16_t* features, int32_t batch,
                              uint16_t* value_out, uint16_t* policy_out);

/* Same network evaluation from compact positions; the device expands them to NHWC fp16
 * (replaces the 23,104-byte-per-leaf host copy of pool/batch.rs:87-96 with 1,448 bytes). */
int32_t dg_engine_forward_packed(dg_engine* engine, const dg_packed_position* positions, int32_t batch,
                                 uint16_t* value_out, uint16_t* policy_out);

/* The same evaluation from raw positions: the device derives the feature planes itself and also returns
 * `Board::is_valid(to_move, .)` for the 361 points of every position (identity orientation) -- what
 * create_initial_policy (pool/policy_helper.rs:39-43) asks the board for.  384 bytes H2D per position. */
int32_t dg_engine_forward_raw(dg_engine* engine, const dg_raw_position* positions, int32_t batch,
                              uint16_t* value_out, uint16_t* policy_out, uint8_t* legal_out /* [batch][361] */);
/* The same, and the priors the search inserts: `create_initial_policy` + `add_valid_candidates` + `normalize_policy(1.0)`
 * (pool/policy_helper.rs:28-134, as the Insert event of pool/worker_thread.rs:88-93 runs them) computed on the device --
 * candidate mask of the position's search options (bits 4.. of `symmetry`: 0 = StandardSearch, 1 = ScoringSearch incl.
 * Benson's unconditional life and the own-eye heuristic, libdg_mcts/options.rs:53-138), orbit folding on symmetric
 * boards, inverse symmetry, renormalisation.  prior_out: batch x 368 floats (-inf = not a candidate), bit-identical to
 * dg_board_prior on the same policy. */
int32_t dg_engine_forward_raw_prior(dg_engine* engine, const dg_raw_position* positions, int32_t batch,
                                    uint16_t* value_out, uint16_t* policy_out, uint8_t* legal_out, float* prior_out);
/* Only the feature stage of dg_engine_forward_raw: the planes as compact positions + the legal masks (tests, tools). */
int32_t dg_engine_features_raw(dg_engine* engine, const dg_raw_position* positions, int32_t batch,
                               dg_packed_position* planes_out, uint8_t* legal_out /* [batch][361] */);

/* ---- leaf-batch queue (replaces pool::Batcher, src/libdg_mcts/pool/batch.rs:61-124) ---------- */

/* The reference gathers leaves under a mutex (23 KB memcpy each), cuts batches of <= --batch-size with at most
 * 2 x devices alive (batch.rs:98-123, predictors/nn.rs:64-67) and the worker that cut a batch blocks in
 * `batch.forward(predictor)` (pool/worker_thread.rs:88-99).  Here a leaf batch is one of the engine's `num_workspaces`
 * in-flight evaluations: any number of producer threads claim slots of its pinned input array lock-free (one
 * compare-and-swap, 384 bytes per leaf), ONE CUDA-graph launch evaluates them -- host-to-device copy, feature planes and
 * legal moves from the raw stones, tower, heads, optionally the ready-to-insert priors, device-to-host copies, all
 * captured once per batch size bucket -- and completion is a word of pinned host memory written by the last kernel of
 * the launch, so waiting costs no driver call and any thread can notice it.  Results stay valid until the batch is reset.
 * Outputs are bit-identical to dg_engine_forward_raw(_prior) on the same positions. */
typedef struct dg_leaf_batch dg_leaf_batch;
#define DG_LEAF_PRIOR 0x1u           /* dg_leaf_batch_submit: also compute the priors (dg_engine_forward_raw_prior) */

/* Takes one of the engine's workspaces (blocks while all are taken) / gives it back (waits for a submit in flight). */
int32_t dg_engine_batch_acquire(dg_engine* engine, dg_leaf_batch** out);
void    dg_engine_batch_release(dg_leaf_batch* batch);
int32_t dg_leaf_batch_capacity(const dg_leaf_batch* batch);              /* = the engine's max_batch */
/* Lock-free multi-producer: appends `n` positions, returns the index of the first one, or -1 when they do not fit or the
 * batch is sealed (submitted and not yet reset) -- `Batcher::push` (batch.rs:87-91); DG_ERR_INVALID_ARGUMENT for a position
 * whose to_move is not a colour (nothing is appended). */
int32_t dg_leaf_batch_push(dg_leaf_batch* batch, const dg_raw_position* positions, int32_t n);
/* Seals the batch and launches its evaluation; returns at once -- `Batcher::get_batch` + `Batch::forward` without the
 * blocking (batch.rs:98-123).  Any thread may call it, once per fill. */
int32_t dg_leaf_batch_submit(dg_leaf_batch* batch, uint32_t outputs);
/* 1 when the results of the last submit are in the output arrays (a plain memory read), 0 while it runs. */
int32_t dg_leaf_batch_ready(dg_leaf_batch* batch);
/* Blocks until ready (spinning on the flag, or in 20 us naps under DG_FLAG_BLOCKING_SYNC); DG_ERR_CUDA if the launch failed. */
int32_t dg_leaf_batch_wait(dg_leaf_batch* batch);
int32_t dg_leaf_batch_size(const dg_leaf_batch* batch);                  /* leaves of the last submit */
dg_raw_position* dg_leaf_batch_slots(dg_leaf_batch* batch);              /* the pinned input array (single-producer use; the writer
                                                                            vouches for to_move = 1 / 2, dg_leaf_batch_push checks it) */
/* Pinned output arrays of the last submit, in push order: [n] fp16, [n][362] fp16, [n][361], [n][368] floats. */
const uint16_t* dg_leaf_batch_value(const dg_leaf_batch* batch);
const uint16_t* dg_leaf_batch_policy(const dg_leaf_batch* batch);
const uint8_t*  dg_leaf_batch_legal(const dg_leaf_batch* batch);
const float*    dg_leaf_batch_prior(const dg_leaf_batch* batch);         /* only after DG_LEAF_PRIOR */
/* Empties the batch for the next fill (after the results have been consumed). */
void    dg_leaf_batch_reset(dg_leaf_batch* batch);

/* ---- weight file (device-independent; usable without an engine) ------------------------------ */

/* Parses `path` exactly like dg_engine_load_weights_json would and reports one tensor:
 * its "s" scale and decoded size in bytes (0 / 0 when `name` is absent or NULL).  Returns DG_OK,
 * DG_ERR_MISSING_WEIGHTS or DG_ERR_MALFORMED_WEIGHTS; num_tensors (optional) = entries found.
 * (The reference's loader test, src/libdg_nn/loader.rs:124-142, asserts on exactly these.) */
int32_t dg_weights_file_probe(const char* path, const char* name, int32_t* num_tensors, float* scale, uint64_t* nbytes);

/* ---- housekeeping -------------------------------------------------------------------------- */

/* `Network::synchronize()` (network.rs:145-159): waits for all in-flight work on the device. */
int32_t dg_engine_synchronize(dg_engine* engine);
/* Pinned host memory the forward calls can DMA from/to directly. */
void*   dg_engine_alloc_host(dg_engine* engine, uint64_t nbytes);
void    dg_engine_free_host(dg_engine* engine, void* ptr);
/* Last error text of this engine: a copy owned by the calling thread, valid until that thread asks again. */
const char* dg_engine_last_error(dg_engine* engine);
/* Network shape discovered from the weights (graph.rs:76-96). */
int32_t dg_engine_num_blocks(dg_engine* engine);
/* The max_batch / num_workspaces (= leaf batches that can be held at once) the engine was created with. */
int32_t dg_engine_max_batch(dg_engine* engine);
int32_t dg_engine_num_workspaces(dg_engine* engine);
/* Library/ABI version, for the FFI shim to assert on. */
int32_t dg_engine_abi_version(void);

/* ---- measurement hooks (bench.py, tests) ---------------------------------------------------- */

/* Runs `iters` forwards of `batch` positions whose inputs are ALREADY resident in device
 * memory (the last batch given to dg_engine_forward_*), each step timed with its own CUDA event
 * pair on the engine's stream; with flush_l2 != 0 a 256 MiB memset evicts L2 between steps
 * (outside the timed pairs).  ms_total = summed device time of all steps.
 * tower_ms (optional) = same for the residual-block convolution launches only (2 per block),
 * measured in a second pass.  launches (optional) = kernels launched per forward. */
int32_t dg_engine_time_resident(dg_engine* engine, int32_t batch, int32_t iters, int32_t flush_l2,
                                float* ms_total, float* tower_ms, int32_t* launches);
/* End-to-end rate of the blocking call: `callers` host threads (the reference runs 2 per device,
 * predictors/nn.rs:64-67) each issue dg_engine_forward_f16 from their own pinned buffers (filled with `features`)
 * until `steps` calls have been made in total.  seconds = wall time of those calls. */
int32_t dg_engine_time_e2e(dg_engine* engine, const uint16_t* features, int32_t batch, int32_t steps, int32_t callers,
                           double* seconds);
/* Re-evaluates the resident batch up to `layer` (0 = up-sample, i = residual block i,
 * -1 = last block) and copies that activation into out[batch][361][128] fp16 (tests). */
int32_t dg_engine_debug_read_tower(dg_engine* engine, int32_t layer, int32_t batch, uint16_t* out);

/* Runs ONE residual-block convolution (block 0, conv_2 with skip) on the resident batch with
 * in-kernel clock64() tracing and copies min(out_len, SMs*3*64) samples to out, laid out
 * [cta][role: 0 TMA producer, 1 MMA issuer, 2 epilogue][64] (perf debugging). */
int32_t dg_engine_debug_conv_trace(dg_engine* engine, int32_t batch, int64_t* out, int32_t out_len);
/* Same for one launch of the persistent tower kernel: [cta][role][1024] samples.  Producer: (before flag wait,
 * after flag wait) per unit; MMA issuer: (unit start, accumulator free, operands of k-half 0 landed, unit issued)
 * per unit; epilogue (first warp): (before accumulator wait, after, unit stored) per unit. */
int32_t dg_engine_debug_tower_trace(dg_engine* engine, int32_t batch, int64_t* out, int32_t out_len);

#ifdef __cplusplus
}
#endif
#endif /* DG_ENGINE_H */
