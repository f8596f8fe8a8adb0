// Residual-tower 3x3 convolution (128 -> 128 channels) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// Replaces cudnnConvolutionBiasActivationForward as driven by `ResidualLayer::forward`
// (src/libdg_nn/layers/residual_block.rs:64-78, conv2d.rs:171-220):
//     y   = relu(conv1(x) + b1)
//     out = relu(g * conv2(y) + (1 - g) * x + fp16(g * b2))
//
// One persistent cluster of two CTAs per SM pair.  Each unit of work is 256 consecutive board rows
// (two 128-row tiles, one per CTA) x all 128 output channels:
//   * each CTA keeps HALF of the filter bank resident in shared memory for the whole launch
//     (its 64 output channels x 9 taps x 128 input channels = 144 KiB, TMA-loaded once); the pair's
//     UMMA (M=256, N=128, K=16) reads both halves, so weights are never re-fetched per tile and each
//     SM reads only 96 B/cycle of operands from shared memory (a single CTA needs 128-192 B/cycle);
//   * activations: per tile and 64-channel k-half ONE TMA box of 170 rows (tile + 21 halo rows
//     each side); the nine taps are UMMA descriptors starting at nine row offsets in that box;
//   * accumulators: 2 x 128 TMEM columns per CTA (double buffered) so the epilogue of unit i
//     overlaps the MMAs of unit i+1;
//   * epilogue (8 warps, one thread per row and 64-channel half): the skip row is prefetched with
//     256-bit loads before the accumulator is ready; tcgen05.ld -> g*acc + b (+ (1-g)*skip) ->
//     ReLU -> halo rows := 0 -> fp16 -> 256-bit stores (every access is one full 32-byte sector).
// warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA only), warps 2..9 = epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "conv_tc.h"
#include "layout.h"
#include "ptx.cuh"

namespace dg {

namespace t2 {
constexpr int kStages = 3;
constexpr int kThreads = 320;
constexpr int kWindowBytes = DG_WINDOW_ROWS * 128;      // 21,760
constexpr int kStageBytes = 22 * 1024;
constexpr int kSlab = 64 * 128;                          // [64 out][64 in] fp16
template <int NH>
struct Smem {
    static constexpr int kWeights = 9 * NH * kSlab;      // 147,456 for the tower (NH = 2)
    static constexpr int kAOff = kWeights;
    static constexpr int kBarOff = kAOff + kStages * kStageBytes;
    static constexpr int kTotal = kBarOff + 1024 + 1024; // + barriers/bias + alignment slack
};
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kIdesc = umma_idesc_f16(256, 128);
}  // namespace t2

template <int NH, int H, int I>
__device__ __forceinline__ void issue_one(uint32_t d_tmem, uint32_t a_lo, uint32_t w_lo, uint32_t idesc) {
    constexpr int tap = I / 4, k = I % 4;
    constexpr int row_off = DG_HALO_ROWS + (tap / 3 - 1) * DG_LINE_STRIDE + (tap % 3 - 1);
    umma_f16_ss_pair<((row_off * 128 + k * 32) >> 4), (((tap * NH + H) * t2::kSlab + k * 32) >> 4)>(
        d_tmem, a_lo, w_lo, kUmmaDescHiSw128, idesc, (H | I) != 0);
}
template <int NH, int H, int... I>
__device__ __forceinline__ void issue_half(uint32_t d_tmem, uint32_t a_lo, uint32_t w_lo, std::integer_sequence<int, I...>,
                                           uint32_t idesc = t2::kIdesc) {
    (issue_one<NH, H, I>(d_tmem, a_lo, w_lo, idesc), ...);
}

// NH = number of 64-channel k-halves of the input: 2 for the tower (128 channels), 1 for the
// up-sampling layer (32 feature planes zero-padded to 64).
template <int NH, bool SKIP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(t2::kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w, ConvTcParams p) {
    using namespace t2;
    constexpr int kWeights = Smem<NH>::kWeights, kAOff = Smem<NH>::kAOff, kBarOff = Smem<NH>::kBarOff;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + kAOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = bars + 1 + kStages;
    uint64_t* acc_full = bars + 1 + 2 * kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* bias_s = reinterpret_cast<float*>(bars + 32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int first_unit = blockIdx.x >> 1;
    const int unit_step = gridDim.x >> 1;
    const int nunits = (p.ntiles + 1) >> 1;
    constexpr bool has_skip = SKIP;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_act);
        tma_prefetch_desc(&tm_w);
        mbar_init(w_full, 1);
        for (int i = 0; i < kStages; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 16);      // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < 128; i += kThreads) bias_s[i] = p.bias[i];
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    griddep_launch_dependents();   // the next layer may start its prologue / weight loads (it waits for us before touching activations)
    const uint32_t tmem_base = *tmem_slot;
    int tr = 0;
#define DG_TRACE(role)                                                                          \
    do {                                                                                        \
        if (p.trace && tr < 64) p.trace[(blockIdx.x * 3 + (role)) * 64 + tr++] = clock64();     \
    } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            // completion bytes of BOTH CTAs are credited to the leader's barriers
            const uint32_t w_full0 = mapa_shared(smem_u32(w_full), 0);
            if (rank == 0) mbar_expect_tx(w_full, 2 * kWeights);
            for (int tap = 0; tap < 9; tap++)
                for (int h = 0; h < NH; h++)
                    tma_load_2d_pair(w_s + (tap * NH + h) * kSlab, &tm_w, w_full0, h * 64, tap * 128 + rank * 64);
            griddep_wait();
            DG_TRACE(0);
            int stage = 0;
            uint32_t phase = 0;
            for (int u = first_unit; u < nunits; u += unit_step) {
                const int tile = 2 * u + rank;
                for (int h = 0; h < NH; h++) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    DG_TRACE(0);
                    if (rank == 0) mbar_expect_tx(&a_full[stage], 2 * kWindowBytes);
                    tma_load_2d_pair(a_s + stage * kStageBytes, &tm_act, mapa_shared(smem_u32(&a_full[stage]), 0), h * 64,
                                     DG_GUARD_ROWS + tile * DG_TILE_M - DG_HALO_ROWS);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one elected lane)
        if (rank == 0) {
            mbar_wait(w_full, 0);
            if (lane == 0) DG_TRACE(1);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            const uint32_t w_lo = umma_desc_lo(smem_u32(w_s));
            for (int u = first_unit; u < nunits; u += unit_step) {
                mbar_wait_cluster(&acc_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 128;
                // k-half 0
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                if (lane == 0) DG_TRACE(1);
                if (elect_one()) {
                    issue_half<NH, 0>(d_tmem, umma_desc_lo(smem_u32(a_s + stage * kStageBytes)), w_lo, std::make_integer_sequence<int, 36>{});
                    umma_commit_pair(&a_empty[stage], 3);
                    if (NH == 1) umma_commit_pair(&acc_full[as], 3);
                }
                __syncwarp();
                if (lane == 0) DG_TRACE(1);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
                if (NH == 2) {   // k-half 1
                    mbar_wait(&a_full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) DG_TRACE(1);
                    if (elect_one()) {
                        issue_half<NH, NH - 1>(d_tmem, umma_desc_lo(smem_u32(a_s + stage * kStageBytes)), w_lo, std::make_integer_sequence<int, 36>{});
                        umma_commit_pair(&a_empty[stage], 3);
                        umma_commit_pair(&acc_full[as], 3);
                    }
                    __syncwarp();
                    if (lane == 0) DG_TRACE(1);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (both CTAs)
        griddep_wait();
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
        const int colhalf = (warp - 2) >> 2;             // which 64 output channels
        const int row = quarter * 32 + lane;
        const bool tracer = (threadIdx.x == 64);
        const uint32_t acc_empty0[2] = {mapa_shared(smem_u32(&acc_empty[0]), 0), mapa_shared(smem_u32(&acc_empty[1]), 0)};
        const float* bias_h = bias_s + colhalf * 64;
        int as = 0;
        uint32_t aphase = 0;
        const float alpha = p.alpha, beta = p.beta;
        // the residual input of unit i+1 is prefetched (256-bit loads) before the epilogue math of unit i
        auto row_info = [&](int u, size_t& off) -> bool {
            const int m = (2 * u + static_cast<int>(rank)) * DG_TILE_M + row;
            const int q = m % DG_POS_ROWS;
            off = static_cast<size_t>(DG_GUARD_ROWS + m) * 128 + colhalf * 64;
            return (m >= p.valid_rows) || (q % DG_LINE_STRIDE == DG_LINE_STRIDE - 1) || (q >= DG_POS_ROWS - DG_LINE_STRIDE);
        };
        auto prefetch_skip = [&](int u, uint32_t (&dst)[32]) {
            size_t o;
            const bool h = (u >= nunits) || row_info(u, o);
#pragma unroll
            for (int i = 0; i < 32; i++) dst[i] = 0;
            if (has_skip && !h) {
#pragma unroll
                for (int i = 0; i < 4; i++) ld_global_256(p.skip + o + i * 16, &dst[i * 8]);
            }
        };
        uint32_t sk[32], sk_next[32];
        prefetch_skip(first_unit, sk);
        for (int u = first_unit; u < nunits; u += unit_step) {
            size_t off;
            const bool halo = row_info(u, off);
            prefetch_skip(u + unit_step, sk_next);
            if (tracer) DG_TRACE(2);
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
            if (tracer) DG_TRACE(2);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 128 + colhalf * 64;
            uint32_t acc0[32], acc1[32];
            tmem_ld_32x32b_x32(taddr, acc0);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 32, acc1);           // in flight while the first 32 channels are finished
            auto finish = [&](const uint32_t (&acc)[32], int part) {
                uint32_t packed[16];
                if (halo) {                                  // halo rows stay zero
#pragma unroll
                    for (int e = 0; e < 16; e++) packed[e] = 0;
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        float v0 = fmaf(alpha, __uint_as_float(acc[2 * e]), bias_h[part * 32 + 2 * e]);
                        float v1 = fmaf(alpha, __uint_as_float(acc[2 * e + 1]), bias_h[part * 32 + 2 * e + 1]);
                        if (has_skip) {
                            const float2 s2 = __half22float2(*reinterpret_cast<const __half2*>(&sk[part * 16 + e]));
                            v0 = fmaf(beta, s2.x, v0);
                            v1 = fmaf(beta, s2.y, v1);
                        }
                        const __half2 hv = __floats2half2_rn(fmaxf(v0, 0.f), fmaxf(v1, 0.f));   // NaN-non-propagating ReLU
                        packed[e] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                }
                st_global_256(p.out + off + part * 32, &packed[0]);
                st_global_256(p.out + off + part * 32 + 16, &packed[8]);
            };
            finish(acc0, 0);
            tmem_ld_wait();
            tc_fence_before();                               // accumulator fully read: hand it back to the MMA issuer
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty0[as]);
            finish(acc1, 1);
            if (tracer) DG_TRACE(2);
#pragma unroll
            for (int i = 0; i < 32; i++) sk[i] = sk_next[i];
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    if (lane == 0 && warp <= 1) DG_TRACE(warp);
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, t2::kTmemCols);
    }
#undef DG_TRACE
}

template <int NH, bool SKIP>
static cudaError_t launch_pair(const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p, int num_sms,
                               cudaStream_t stream, bool pdl) {
    static std::atomic<unsigned long long> configured{0};
    auto kernel = conv3x3_pair_kernel<NH, SKIP>;
    if (cudaError_t e = opt_in_shared_memory(kernel, t2::Smem<NH>::kTotal, configured); e != cudaSuccess) return e;
    const int nunits = (p.ntiles + 1) / 2;
    int pairs = num_sms / 2;
    if (pairs > nunits) pairs = nunits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(t2::kThreads);
    cfg.dynamicSmemBytes = t2::Smem<NH>::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, tm_act, tm_w, p);
}

cudaError_t launch_conv_pair(int k_halves, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p, int num_sms,
                             cudaStream_t stream, bool pdl) {
    if (k_halves == 1) return launch_pair<1, false>(tm_act, tm_w, p, num_sms, stream, pdl);
    return p.skip ? launch_pair<2, true>(tm_act, tm_w, p, num_sms, stream, pdl) : launch_pair<2, false>(tm_act, tm_w, p, num_sms, stream, pdl);
}

// =============================================================================================
// The whole tower (up-sampling layer + every residual-block convolution + the head convolution) as ONE
// persistent launch.
//
// Same CTA-pair machinery as conv3x3_pair_kernel, plus:
//   * no global barrier between layers: a 256-row unit of layer l only needs units u-1, u, u+1 of
//     layer l-1, published through per-tile progress flags (st.release by the epilogue; the TMA producer
//     reads a unit's three flags together with relaxed loads, then fence.acq_rel + fence.proxy.async).
//     The same RAW chain also covers every WAR hazard of the x/y ping-pong buffers;
//   * the last layer may be the head convolution (wrows = 16: 8 policy + 2 value samples + 6 zero channels,
//     N = 16 MMAs, 1 KiB slabs at the start of the slab slots) whose epilogue writes pbuf / vbuf;
//   * the filter bank of layer l+1 replaces layer l's in shared memory slab by slab (one 8 KiB slab per
//     (tap, k-half), in the order the MMAs consume them): in the last unit of a layer every tap commits
//     its own w_free barrier when its MMAs retire, the producer re-fetches that slab immediately, and in
//     the first unit of the next layer every tap waits for its own w_full barrier;
//   * the unit -> pair assignment rotates by `rot` pairs per layer so that the pairs that get the
//     extra (6th) unit differ from layer to layer.
// The grid must be fully co-resident (<= one CTA per SM, cooperative launch).
constexpr int kTowerSmem = t2::Smem<2>::kBarOff + 512 + 4 * 512 + 1024;   // + barriers + bias[2][2][128] + alignment slack

// Tiles of the tower kernel are aligned to positions: three 128-row tiles cover rows 0..383 of a position's 400 board rows;
// rows 384..399 lie in the all-zero halo line y = 19, are never computed and never written (they are zero from the
// allocation on).  3 tiles instead of 3.125 per position: 4 % fewer MMAs than tiling the row space blindly.
__device__ __forceinline__ int tower_tile_base(int tile) {
    const int n = tile / 3;
    return n * DG_POS_ROWS + (tile - 3 * n) * DG_TILE_M;
}

template <int UNUSED = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(t2::kThreads, 1)
tower_kernel(const __grid_constant__ TowerParams p) {
    using namespace t2;
    constexpr int kAOff = Smem<2>::kAOff, kBarOff = Smem<2>::kBarOff;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + kAOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* a_full = bars;                 // [kStages]
    uint64_t* a_empty = bars + kStages;      // [kStages]
    uint64_t* acc_full = bars + 2 * kStages; // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint64_t* w_full = acc_empty + 2;        // [18] one per filter slab (tap, k-half): slab = tap * 2 + h
    uint64_t* w_free = w_full + 18;          // [18]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
    float* bias_s = reinterpret_cast<float*>(bars + 64);      // [2 groups][2][128]
    // barriers (512 B) + two bias buffers (1 KiB) follow the activation ring; see kTowerSmem

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;
    const int nunits = (p.ntiles + 1) >> 1;
    auto first_unit = [&](int l) { return (pair + npairs - (l * p.rot) % npairs) % npairs; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; i++) tma_prefetch_desc(&p.act[i]);
        for (int i = 0; i < kStages; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8);           // 4 warps of one epilogue group x 2 CTAs
        }
        for (int i = 0; i < 18; i++) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_free[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    griddep_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;
    int tr = 0;
#define DG_TRACE(role)                                                                          \
    do {                                                                                        \
        if (p.trace && tr < 1024) p.trace[(blockIdx.x * 3 + (role)) * 1024 + tr++] = clock64(); \
    } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t a_full0[kStages];
            for (int i = 0; i < kStages; i++) a_full0[i] = mapa_shared(smem_u32(&a_full[i]), 0);
            int stage = 0;
            uint32_t phase = 0;
            const int ntiles_even = 2 * nunits;
            for (int l = 0; l < p.nlayers; l++) {
                const int nh = p.layer[l].nh;
                const CUtensorMap* tm_w = &p.w[l];
                const CUtensorMap* tm_a = &p.act[p.layer[l].in_map];
                // slab by slab, in the order the MMAs consume them: a slab of layer l is fetched as soon as the last unit
                // of layer l-1 is done with it, so the reload spreads over a whole unit time instead of two bursts
                // a layer's filter bank: per tap `wrows` output channels (128; 16 for the head convolution), half of them
                // in each CTA -- a slab is [wrows / 2 out][64 in] at the start of its 8 KiB slot
                const int wrows = p.layer[l].wrows, wbytes = wrows * 64;
                auto load_weights = [&](int h) {
                    for (int tap = 0; tap < 9; tap++) {
                        const int s = tap * 2 + h;
                        if (l > 0) mbar_wait(&w_free[s], (l - 1) & 1);      // layer l-1 no longer reads this slab
                        if (rank == 0) mbar_expect_tx(&w_full[s], 2 * wbytes);
                        tma_load_2d_pair(w_s + s * kSlab, tm_w, mapa_shared(smem_u32(&w_full[s]), 0), h * 64,
                                         tap * wrows + rank * (wrows >> 1));
                    }
                };
                // The first unit's activation windows do not depend on the filter bank: they are requested BEFORE the
                // producer sits in the w_free waits of the weight swap (p.late_a = the earlier order, for A/B runs).
                if (l == 0 || p.late_a) load_weights(0);
                if (l == 0) { griddep_wait(); DG_TRACE(0); }
                const int u0 = first_unit(l);
                for (int u = u0; u < nunits; u += npairs) {
                    const int tile = 2 * u + rank;
                    DG_TRACE(0);
                    if (l > 0) {        // rows tile*128-21 .. +149 of layer l-1 must be complete and visible
                        // the three flags are read together (one L2 round trip, not three dependent ones), then one fence
                        // turns the observation into an acquire
                        const uint32_t want = p.gen + l;
                        const uint32_t* f0 = p.done + (tile > 0 ? tile - 1 : tile);
                        const uint32_t* f2 = p.done + (tile + 1 < ntiles_even ? tile + 1 : tile);
                        for (;;) {
                            const uint32_t a = ld_relaxed_gpu(f0), b = ld_relaxed_gpu(p.done + tile), c = ld_relaxed_gpu(f2);
                            if (static_cast<int32_t>(a - want) >= 0 && static_cast<int32_t>(b - want) >= 0 && static_cast<int32_t>(c - want) >= 0) break;
                        }
                        fence_acq_rel_gpu();
                        fence_proxy_async_global();
                    }
                    for (int h = 0; h < nh; h++) {
                        if (u == u0 && h == 1 && p.late_a) load_weights(1);
                        if (h == 0) DG_TRACE(0);
                        mbar_wait(&a_empty[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&a_full[stage], 2 * kWindowBytes);
                        tma_load_2d_pair(a_s + stage * kStageBytes, tm_a, a_full0[stage], h * 64,
                                         DG_GUARD_ROWS + tower_tile_base(tile) - DG_HALO_ROWS);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        if (u == u0 && !p.late_a && (l > 0 || h == 1)) load_weights(h);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        if (rank == 0) {
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0, wfull_phase[2] = {0, 0};
            const uint32_t w_lo = umma_desc_lo(smem_u32(w_s));
            for (int l = 0; l < p.nlayers; l++) {
                const int nh = p.layer[l].nh;
                const uint32_t idesc = umma_idesc_f16(256, p.layer[l].wrows);      // N = 128, or 16 for the head convolution
                const int u0 = first_unit(l);
                for (int u = u0; u < nunits; u += npairs) {
                    const bool last = (u + npairs >= nunits);
                    if (lane == 0) DG_TRACE(1);
                    mbar_wait_cluster(&acc_empty[as], aphase ^ 1);
                    tc_fence_after();
                    if (lane == 0) DG_TRACE(1);
                    const uint32_t d_tmem = tmem_base + as * 128;
                    const bool first = (u == u0);
                    // In the first unit of a layer every tap waits for its own slab of the new filter bank; in the
                    // last unit every tap hands its slab back as soon as its MMAs have retired.
#define DG_TAP(H, T)                                                                                          \
    do {                                                                                                      \
        if (first) { mbar_wait(&w_full[(T) * 2 + (H)], wfull_phase[H]); tc_fence_after(); }                    \
        if (elect_one()) {                                                                                    \
            issue_half<2, H>(d_tmem, a_lo, w_lo, std::integer_sequence<int, (T) * 4, (T) * 4 + 1, (T) * 4 + 2, (T) * 4 + 3>{}, idesc); \
            if (last) umma_commit_pair(&w_free[(T) * 2 + (H)], 3);                                             \
        }                                                                                                     \
        __syncwarp();                                                                                         \
    } while (0)
#define DG_TAPS(H) DG_TAP(H, 0); DG_TAP(H, 1); DG_TAP(H, 2); DG_TAP(H, 3); DG_TAP(H, 4); DG_TAP(H, 5); DG_TAP(H, 6); DG_TAP(H, 7); DG_TAP(H, 8)
                    // k-half 0
                    mbar_wait(&a_full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) DG_TRACE(1);
                    {
                        const uint32_t a_lo = umma_desc_lo(smem_u32(a_s + stage * kStageBytes));
                        if (first || last) {
                            DG_TAPS(0);
                            if (first) wfull_phase[0] ^= 1;
                        } else if (elect_one()) {
                            issue_half<2, 0>(d_tmem, a_lo, w_lo, std::make_integer_sequence<int, 36>{}, idesc);
                        }
                        __syncwarp();
                        if (elect_one()) {
                            umma_commit_pair(&a_empty[stage], 3);
                            if (nh == 1) {
                                if (last)
                                    for (int t = 0; t < 9; t++) umma_commit_pair(&w_free[t * 2 + 1], 3);
                                umma_commit_pair(&acc_full[as], 3);
                            }
                        }
                        __syncwarp();
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    if (nh == 2) {   // k-half 1
                        mbar_wait(&a_full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_lo = umma_desc_lo(smem_u32(a_s + stage * kStageBytes));
                        if (first || last) {
                            DG_TAPS(1);
                            if (first) wfull_phase[1] ^= 1;
                        } else if (elect_one()) {
                            issue_half<2, 1>(d_tmem, a_lo, w_lo, std::make_integer_sequence<int, 36>{}, idesc);
                        }
                        __syncwarp();
                        if (elect_one()) {
                            umma_commit_pair(&a_empty[stage], 3);
                            umma_commit_pair(&acc_full[as], 3);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
#undef DG_TAPS
#undef DG_TAP
                    if (lane == 0) DG_TRACE(1);
                    if (++as == 2) { as = 0; aphase ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (both CTAs)
        // Two independent groups of four warps: group g drains accumulator stage g, i.e. every second unit
        // of this pair, one thread per row and all 128 output channels.  The groups run out of phase (one
        // reads TMEM / does the math while the other one's stores drain), and each has two unit-times per unit.
        griddep_wait();
        const int quarter = warp & 3;
        const int g = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const bool tracer = (threadIdx.x == 64);
        const bool leader = ((threadIdx.x - 64) & 127) == 0;
        const uint32_t acc_empty0 = mapa_shared(smem_u32(&acc_empty[g]), 0);
        float* bias_g = bias_s + g * 256;                  // [2][128] per group
        uint32_t aphase = 0;
        auto next_item = [&](int& l, int& u) {
            u += npairs;
            if (u >= nunits) {
                l++;
                if (l < p.nlayers) u = first_unit(l);
            }
        };
        auto row_info = [&](int u, size_t& off) -> bool {
            const int m = tower_tile_base(2 * u + static_cast<int>(rank)) + row;
            const int q = m % DG_POS_ROWS;
            off = static_cast<size_t>(DG_GUARD_ROWS + m) * 128;
            return (m >= p.valid_rows) || (q % DG_LINE_STRIDE == DG_LINE_STRIDE - 1) || (q >= DG_POS_ROWS - DG_LINE_STRIDE);
        };
        uint32_t sk[64];
        // residual input of unit (l, u): written two layers ago, possibly by another pair -> check its flag, then
        // read it with L2-coherent 256-bit loads; issued as soon as the previous unit of this group is done
        auto prefetch_skip = [&](int l, int u) {
            if (!p.layer[l].has_skip) return;
            size_t o;
            const bool h = row_info(u, o);
            if (lane == 0) {
                const uint32_t want = p.gen + l - 1;
                while (static_cast<int32_t>(ld_acquire_gpu(p.done + 2 * u + rank) - want) < 0) {
                }
            }
            __syncwarp();
            if (!h) {
                const __half* src = p.layer[l].skip + o;
#pragma unroll
                for (int i = 0; i < 8; i++) ld_global_cg_256(src + i * 16, &sk[i * 8]);
            }
        };
        int l = 0, u = first_unit(0);
        if (g == 1) next_item(l, u);
        int bias_layer = -1;
        if (l < p.nlayers) prefetch_skip(l, u);
        while (l < p.nlayers) {
            const TowerLayer& L = p.layer[l];
            float* bias_l = bias_g + (l & 1) * 128;
            if (bias_layer != l) {
                bias_l[(threadIdx.x - 64) & 127] = L.bias[(threadIdx.x - 64) & 127];
                named_bar_sync(3 + g, 128);
                bias_layer = l;
            }
            const bool has_skip = L.has_skip != 0;
            const float alpha = L.alpha, beta = L.beta;
            __half* out = L.out;
            size_t off;
            const bool halo = row_info(u, off);
            if (tracer) DG_TRACE(2);
            mbar_wait(&acc_full[g], aphase);
            aphase ^= 1;
            tc_fence_after();
            if (tracer) DG_TRACE(2);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * 128;
            auto finish = [&](const uint32_t (&acc)[32], int part) {
                uint32_t packed[16];
                if (halo) {
#pragma unroll
                    for (int e = 0; e < 16; e++) packed[e] = 0;
                } else if (has_skip) {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        float v0 = fmaf(alpha, __uint_as_float(acc[2 * e]), bias_l[part * 32 + 2 * e]);
                        float v1 = fmaf(alpha, __uint_as_float(acc[2 * e + 1]), bias_l[part * 32 + 2 * e + 1]);
                        const float2 s2 = __half22float2(*reinterpret_cast<const __half2*>(&sk[part * 16 + e]));
                        v0 = fmaf(beta, s2.x, v0);
                        v1 = fmaf(beta, s2.y, v1);
                        const __half2 hv = __floats2half2_rn(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
                        packed[e] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const float v0 = fmaf(alpha, __uint_as_float(acc[2 * e]), bias_l[part * 32 + 2 * e]);
                        const float v1 = fmaf(alpha, __uint_as_float(acc[2 * e + 1]), bias_l[part * 32 + 2 * e + 1]);
                        const __half2 hv = __floats2half2_rn(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
                        packed[e] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                }
                st_global_256(out + off + part * 32, &packed[0]);
                st_global_256(out + off + part * 32 + 16, &packed[8]);
            };
            if (L.wrows == 16) {
                // the head convolution (policy_head.rs:49-52, value_head.rs:45-49): 8 policy + 2 value samples (+ 6 zero channels)
                // per board row -> pbuf[row][8] (the A operand of the policy FC) and vbuf[row][2]
                uint32_t acc[16];
                tmem_ld_32x32b_x16(taddr, acc);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty0);
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float v0 = fmaf(alpha, __uint_as_float(acc[2 * j]), bias_l[2 * j]);
                    float v1 = fmaf(alpha, __uint_as_float(acc[2 * j + 1]), bias_l[2 * j + 1]);
                    v0 = (v0 > 0.f && !halo) ? v0 : 0.f;     // NaN-non-propagating ReLU; halo rows stay zero
                    v1 = (v1 > 0.f && !halo) ? v1 : 0.f;
                    const __half2 hv = __floats2half2_rn(v0, v1);
                    packed[j] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                const size_t grow = off >> 7;
                *reinterpret_cast<uint4*>(out + grow * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint32_t*>(L.out2 + grow * 2) = packed[4];
                named_bar_sync(1 + g, 128);
                if (leader) {
                    __threadfence();
                    st_release_gpu(p.done + 2 * u + rank, p.gen + l + 1);
                }
                if (tracer) DG_TRACE(2);
                next_item(l, u);
                if (l < p.nlayers) next_item(l, u);
                if (l < p.nlayers) prefetch_skip(l, u);
                continue;
            }
            uint32_t acc0[32], acc1[32];
            tmem_ld_32x32b_x32(taddr, acc0);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 32, acc1);
            finish(acc0, 0);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 64, acc0);
            finish(acc1, 1);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 96, acc1);
            finish(acc0, 2);
            tmem_ld_wait();
            tc_fence_before();                          // accumulator stage fully read: hand it back to the MMA issuer
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty0);
            finish(acc1, 3);
            // publish: this CTA's tile of layer l is complete once the four warps of the group have stored their rows
            named_bar_sync(1 + g, 128);
            if (leader) {
                __threadfence();
                st_release_gpu(p.done + 2 * u + rank, p.gen + l + 1);
            }
            if (tracer) DG_TRACE(2);
            next_item(l, u);
            if (l < p.nlayers) next_item(l, u);
            if (l < p.nlayers) prefetch_skip(l, u);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, t2::kTmemCols);
    }
#undef DG_TRACE
}

cudaError_t launch_tower(const TowerParams& p, int num_sms, cudaStream_t stream) {
    static std::atomic<unsigned long long> configured{0};
    auto kernel = tower_kernel<0>;
    if (cudaError_t e = opt_in_shared_memory(kernel, kTowerSmem, configured); e != cudaSuccess) return e;
    const int nunits = (p.ntiles + 1) / 2;
    int pairs = num_sms / 2;
    if (pairs > nunits) pairs = nunits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(t2::kThreads);
    cfg.dynamicSmemBytes = kTowerSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;     // every CTA must be resident: CTAs wait on each other's flags
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

}  // namespace dg
