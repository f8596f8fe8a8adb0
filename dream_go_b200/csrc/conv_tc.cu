// Head 3x3 convolutions (128 -> 8 policy samples + 2 value samples, one N=16 implicit GEMM) on the
// 5th-generation tensor cores.
//
// Replaces the two cudnnConvolutionBiasActivationForward calls of policy_head.rs:49-52 and
// value_head.rs:45-49:  p1 = relu(conv3x3(x, Wp) + bp),  v1 = relu(conv3x3(x, Wv) + bv).
// x is a board-row buffer (layout.h), the filters are KRSC fp16 re-laid out per tap.
//
// Design (one persistent CTA per SM, warp-specialised, 192 threads):
//   * weights of this CTA's output-channel slice stay RESIDENT in shared memory for the whole
//     launch (9 taps x NH k-halves x [BN x 64] fp16, 128B-swizzled, loaded once by TMA);
//   * per 128-row tile ONE TMA box of 170 rows x 64 channels (tile + 21-row halo each side) is
//     staged per k-half; the nine taps are nine UMMA descriptors that start at different ROW
//     offsets inside that box -- no im2col, activations are read from L2 1.33x instead of 9x.
//     (The 128B swizzle is a function of the absolute shared-memory address, so a descriptor may
//     start at any 128-byte row of a TMA-written box with base_offset = 0; verified on B200.)
//   * tcgen05.mma (M=128, N=BN, K=16) accumulates in TMEM (double-buffered accumulator) so the
//     epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue: tcgen05.ld -> alpha*acc + bias (+ beta*skip) -> ReLU -> zero the halo rows ->
//     fp16 -> global.
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.h"
#include "layout.h"
#include "ptx.cuh"

namespace dg {

constexpr int kStages = 3;                       // A half-window ring
constexpr int kWindowBytes = DG_WINDOW_ROWS * 128;   // 21,760 bytes landed per TMA box
constexpr int kStageBytes = 22 * 1024;           // 1024-aligned slot
constexpr int kThreads = 192;

template <int NH, int BN>
struct ConvSmem {
    static constexpr int kSlab = BN * 128;                 // one tap, one k-half: [BN][64] fp16
    static constexpr int kWeights = 9 * NH * kSlab;
    static constexpr int kBars = kWeights + kStages * kStageBytes;
    static constexpr int kTotal = kBars + 256 + BN * 4 + 1024;   // + barriers + bias + alignment slack
};

template <int NH, int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w, ConvTcParams p) {
    using L = ConvSmem<NH, BN>;
    constexpr int kNSplit = (BN == 64) ? 2 : 1;
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
    constexpr uint32_t kIdesc = umma_idesc_f16(DG_TILE_M, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 128B swizzle needs 1024-byte alignment
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + L::kWeights;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;
    uint64_t* a_empty = bars + 1 + kStages;
    uint64_t* acc_full = bars + 1 + 2 * kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* bias_s = reinterpret_cast<float*>(bars + 32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nslice = blockIdx.x % kNSplit;           // which BN-wide slice of the output channels
    const int first_tile = blockIdx.x / kNSplit;
    const int tile_step = gridDim.x / kNSplit;

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
            mbar_init(&acc_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < BN; i += kThreads) bias_s[i] = p.bias[nslice * BN + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    int tr = 0;      // trace cursor of this role
#define DG_TRACE(role)                                                                          \
    do {                                                                                        \
        if (p.trace && tr < 64) p.trace[(blockIdx.x * 3 + (role)) * 64 + tr++] = clock64();     \
    } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_expect_tx(w_full, L::kWeights);
            for (int tap = 0; tap < 9; tap++)
                for (int h = 0; h < NH; h++)
                    tma_load_2d(w_s + (tap * NH + h) * L::kSlab, &tm_w, w_full, h * 64, tap * (kNSplit * BN) + nslice * BN);
            griddep_wait();   // activations of the previous launch must be complete before we read them
            DG_TRACE(0);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = first_tile; t < p.ntiles; t += tile_step) {
                for (int h = 0; h < NH; h++) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    DG_TRACE(0);
                    mbar_expect_tx(&a_full[stage], kWindowBytes);
                    tma_load_2d(a_s + stage * kStageBytes, &tm_act, &a_full[stage], h * 64,
                                DG_GUARD_ROWS + t * DG_TILE_M - DG_HALO_ROWS);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the loop (so every address stays in uniform registers); one elected
        // lane issues the MMAs and the commits.  Per MMA only the low descriptor words change, by
        // compile-time constants.
        mbar_wait(w_full, 0);
        if (lane == 0) DG_TRACE(1);
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        const uint32_t w_lo = umma_desc_lo(smem_u32(w_s));
        for (int t = first_tile; t < p.ntiles; t += tile_step) {
            mbar_wait(&acc_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
#pragma unroll
            for (int h = 0; h < NH; h++) {
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                if (lane == 0) DG_TRACE(1);
                const uint32_t a_lo = umma_desc_lo(smem_u32(a_s + stage * kStageBytes));
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const int row_off = DG_HALO_ROWS + (tap / 3 - 1) * DG_LINE_STRIDE + (tap % 3 - 1);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma_f16_ss_lo(d_tmem, a_lo + ((row_off * 128 + k * 32) >> 4),
                                           w_lo + (((tap * NH + h) * L::kSlab + k * 32) >> 4), kUmmaDescHiSw128, kIdesc,
                                           (h | tap | k) != 0);
                    }
                    umma_commit(&a_empty[stage]);      // window slot may be refilled once these MMAs retire
                    if (h == NH - 1) umma_commit(&acc_full[as]);   // accumulator complete -> epilogue
                }
                if (lane == 0) DG_TRACE(1);
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps
        griddep_wait();
        const int quarter = warp & 3;
        int as = 0;
        uint32_t aphase = 0;
        const float alpha = p.alpha;
        for (int t = first_tile; t < p.ntiles; t += tile_step) {
            const int m = t * DG_TILE_M + quarter * 32 + lane;
            const int q = m % DG_POS_ROWS;
            const bool halo = (m >= p.valid_rows) || (q % DG_LINE_STRIDE == DG_LINE_STRIDE - 1) ||
                              (q >= DG_POS_ROWS - DG_LINE_STRIDE);
            const size_t grow = static_cast<size_t>(DG_GUARD_ROWS + m);

            if (warp == 2 && lane == 0) DG_TRACE(2);
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
            if (warp == 2 && lane == 0) DG_TRACE(2);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            static_assert(BN == 16, "this kernel is instantiated for the 16-channel head convolution only");
            uint32_t acc[16];
            tmem_ld_32x32b_x16(taddr, acc);
            tmem_ld_wait();
            uint32_t packed[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float v0 = fmaf(alpha, __uint_as_float(acc[2 * j]), bias_s[2 * j]);
                float v1 = fmaf(alpha, __uint_as_float(acc[2 * j + 1]), bias_s[2 * j + 1]);
                v0 = (v0 > 0.f && !halo) ? v0 : 0.f;     // NaN-non-propagating ReLU; halo rows stay zero
                v1 = (v1 > 0.f && !halo) ? v1 : 0.f;
                const __half2 hv = __floats2half2_rn(v0, v1);
                packed[j] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            // channels 0..7 = policy samples -> out[row][8]; channels 8..9 = value samples -> out2[row][2]
            *reinterpret_cast<uint4*>(p.out + grow * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint32_t*>(p.out2 + grow * 2) = packed[4];
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    if (lane == 0 && (warp <= 2)) DG_TRACE(warp);
    tc_fence_before();
    __syncthreads();
    griddep_launch_dependents();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int NH, int BN>
static cudaError_t launch_one(const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p, int num_sms,
                              cudaStream_t stream, bool pdl) {
    using L = ConvSmem<NH, BN>;
    static std::atomic<unsigned long long> configured{0};   // per (NH, BN) instantiation, one bit per device
    auto kernel = conv3x3_tc_kernel<NH, BN>;
    if (cudaError_t e = opt_in_shared_memory(kernel, L::kTotal, configured); e != cudaSuccess) return e;
    constexpr int nsplit = (BN == 64) ? 2 : 1;
    int grid = num_sms - (num_sms % nsplit);
    const int work = p.ntiles * nsplit;
    if (grid > work) grid = work;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, tm_act, tm_w, p);
}

cudaError_t launch_conv_tc(ConvTcShape shape, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p,
                           int num_sms, cudaStream_t stream, bool pdl) {
    if (shape != ConvTcShape::kHeads) return cudaErrorInvalidValue;   // tower / up-sampling run in conv_tc2.cu
    return launch_one<2, 16>(tm_act, tm_w, p, num_sms, stream, pdl);
}

}  // namespace dg
