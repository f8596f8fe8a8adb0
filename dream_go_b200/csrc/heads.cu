// Policy / value heads after the head convolution.
//
// Replaces (policy_head.rs:79-103)  softmax(tau * (flat(p1) @ Wl) + fp16(tau * bl))   -- dense.rs:197-220 + softmax.rs:58-79
// and      (value_head.rs:67-84)    tanh(flat(v1) @ wl + bl)                           -- dense.rs + activation.rs:60-83
//
// The policy fully-connected layer is a [batch x 3200] x [3200 x 384] GEMM on tcgen05: the head
// convolution leaves the 8 policy samples of every board ROW (halo rows included, as zeros) in a
// dense [row][8] buffer, so one position is a contiguous 3200-element K-major row and the FC weight
// is re-laid out once at load time to [384 out][3200 in] with zero columns for the halo rows.
// K is split over blockIdx.z; the fp32 partial sums are reduced IN A FIXED ORDER by the finishing
// kernel (no float atomics: results are bit-reproducible run to run), which also applies tau,
// the bias, the fp16 rounding of the dense output, the softmax, and the whole value head.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.h"
#include "kernels.h"
#include "layout.h"
#include "ptx.cuh"

namespace dg {

constexpr int kFcStages = kPolicyFcKBlocks / kPolicyFcSplit;     // k-blocks (of 64) per CTA, all in flight at once
constexpr int kFcStageBytes = 2 * 128 * 128;                      // A tile + B tile, [128][64] fp16 each
constexpr int kFcSmem = kFcStages * kFcStageBytes + 1024 + 1024;

__global__ void __launch_bounds__(128, 1)
policy_fc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ part, int m_total) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFcStages * kFcStageBytes);
    uint64_t* acc_bar = full + kFcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128, kb0 = blockIdx.z * kFcStages;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        for (int i = 0; i < kFcStages; i++) mbar_init(&full[i], 1);
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    griddep_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kFcStages; i++) {           // weights do not depend on the previous kernel
                mbar_expect_tx(&full[i], kFcStageBytes);
                tma_load_2d(smem + i * kFcStageBytes + 128 * 128, &tm_b, &full[i], (kb0 + i) * 64, n0);
            }
            griddep_wait();
            for (int i = 0; i < kFcStages; i++) tma_load_2d(smem + i * kFcStageBytes, &tm_a, &full[i], (kb0 + i) * 64, m0);
        }
        __syncwarp();
        for (int i = 0; i < kFcStages; i++) {
            mbar_wait(&full[i], 0);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = umma_desc_lo(smem_u32(smem + i * kFcStageBytes));
                const uint32_t b_lo = umma_desc_lo(smem_u32(smem + i * kFcStageBytes + 128 * 128));
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16_ss_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, kUmmaDescHiSw128, umma_idesc_f16(128, 128), (i | k) != 0);
                if (i == kFcStages - 1) umma_commit(acc_bar);
            }
            __syncwarp();
        }
    }
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    float* dst = part + (static_cast<size_t>(blockIdx.z) * m_total + m) * kPolicyFcN + n0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; j++) st_global_256(dst + c * 32 + j * 8, &acc[j * 8]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// One block per position: fixed-order reduction of the K-split partials, tau / bias / fp16 rounding,
// softmax (max-subtracted, "ACCURATE"), and the value head (722 -> 1 dot product + tanh).
__global__ void __launch_bounds__(384, 1)
heads_finish_kernel(const float* __restrict__ part, int m_total, const float* __restrict__ bp, float tau,
                    const __half* __restrict__ vbuf, const __half* __restrict__ wv, float bv, __half* __restrict__ policy,
                    __half* __restrict__ value) {
    __shared__ float red[12];
    __shared__ float red2[12];
    griddep_wait();
    const int n = blockIdx.x, o = threadIdx.x, warp = o >> 5, lane = o & 31;
    float logit = -INFINITY;
    if (o < 362) {
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < kPolicyFcSplit; s++) acc += part[(static_cast<size_t>(s) * m_total + n) * kPolicyFcN + o];
        // the dense layer's output tensor is fp16 (dense.rs:137-152): round before the softmax
        logit = __half2float(__float2half_rn(fmaf(tau, acc, bp[o])));
    }
    // value samples of this position: rows n*400 .. +399, two fp16 each; halo rows are zero
    float vacc = 0.f;
    for (int r = o; r < DG_POS_ROWS; r += 384) {
        const __half2 v2 = *reinterpret_cast<const __half2*>(vbuf + (static_cast<size_t>(DG_GUARD_ROWS) + n * DG_POS_ROWS + r) * 2);
        const __half2 w2 = *reinterpret_cast<const __half2*>(wv + r * 2);
        vacc = fmaf(__low2float(v2), __low2float(w2), vacc);
        vacc = fmaf(__high2float(v2), __high2float(w2), vacc);
    }
    float mx = logit;
    for (int d = 16; d > 0; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        vacc += __shfl_xor_sync(0xffffffffu, vacc, d);
    }
    if (lane == 0) { red[warp] = mx; red2[warp] = vacc; }
    __syncthreads();
    mx = red[0];
    float vsum = red2[0];
    for (int i = 1; i < 12; i++) { mx = fmaxf(mx, red[i]); vsum += red2[i]; }
    __syncthreads();
    const float ex = (o < 362) ? expf(logit - mx) : 0.f;
    float sum = ex;
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = red[0];
    for (int i = 1; i < 12; i++) sum += red[i];
    if (o < 362) policy[static_cast<size_t>(n) * 362 + o] = __float2half_rn(ex / sum);
    if (o == 0) {
        const float pre = __half2float(__float2half_rn(vsum + bv));     // dense output is fp16, tanh applied in place
        value[n] = __float2half_rn(tanhf(pre));
    }
}

static cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, func, args);
}

cudaError_t launch_policy_fc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, float* part, int batch, cudaStream_t s) {
    static std::atomic<unsigned long long> configured{0};
    if (cudaError_t e = opt_in_shared_memory(policy_fc_kernel, kFcSmem, configured); e != cudaSuccess) return e;
    const int m_tiles = (batch + 127) / 128;
    int m_total = m_tiles * 128;
    void* args[] = {const_cast<CUtensorMap*>(&tm_a), const_cast<CUtensorMap*>(&tm_b), &part, &m_total};
    return launch_pdl(reinterpret_cast<const void*>(policy_fc_kernel), dim3(kPolicyFcN / 128, m_tiles, kPolicyFcSplit), dim3(128), kFcSmem, s, args);
}

cudaError_t launch_heads_finish(const float* part, int batch, const float* bp, float tau, const __half* vbuf, const __half* wv,
                                float bv, __half* policy, __half* value, cudaStream_t s) {
    int m_total = ((batch + 127) / 128) * 128;
    void* args[] = {&part, &m_total, &bp, &tau, &vbuf, &wv, &bv, &policy, &value};
    return launch_pdl(reinterpret_cast<const void*>(heads_finish_kernel), dim3(batch), dim3(384), 0, s, args);
}

}  // namespace dg
