// Bandwidth-bound helper kernels around the tensor-core tower: feature packing, the head
// fully-connected layers (+softmax / tanh), and a slow direct convolution used by tests as an
// on-device cross-check of the tcgen05 kernel.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "layout.h"

namespace dg {

// ---------------------------------------------------------------------------------------------
// features NHWC fp16 [B][361][32]  ->  board rows [.][64] fp16 (channels 32..63 and halo rows = 0)
// Replaces the H2D-then-conv input of graph.rs:130-131.  One thread per 16-byte chunk of a row:
// reads are 64 contiguous bytes per point, writes are fully coalesced 128-byte rows.
__global__ void pack_features_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int total_rows, int valid_rows) {
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const long m = idx >> 3;
    const int chunk = idx & 7;
    if (m >= total_rows) return;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int q = static_cast<int>(m % DG_POS_ROWS);
    const int x = q % DG_LINE_STRIDE, y = q / DG_LINE_STRIDE;
    if (chunk < 4 && m < valid_rows && x < 19 && y < 19) {
        const long n = m / DG_POS_ROWS;
        v = __ldg(in + (n * 361 + y * 19 + x) * 4 + chunk);
    }
    out[(DG_GUARD_ROWS + m) * 8 + chunk] = v;
}

// compact positions (361 plane masks + k) -> board rows [.][64] fp16.
// `pos` may live in mapped pinned host memory (the leaf queue) or in device memory.
__global__ void pack_compact_kernel(const uint32_t* __restrict__ pos, uint4* __restrict__ out, int total_rows, int valid_rows) {
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const long m = idx >> 3;
    const int chunk = idx & 7;
    if (m >= total_rows) return;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int q = static_cast<int>(m % DG_POS_ROWS);
    const int x = q % DG_LINE_STRIDE, y = q / DG_LINE_STRIDE;
    if (chunk < 4 && m < valid_rows && x < 19 && y < 19) {
        const long n = m / DG_POS_ROWS;
        const uint32_t* p = pos + n * 362;                 // sizeof(dg_packed_position) / 4
        const uint32_t mask = p[y * 19 + x] >> (chunk * 8);
        const uint32_t kbits = p[361] & 0xffffu;
        uint32_t h[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t one = (chunk == 0 && j < 2) ? kbits : 0x3c00u;
            h[j] = ((mask >> j) & 1u) ? one : 0u;
        }
        v = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
    out[(DG_GUARD_ROWS + m) * 8 + chunk] = v;
}

cudaError_t launch_pack_features(const void* in, void* out, int batch, cudaStream_t s) {
    const int total = dg_num_tiles(batch) * DG_TILE_M;
    const long threads = static_cast<long>(total) * 8;
    pack_features_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), total, batch * DG_POS_ROWS);
    return cudaGetLastError();
}
cudaError_t launch_pack_compact(const void* pos, void* out, int batch, cudaStream_t s) {
    const int total = dg_num_tiles(batch) * DG_TILE_M;
    const long threads = static_cast<long>(total) * 8;
    pack_compact_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(
        static_cast<const uint32_t*>(pos), static_cast<uint4*>(out), total, batch * DG_POS_ROWS);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Direct 3x3 convolution, one thread per (row, output channel).  TEST CROSS-CHECK ONLY
// (DG_FLAG_DEBUG_DIRECT_CONV); same operands and epilogue as conv3x3_tc_kernel.
__global__ void conv3x3_direct_kernel(const __half* __restrict__ in, int cin, const __half* __restrict__ w, int ntot,
                                      const float* __restrict__ bias, float alpha, float beta, const __half* skip,
                                      int skip_stride, __half* out, int out_stride, int total_rows, int valid_rows) {
    const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    const int n = static_cast<int>(idx % ntot);
    const long m = idx / ntot;
    if (m >= total_rows) return;
    const int q = static_cast<int>(m % DG_POS_ROWS);
    const bool halo = (m >= valid_rows) || (q % DG_LINE_STRIDE == DG_LINE_STRIDE - 1) || (q >= DG_POS_ROWS - DG_LINE_STRIDE);
    float acc = 0.f;
    for (int tap = 0; tap < 9; tap++) {
        const long row = DG_GUARD_ROWS + m + (tap / 3 - 1) * DG_LINE_STRIDE + (tap % 3 - 1);
        const uint4* a = reinterpret_cast<const uint4*>(in + row * cin);
        const uint4* b = reinterpret_cast<const uint4*>(w + (static_cast<long>(tap) * ntot + n) * cin);
        for (int c = 0; c < cin / 8; c++) {
            const uint4 av = a[c], bv = b[c];
            const __half2* ah = reinterpret_cast<const __half2*>(&av);
            const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 af = __half22float2(ah[j]), bf = __half22float2(bh[j]);
                acc = fmaf(af.x, bf.x, acc);
                acc = fmaf(af.y, bf.y, acc);
            }
        }
    }
    const long grow = DG_GUARD_ROWS + m;
    float v = fmaf(alpha, acc, bias[n]);
    if (skip) v = fmaf(beta, __half2float(skip[grow * skip_stride + n]), v);
    v = (v > 0.f && !halo) ? v : 0.f;
    out[grow * out_stride + n] = __float2half_rn(v);
}

cudaError_t launch_conv_direct(const __half* in, int cin, const __half* w, int ntot, const float* bias, float alpha, float beta,
                               const __half* skip, int skip_stride, __half* out, int out_stride, int batch, cudaStream_t s) {
    const int total = dg_num_tiles(batch) * DG_TILE_M;
    const long threads = static_cast<long>(total) * ntot;
    conv3x3_direct_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(
        in, cin, w, ntot, bias, alpha, beta, skip, skip_stride, out, out_stride, total, batch * DG_POS_ROWS);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Heads: policy FC 2888->362 (+tau, softmax) and value FC 722->1 (+tanh) over the 16-channel head
// convolution output (channels 0..7 policy samples, 8..9 value samples).
// Replaces dense.rs:197-220 + softmax.rs:58-79 (policy_head.rs:79-103) and dense + tanh
// (value_head.rs:67-84).  One block evaluates kHeadPos positions; thread (ks, op) accumulates
// outputs 2*op, 2*op+1 over a quarter of the 2888 inputs, weights read as coalesced half2.
constexpr int kHeadPos = 2;
constexpr int kHeadKSplit = 4;
constexpr int kHeadThreads = 768;     // >= 181 * 4

__global__ void __launch_bounds__(kHeadThreads, 1)
heads_fc_kernel(const __half* __restrict__ hbuf, const __half2* __restrict__ wp, const float* __restrict__ bp, float tau,
                const __half* __restrict__ wv, float bv, int batch, __half* __restrict__ policy, __half* __restrict__ value) {
    __shared__ __half2 xs[2888];                       // [i] -> (position 0, position 1)
    __shared__ float vsum[kHeadPos][32];
    __shared__ float part[kHeadKSplit][kHeadPos][364];
    __shared__ float logits[kHeadPos][364];
    const int n0 = blockIdx.x * kHeadPos;
    const int tid = threadIdx.x;

    // stage the policy samples and reduce the value dot product on the way
    float vacc[kHeadPos] = {0.f, 0.f};
    for (int h = tid; h < 361; h += kHeadThreads) {
        __half pv[kHeadPos][16];
#pragma unroll
        for (int p = 0; p < kHeadPos; p++) {
            const int n = n0 + p;
            if (n < batch) {
                const long row = DG_GUARD_ROWS + static_cast<long>(n) * DG_POS_ROWS + (h / 19) * DG_LINE_STRIDE + h % 19;
                const uint4* src = reinterpret_cast<const uint4*>(hbuf + row * 16);
                *reinterpret_cast<uint4*>(&pv[p][0]) = src[0];
                *reinterpret_cast<uint4*>(&pv[p][8]) = src[1];
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) pv[p][j] = __float2half(0.f);
            }
            vacc[p] += __half2float(pv[p][8]) * __half2float(wv[2 * h]) + __half2float(pv[p][9]) * __half2float(wv[2 * h + 1]);
        }
#pragma unroll
        for (int s = 0; s < 8; s++) xs[8 * h + s] = __halves2half2(pv[0][s], pv[1][s]);
    }
#pragma unroll
    for (int p = 0; p < kHeadPos; p++) {
        float v = vacc[p];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) vsum[p][tid >> 5] = v;
    }
    __syncthreads();

    const int ks = tid / 181, op = tid % 181;
    if (ks < kHeadKSplit) {
        float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;   // [position][output parity]
        const int i0 = ks * 722;
#pragma unroll 4
        for (int i = i0; i < i0 + 722; i++) {
            const float2 w = __half22float2(wp[i * 181 + op]);
            const float2 x = __half22float2(xs[i]);
            a00 = fmaf(x.x, w.x, a00);
            a01 = fmaf(x.x, w.y, a01);
            a10 = fmaf(x.y, w.x, a10);
            a11 = fmaf(x.y, w.y, a11);
        }
        part[ks][0][2 * op] = a00;
        part[ks][0][2 * op + 1] = a01;
        part[ks][1][2 * op] = a10;
        part[ks][1][2 * op + 1] = a11;
    }
    __syncthreads();
    for (int i = tid; i < kHeadPos * 362; i += kHeadThreads) {
        const int p = i / 362, o = i % 362;
        const float acc = (part[0][p][o] + part[1][p][o]) + (part[2][p][o] + part[3][p][o]);
        // the dense layer's output tensor is fp16 (dense.rs:137-152): round before the softmax
        logits[p][o] = __half2float(__float2half_rn(fmaf(tau, acc, bp[o])));
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    if (warp < kHeadPos && n0 + warp < batch) {
        const int p = warp;
        float mx = -INFINITY;
        for (int o = lane; o < 362; o += 32) mx = fmaxf(mx, logits[p][o]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int o = lane; o < 362; o += 32) sum += expf(logits[p][o] - mx);
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        __half* dst = policy + static_cast<long>(n0 + p) * 362;
        for (int o = lane; o < 362; o += 32) dst[o] = __float2half_rn(expf(logits[p][o] - mx) * inv);
    } else if (warp >= 8 && warp < 8 + kHeadPos && n0 + (warp - 8) < batch) {
        const int p = warp - 8;
        float v = (lane < kHeadThreads / 32) ? vsum[p][lane] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            const float pre = __half2float(__float2half_rn(v + bv));   // dense output is fp16, then tanh in place
            value[n0 + p] = __float2half_rn(tanhf(pre));
        }
    }
}

cudaError_t launch_heads_fc(const __half* hbuf, const __half* wp, const float* bp, float tau, const __half* wv, float bv,
                            int batch, __half* policy, __half* value, cudaStream_t s) {
    heads_fc_kernel<<<(batch + kHeadPos - 1) / kHeadPos, kHeadThreads, 0, s>>>(
        hbuf, reinterpret_cast<const __half2*>(wp), bp, tau, wv, bv, batch, policy, value);
    return cudaGetLastError();
}

}  // namespace dg
