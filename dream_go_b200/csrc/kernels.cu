// Bandwidth-bound helper kernels around the tensor-core tower: feature packing and a slow direct
// convolution used by tests as an on-device cross-check of the tcgen05 kernels.
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

// Debug direct path only: [row][16] head-convolution output -> [row][8] policy samples + [row][2] value samples.
__global__ void split_heads_kernel(const __half* __restrict__ h, __half* __restrict__ pbuf, __half* __restrict__ vbuf, long rows) {
    const long r = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
    if (r >= rows) return;
    const long g = DG_GUARD_ROWS + r;
    for (int i = 0; i < 8; i++) pbuf[g * 8 + i] = h[g * 16 + i];
    vbuf[g * 2] = h[g * 16 + 8];
    vbuf[g * 2 + 1] = h[g * 16 + 9];
}
cudaError_t launch_split_heads(const __half* h, __half* pbuf, __half* vbuf, int batch, cudaStream_t s) {
    const long rows = static_cast<long>(dg_num_tiles(batch)) * DG_TILE_M;
    split_heads_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, s>>>(h, pbuf, vbuf, rows);
    return cudaGetLastError();
}

// Completion signal of a leaf batch: runs last on the batch's stream, after the device-to-host copies of the results,
// and writes `value` into a word of pinned host memory that the host polls (no driver call per poll).
__global__ void signal_host_kernel(volatile uint32_t* flag, uint32_t value) {
    __threadfence_system();
    *flag = value;
}
cudaError_t launch_signal_host(uint32_t* flag, uint32_t value, cudaStream_t s) {
    signal_host_kernel<<<1, 1, 0, s>>>(flag, value);
    return cudaGetLastError();
}

}  // namespace dg
