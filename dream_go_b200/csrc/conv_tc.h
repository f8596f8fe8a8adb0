#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>

namespace dg {

// cudaFuncSetAttribute applies to the CURRENT device only, and one process may hold an engine per device
// (the reference drives every GPU from one process, predictors/nn.rs:84-92): `done` keeps one bit per device.
template <typename Kernel>
inline cudaError_t opt_in_shared_memory(Kernel kernel, int bytes, std::atomic<unsigned long long>& done) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (device & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

struct ConvTcParams {
    int ntiles;                 // 128-row tiles to process
    int valid_rows;             // batch * 400; rows beyond are written as zero
    __half* out;                // board-row buffer (row 0 = guard start)
    int out_stride;             // elements per output row
    __half* out2;               // head convolution only: value samples [row][2]
    const __half* skip;         // optional residual input (same rows), nullptr if none
    int skip_stride;
    const float* bias;          // [Cout] fp32 (already scaled and fp16-rounded where the reference does so)
    float alpha;                // scale of the convolution result
    float beta;                 // scale of the skip input
    long long* trace;           // optional [grid][3 roles][64] clock64() samples (debug), nullptr normally
};

enum class ConvTcShape {
    kUp,      // 64 (32 real + 32 zero) -> 128 channels, two 64-channel CTAs per tile
    kTower,   // 128 -> 128 channels
    kHeads,   // 128 -> 16 channels (8 policy + 2 value + 6 zero)
};

cudaError_t launch_conv_tc(ConvTcShape shape, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p,
                           int num_sms, cudaStream_t stream, bool pdl);

// Tower (k_halves = 2) / up-sampling (k_halves = 1) convolution on CTA pairs (conv_tc2.cu).
cudaError_t launch_conv_pair(int k_halves, const CUtensorMap& tm_act, const CUtensorMap& tm_w, const ConvTcParams& p, int num_sms,
                             cudaStream_t stream, bool pdl);

// ---------------------------------------------------------------------------------------------
// Whole residual tower in ONE persistent launch (conv_tc2.cu: tower_kernel).
constexpr int kTowerMaxLayers = 41;          // up-sampling layer + 2 convolutions x 20 blocks

struct TowerLayer {
    int in_map;                 // activation tensor map this layer reads: 0 = features, 1 = x, 2 = y
    int nh;                     // 64-channel k-halves of the input (1 for the up-sampling layer, else 2)
    int has_skip;               // add beta * skip (the residual input) in the epilogue
    float alpha, beta;
    const float* bias;          // [128] fp32
    __half* out;                // output board-row buffer (row 0 = guard start)
    const __half* skip;         // residual input buffer or nullptr
    int wrows;                  // output channels of the layer: 128, or 16 = the head convolution (8 policy + 2 value samples)
    __half* out2;               // head convolution only: value samples [row][2] (`out` = policy samples [row][8])
};

struct TowerParams {
    CUtensorMap act[3];                     // 170-row x 64-channel load windows over features / x / y
    CUtensorMap w[kTowerMaxLayers];         // per-layer filter banks, [64 out x 64 in] boxes
    TowerLayer layer[kTowerMaxLayers];
    int nlayers;
    int ntiles;
    int valid_rows;
    int rot;                                // per-layer rotation of the unit -> CTA-pair assignment
    uint32_t gen;                           // flag base of this launch: done[tile] = gen + layers completed
    uint32_t* done;                         // [ntiles rounded up to even] progress flags
    long long* trace;
    int late_a;                             // debug (DG_FLAG_TOWER_LATE_A): request a layer's first windows after its filter slabs
};

cudaError_t launch_tower(const TowerParams& p, int num_sms, cudaStream_t stream);

}  // namespace dg
