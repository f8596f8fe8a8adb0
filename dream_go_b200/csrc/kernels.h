#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace dg {

constexpr int DG_POS_ROWS_ = 400;   // == DG_POS_ROWS (layout.h)

cudaError_t launch_pack_features(const void* in_nhwc_f16, void* out_rows64, int batch, cudaStream_t s);
cudaError_t launch_pack_compact(const void* positions, void* out_rows64, int batch, cudaStream_t s);
// features.cu -- V1 feature planes + legal mask from raw stones (dg_raw_position -> dg_packed_position)
cudaError_t upload_feature_tables(const unsigned long long* zobrist, const uint16_t* symmetry);
cudaError_t launch_planes_from_stones(const void* raw, void* planes, void* legal, void* rows64, void* cand, void* rep, int batch,
                                      cudaStream_t s);
cudaError_t launch_prior_from_policy(const void* raw, const void* policy, const void* cand, const void* rep, float* prior, int batch,
                                     cudaStream_t s);
cudaError_t launch_conv_direct(const __half* in, int cin, const __half* w, int ntot, const float* bias, float alpha, float beta,
                               const __half* skip, int skip_stride, __half* out, int out_stride, int batch, cudaStream_t s);
cudaError_t launch_signal_host(uint32_t* flag_in_pinned_host_memory, uint32_t value, cudaStream_t s);
cudaError_t launch_split_heads(const __half* h, __half* pbuf, __half* vbuf, int batch, cudaStream_t s);
// heads.cu -- policy FC as a K-split tcgen05 GEMM + finishing kernel (softmax, value head)
constexpr int kPolicyFcK = DG_POS_ROWS_ * 8;        // 3200: 8 policy samples of each of the 400 board rows of a position
constexpr int kPolicyFcKBlocks = kPolicyFcK / 64;   // 50
constexpr int kPolicyFcSplit = 10;                  // K split (blockIdx.z)
constexpr int kPolicyFcN = 384;                     // 362 outputs padded to 3 x 128
cudaError_t launch_policy_fc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, float* part, int batch, cudaStream_t s);
cudaError_t launch_heads_finish(const float* part, int batch, const float* bp, float tau, const __half* vbuf, const __half* wv,
                                float bv, __half* policy, __half* value, cudaStream_t s);

}  // namespace dg
