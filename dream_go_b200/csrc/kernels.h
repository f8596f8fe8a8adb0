#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace dg {

cudaError_t launch_pack_features(const void* in_nhwc_f16, void* out_rows64, int batch, cudaStream_t s);
cudaError_t launch_pack_compact(const void* positions, void* out_rows64, int batch, cudaStream_t s);
cudaError_t launch_conv_direct(const __half* in, int cin, const __half* w, int ntot, const float* bias, float alpha, float beta,
                               const __half* skip, int skip_stride, __half* out, int out_stride, int batch, cudaStream_t s);
cudaError_t launch_heads_fc(const __half* hbuf, const __half* wp, const float* bp, float tau, const __half* wv, float bv,
                            int batch, __half* policy, __half* value, cudaStream_t s);

}  // namespace dg
