// Cooperative fused layer kernels of the reference ConvNet (one CTA per image; see fused_convnet.cu).
#pragma once
#include <cuda_runtime.h>

#include "ops_kernels.h"

namespace pdt {

// Two words of device memory (zero before first use) shared by every cooperative kernel of a device.
struct GridSync {
  unsigned int* count;
  unsigned int* gen;
};

// One CTA per image, all co-resident: the batch must not exceed the number of SMs.
bool fused_convnet_supported(int B);

// x [B,28,28] → y [B,28,28,16] (conv1 + bias, kept for backward), out [B,14,14,16] (BN + ReLU + pool), saved [32] = mean, invstd.
// partials: B·32 floats of scratch.
void launch_convnet_l1_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta, float* y, float* out,
                           float* saved, float* running_mean, float* running_var, long long* nbt, float momentum, float eps, int B,
                           float* partials, GridSync gs, cudaStream_t st);
// dp [B,14,14,16] → dgamma/dbeta [16], dw [16,1,5,5], db [16].  partials: B·32, partials_w: B·416 floats.
void launch_convnet_l1_bwd(const float* dp, const float* y, const float* x, const float* saved, const float* gamma, const float* beta,
                           float* dgamma, float* dbeta, float* dw, float* db, int B, float* partials, float* partials_w, GridSync gs,
                           cudaStream_t st);
// x [B,14,14,16] NHWC → y [B,14,14,32], out [B,32,7,7] NCHW, saved [64]; logits [B,ncls] = fc(out) when logits != nullptr.
// partials: B·64 floats.
void launch_convnet_l2_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta, float* y, float* out,
                           float* saved, float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                           const float* fcw, const float* fcb, float* logits, int ncls, int B, float* partials, GridSync gs, cudaStream_t st);
// dout [B,32,7,7] → dgamma/dbeta [32], dy [B,14,14,32] (gradient at the conv2 output), dx [B,14,14,16] (data gradient).
void launch_convnet_l2_bwd(const float* dout, const float* y, const float* saved, const float* gamma, const float* beta, const float* w,
                           float* dgamma, float* dbeta, float* dy, float* dx, int B, float* partials, GridSync gs, cudaStream_t st);

}  // namespace pdt
