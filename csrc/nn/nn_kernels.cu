// Small fused NN kernels for sm_100a: Adam (+ bf16 weight shadow), input preprocessing,
// un-pooling of gradients. SURVEY.md §2.4 K17.
#include <cuda_bf16.h>

#include "../he/kernels.h"
#include "nn.h"

namespace hefl {
namespace nn {

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n,
                            const int64_t* __restrict__ step, const float* __restrict__ lr_scale,
                            float lr, float decay, float beta1, float beta2, float eps) {
  const float t = (float)(*step);
  const float lr_t = lr * (lr_scale ? *lr_scale : 1.0f) / (1.0f + decay * (t - 1.0f));
  const float bc1 = 1.0f - powf(beta1, t);
  const float bc2 = 1.0f - powf(beta2, t);
  const float alpha = lr_t * sqrtf(bc2) / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float pi = p[i] - alpha * mi / (sqrtf(vi) + eps);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    g[i] = 0.0f;
    if (shadow) shadow[i] = __float2bfloat16(pi);
  }
}

void adam_step(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n,
               const int64_t* step, const float* lr_scale, float lr, float decay, float beta1,
               float beta2, float eps, cudaStream_t st) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n,
                                      step, lr_scale, lr, decay, beta1, beta2, eps);
  hefl::cuda::note_launch();
}

}  // namespace nn
}  // namespace hefl
