// Launchers of the NN kernels (csrc/nn/*.cu). Raw-pointer C++ API; see nn_bindings.cpp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hefl {
namespace nn {

// Fused Adam with Keras' legacy time-based decay (FLPyfhelin.py:140):
//   lr_t = lr * lr_scale / (1 + decay * (t-1)),  t = *step (already incremented)
// Updates p, m, v in place, zeroes g, optionally writes a bf16 shadow of p.
void adam_step(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n,
               const int64_t* step, const float* lr_scale, float lr, float decay, float beta1,
               float beta2, float eps, cudaStream_t st);

}  // namespace nn
}  // namespace hefl
