// torch.ops.hefl.* bindings for the NN kernels.
#include <ATen/cuda/CUDAContext.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <cmath>

#include "nn.h"

namespace {

using at::Tensor;

void adam_step_(Tensor p, Tensor g, Tensor m, Tensor v, const c10::optional<Tensor>& shadow,
                const Tensor& step, const c10::optional<Tensor>& lr_scale, double lr, double decay,
                double beta1, double beta2, double eps) {
  TORCH_CHECK(p.scalar_type() == at::kFloat && p.is_contiguous(), "p must be contiguous float32");
  TORCH_CHECK(g.numel() == p.numel() && m.numel() == p.numel() && v.numel() == p.numel(), "size mismatch");
  TORCH_CHECK(step.scalar_type() == at::kLong && step.numel() == 1, "step must be a 1-element int64 tensor");
  const int64_t n = p.numel();
  if (p.is_cuda()) {
    void* sh = nullptr;
    if (shadow.has_value()) {
      TORCH_CHECK(shadow->scalar_type() == at::kBFloat16 && shadow->numel() >= n, "bad shadow");
      sh = shadow->data_ptr();
    }
    hefl::nn::adam_step(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(),
                        sh, n, step.data_ptr<int64_t>(),
                        lr_scale.has_value() ? lr_scale->data_ptr<float>() : nullptr, (float)lr,
                        (float)decay, (float)beta1, (float)beta2, (float)eps,
                        at::cuda::getCurrentCUDAStream().stream());
    return;
  }
  const float t = (float)step.item<int64_t>();
  const float scale = lr_scale.has_value() ? lr_scale->item<float>() : 1.0f;
  const float lr_t = (float)lr * scale / (1.0f + (float)decay * (t - 1.0f));
  const float alpha = lr_t * std::sqrt(1.0f - std::pow((float)beta2, t)) / (1.0f - std::pow((float)beta1, t));
  float* pp = p.data_ptr<float>();
  float* gp = g.data_ptr<float>();
  float* mp = m.data_ptr<float>();
  float* vp = v.data_ptr<float>();
  for (int64_t i = 0; i < n; ++i) {
    const float gi = gp[i];
    mp[i] = (float)beta1 * mp[i] + (1.0f - (float)beta1) * gi;
    vp[i] = (float)beta2 * vp[i] + (1.0f - (float)beta2) * gi * gi;
    pp[i] -= alpha * mp[i] / (std::sqrt(vp[i]) + (float)eps);
    gp[i] = 0.0f;
  }
  if (shadow.has_value()) shadow->copy_(p);
}

// Native data-loader primitive: copy rows `indices` of a pinned host tensor straight into a
// device batch buffer with one cudaMemcpyAsync per row on the current stream (no CPU gather).
void gather_h2d_(Tensor dst, const Tensor& src, const Tensor& indices) {
  TORCH_CHECK(dst.is_cuda() && src.is_cpu() && indices.is_cpu(), "dst on GPU, src/indices on CPU");
  TORCH_CHECK(dst.is_contiguous() && src.is_contiguous(), "contiguous tensors required");
  TORCH_CHECK(indices.scalar_type() == at::kLong, "indices must be int64");
  TORCH_CHECK(dst.scalar_type() == src.scalar_type(), "dtype mismatch");
  const int64_t rows = indices.numel();
  TORCH_CHECK(dst.size(0) >= rows, "dst too small");
  const size_t row_bytes = (size_t)(src.numel() / src.size(0)) * src.element_size();
  TORCH_CHECK((size_t)(dst.numel() / dst.size(0)) * dst.element_size() == row_bytes, "row size mismatch");
  const char* sp = static_cast<const char*>(src.data_ptr());
  char* dp = static_cast<char*>(dst.data_ptr());
  const int64_t* ix = indices.data_ptr<int64_t>();
  cudaStream_t st = at::cuda::getCurrentCUDAStream().stream();
  for (int64_t r = 0; r < rows; ++r) {
    TORCH_CHECK(ix[r] >= 0 && ix[r] < src.size(0), "index out of range");
    cudaMemcpyAsync(dp + (size_t)r * row_bytes, sp + (size_t)ix[r] * row_bytes, row_bytes,
                    cudaMemcpyHostToDevice, st);
  }
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(hefl, m) {
  m.def("adam_step_(Tensor(a!) p, Tensor(b!) g, Tensor(c!) m, Tensor(d!) v, Tensor? shadow, Tensor step, Tensor? lr_scale, float lr, float decay, float beta1, float beta2, float eps) -> ()", &adam_step_);
  m.def("gather_h2d_(Tensor(a!) dst, Tensor src, Tensor indices) -> ()", &gather_h2d_);
}
