// Small fused NN kernels for sm_100a: Adam (+ bf16 weight shadow), input preprocessing,
// un-pooling of gradients. SURVEY.md §2.4 K17.
#include <cuda_bf16.h>

#include "../he/kernels.h"
#include "../he/philox.h"
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

// ------------------------------------------------------------------------------------------
// Data-movement kernels around the tcgen05 convolutions
// ------------------------------------------------------------------------------------------
namespace hefl {
namespace nn {

// uint8 NHWC (3 channels) -> bf16 [P,16] (channels 3..15 zero), with 1/255 rescale and an optional per-sample affine warp (Keras shear/zoom/flip,
// FLPyfhelin.py:80-86) sampled bilinearly with border clamp (torch.grid_sample semantics,
// align_corners=False).
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ x, const float* __restrict__ theta,
                                     __nv_bfloat16* __restrict__ X, int B, int H, int W, uint64_t aug_seed,
                                     const int64_t* __restrict__ step) {
  const int b = blockIdx.y;
  const int HW = H * W;
  float tl[6];
  bool warp_img = theta != nullptr;
  if (theta) {
#pragma unroll
    for (int i = 0; i < 6; ++i) tl[i] = theta[b * 6 + i];
  } else if (aug_seed != 0) {
    // Keras ImageDataGenerator(shear_range=0.2 deg, zoom_range=0.2, horizontal_flip) drawn per sample
    // from a counter-based generator keyed by (seed, optimiser step, sample).
    const uint64_t st = step ? (uint64_t)*step : 0ull;
    const Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)st, (uint32_t)(st >> 32), 7u, (uint32_t)aug_seed,
                                    (uint32_t)(aug_seed >> 32));
    const float u0 = r.x * 2.3283064e-10f, u1 = r.y * 2.3283064e-10f, u2 = r.z * 2.3283064e-10f;
    const float sh = (u0 * 2.f - 1.f) * 0.2f * 0.017453292f;
    const float zx = 1.f + (u1 * 2.f - 1.f) * 0.2f, zy = 1.f + (u2 * 2.f - 1.f) * 0.2f;
    const float flip = (r.w & 1u) ? -1.f : 1.f;
    tl[0] = zx * flip; tl[1] = -sinf(sh) * zx; tl[2] = 0.f;
    tl[3] = 0.f; tl[4] = cosf(sh) * zy; tl[5] = 0.f;
    warp_img = true;
  }
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int w = p % W;
    const int h = p / W;
    const int64_t m = (int64_t)b * HW + p;
    float c[3];
    if (warp_img) {
      const float* t = tl;
      const float xn = (2.f * w + 1.f) / W - 1.f, yn = (2.f * h + 1.f) / H - 1.f;
      const float sx = t[0] * xn + t[1] * yn + t[2], sy = t[3] * xn + t[4] * yn + t[5];
      float ix = ((sx + 1.f) * W - 1.f) * 0.5f, iy = ((sy + 1.f) * H - 1.f) * 0.5f;
      ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
      iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
      const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
      const float fx = ix - x0, fy = iy - y0;
      const uint8_t* img = x + (int64_t)b * H * W * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float v00 = img[((int64_t)y0 * W + x0) * 3 + k], v01 = img[((int64_t)y0 * W + x1) * 3 + k];
        const float v10 = img[((int64_t)y1 * W + x0) * 3 + k], v11 = img[((int64_t)y1 * W + x1) * 3 + k];
        c[k] = ((v00 * (1.f - fx) + v01 * fx) * (1.f - fy) + (v10 * (1.f - fx) + v11 * fx) * fy) * (1.f / 255.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = (float)x[m * 3 + k] * (1.f / 255.f);
    }
    const uint32_t b0 = __bfloat16_as_ushort(__float2bfloat16(c[0]));
    const uint32_t b1 = __bfloat16_as_ushort(__float2bfloat16(c[1]));
    const uint32_t b2 = __bfloat16_as_ushort(__float2bfloat16(c[2]));
    uint4* dst = reinterpret_cast<uint4*>(X + m * 16);
    dst[0] = make_uint4(b0 | (b1 << 16), b2, 0u, 0u);
    dst[1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

void preprocess_u8(const uint8_t* x, const float* theta, void* X, int B, int H, int W, uint64_t aug_seed,
                   const int64_t* step, cudaStream_t st) {
  int bx = (H * W + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, B);
  preprocess_u8_kernel<<<grid, 256, 0, st>>>(x, theta, reinterpret_cast<__nv_bfloat16*>(X), B, H, W, aug_seed, step);
  hefl::cuda::note_launch();
}

// Gradient of (ReLU -> 2x2 max-pool): scatter the pooled gradient to the arg-max position on the
// conv-input grid, zero elsewhere (including the invalid border). Output [P, Co] bf16 feeds both
// dgrad (K-major A operand) and wgrad (MN-major B operand). One thread = one pixel x 8 channels.
__global__ void unpool_relu_kernel(const __nv_bfloat16* __restrict__ g, const uint8_t* __restrict__ amax,
                                   const __nv_bfloat16* __restrict__ ypool, __nv_bfloat16* __restrict__ dY, int B,
                                   int H, int W, int Hp, int Wp, int Co) {
  const int groups = Co >> 3;
  const int row = blockIdx.y;                 // b*H + h
  const int b = row / H, h = row - b * H;
  const int hp = h >> 1;
  const int per_row = W * groups;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_row; t += gridDim.x * blockDim.x) {
    const int w = t / groups;
    const int cg = t - w * groups;
    const int wp = w >> 1;
    uint4 outv = make_uint4(0u, 0u, 0u, 0u);
    if (hp < Hp && wp < Wp) {
      const int64_t o = (((int64_t)b * Hp + hp) * Wp + wp) * Co + cg * 8;
      const uint4 gv = *reinterpret_cast<const uint4*>(g + o);
      const uint4 yv = *reinterpret_cast<const uint4*>(ypool + o);
      const uint2 av = *reinterpret_cast<const uint2*>(amax + o);
      const uint32_t pos = (uint32_t)((h & 1) * 2 + (w & 1));
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
      const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
      const uint32_t aw[2] = {av.x, av.y};
      uint32_t ow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = i * 2 + j;
          const uint32_t a8 = (aw[c >> 2] >> ((c & 3) * 8)) & 0xFFu;
          const uint32_t yb = (yw[i] >> (16 * j)) & 0xFFFFu;
          const bool on = a8 == pos && yb != 0u && (yb & 0x8000u) == 0u;   // pooled > 0
          if (on) r |= ((gw[i] >> (16 * j)) & 0xFFFFu) << (16 * j);
        }
        ow[i] = r;
      }
      outv = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    *reinterpret_cast<uint4*>(dY + ((int64_t)row * W + w) * Co + cg * 8) = outv;
  }
}

void unpool_relu(const void* g, const uint8_t* amax, const void* ypool, void* dY, int B, int H, int W, int Hp,
                 int Wp, int Co, cudaStream_t st) {
  const int per_row = W * (Co / 8);
  dim3 grid((per_row + 255) / 256, B * H);
  unpool_relu_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), amax,
                                             reinterpret_cast<const __nv_bfloat16*>(ypool),
                                             reinterpret_cast<__nv_bfloat16*>(dY), B, H, W, Hp, Wp, Co);
  hefl::cuda::note_launch();
}

// bf16 shadow of the flat parameters ([Co][Ci][3][3] per conv) -> Wf [tap][Co][CK] (forward B
// operand, channel-padded) and Wd [tap][Ci][Co] (dgrad B operand).
__global__ void conv_weight_relayout_kernel(const __nv_bfloat16* __restrict__ shadow, const ConvLayerTable t,
                                            __nv_bfloat16* __restrict__ Wf, __nv_bfloat16* __restrict__ Wd) {
  const int l = blockIdx.y;
  const int Ci = t.Ci[l], CK = t.CK[l], Co = t.Co[l];
  const __nv_bfloat16* w = shadow + t.w_off[l];
  const int nf = 9 * Co * CK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
    const int ck = i % CK, co = (i / CK) % Co, tap = i / (CK * Co);
    Wf[t.wf_off[l] + i] = ck < Ci ? w[(co * Ci + ck) * 9 + tap] : __float2bfloat16(0.f);
  }
  if (l > 0) {
    const int nd = 9 * Ci * Co;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
      const int co = i % Co, ci = (i / Co) % Ci, tap = i / (Co * Ci);
      Wd[t.wd_off[l] + i] = w[(co * Ci + ci) * 9 + tap];
    }
  }
}

void conv_weight_relayout(const void* shadow, const ConvLayerTable& t, void* Wf, void* Wd, cudaStream_t st) {
  dim3 grid(32, t.n);
  conv_weight_relayout_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(shadow), t,
                                                    reinterpret_cast<__nv_bfloat16*>(Wf),
                                                    reinterpret_cast<__nv_bfloat16*>(Wd));
  hefl::cuda::note_launch();
}

// dW32 [9*CK+1][Co] (wgrad output, row 9*CK = bias gradient) -> flat fp32 gradient in the
// parameter layout ([Co][Ci][3][3], then bias); clears dW32 for the next step.
__global__ void conv_grad_finalize_kernel(float* __restrict__ dW32, const ConvLayerTable t,
                                          float* __restrict__ grad) {
  const int l = blockIdx.y;
  const int Ci = t.Ci[l], CK = t.CK[l], Co = t.Co[l];
  float* src = dW32 + t.dw_off[l];
  const int nw = Co * Ci * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
    grad[t.w_off[l] + i] = src[(size_t)(tap * CK + ci) * Co + co];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Co; i += gridDim.x * blockDim.x)
    grad[t.b_off[l] + i] = src[(size_t)9 * CK * Co + i];
}
__global__ void zero_f32_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

void conv_grad_finalize(float* dW32, const ConvLayerTable& t, float* grad, cudaStream_t st) {
  dim3 grid(32, t.n);
  conv_grad_finalize_kernel<<<grid, 256, 0, st>>>(dW32, t, grad);
  const int l = t.n - 1;
  const int64_t total = t.dw_off[l] + (int64_t)(9 * t.CK[l] + 1) * t.Co[l];
  zero_f32_kernel<<<64, 256, 0, st>>>(dW32, total);
  hefl::cuda::note_launch(2);
}

}  // namespace nn
}  // namespace hefl
