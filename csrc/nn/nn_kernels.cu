// Small fused NN kernels for sm_100a: Adam (+ bf16 weight shadow), input preprocessing,
// un-pooling of gradients. SURVEY.md §2.4 K17.
#include <cuda_bf16.h>

#include "../he/kernels.h"
#include "../he/philox.h"
#include "nn.h"
#include "launch.cuh"

namespace hefl {
namespace nn {

int g_pdl = 1;
void set_pdl(int on) { g_pdl = on; }

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n,
                            const int64_t* __restrict__ step, const float* __restrict__ lr_scale,
                            float lr, float decay, float beta1, float beta2, float eps) {
  pdl_wait();   // parameter writer: no early trigger (see launch.cuh)
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
  launch_pdl(adam_kernel, dim3(blocks), dim3(256), 0, st, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n,
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
                                     const int64_t* __restrict__ step, int spack) {
  pdl_prologue();
  const int b = blockIdx.y;
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
  // One warp = one run of consecutive pixels of an image row. With `spack` (s-packed layer-1 input: channels
  // 0-2 = this pixel, 3-5 = pixel w+1, 6-8 = pixel w+2, so the three horizontal filter taps become part of the
  // GEMM K dimension) a warp stores 30 pixels and lanes 30/31 only feed their values to the shuffles.
  const int lane = threadIdx.x & 31;
  const int PW = spack ? 30 : 32;
  const int chunks = (W + PW - 1) / PW;
  const int tasks = H * chunks;
  const int wpb = blockDim.x >> 5;
  const uint8_t* img = x + (int64_t)b * H * W * 3;
  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < tasks; task += gridDim.x * wpb) {
    const int h = task / chunks;
    const int w = (task - h * chunks) * PW + lane;
    const int wc = w < W ? w : W - 1;
    float c[3];
    if (warp_img) {
      const float* t = tl;
      const float xn = (2.f * wc + 1.f) / W - 1.f, yn = (2.f * h + 1.f) / H - 1.f;
      const float sx = t[0] * xn + t[1] * yn + t[2], sy = t[3] * xn + t[4] * yn + t[5];
      float ix = ((sx + 1.f) * W - 1.f) * 0.5f, iy = ((sy + 1.f) * H - 1.f) * 0.5f;
      ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
      iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
      const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
      const float fx = ix - x0, fy = iy - y0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float v00 = img[((int64_t)y0 * W + x0) * 3 + k], v01 = img[((int64_t)y0 * W + x1) * 3 + k];
        const float v10 = img[((int64_t)y1 * W + x0) * 3 + k], v11 = img[((int64_t)y1 * W + x1) * 3 + k];
        c[k] = ((v00 * (1.f - fx) + v01 * fx) * (1.f - fy) + (v10 * (1.f - fx) + v11 * fx) * fy) * (1.f / 255.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = (float)img[((int64_t)h * W + wc) * 3 + k] * (1.f / 255.f);
    }
    uint32_t p01 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(c[0])) |
                   ((uint32_t)__bfloat16_as_ushort(__float2bfloat16(c[1])) << 16);
    uint32_t p2 = __bfloat16_as_ushort(__float2bfloat16(c[2]));
    if (w >= W) { p01 = 0u; p2 = 0u; }                       // past the row end: zero neighbours
    const uint32_t n1_01 = __shfl_down_sync(0xffffffffu, p01, 1), n1_2 = __shfl_down_sync(0xffffffffu, p2, 1);
    const uint32_t n2_01 = __shfl_down_sync(0xffffffffu, p01, 2), n2_2 = __shfl_down_sync(0xffffffffu, p2, 2);
    if (lane < PW && w < W) {
      uint4* dst = reinterpret_cast<uint4*>(X + ((int64_t)(b * H + h) * W + w) * 16);
      if (spack) {
        dst[0] = make_uint4(p01, p2 | (n1_01 << 16), (n1_01 >> 16) | (n1_2 << 16), n2_01);
        dst[1] = make_uint4(n2_2, 0u, 0u, 0u);
      } else {
        dst[0] = make_uint4(p01, p2, 0u, 0u);
        dst[1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

void preprocess_u8(const uint8_t* x, const float* theta, void* X, int B, int H, int W, uint64_t aug_seed,
                   const int64_t* step, int spack, cudaStream_t st) {
  const int tasks = H * ((W + (spack ? 29 : 31)) / (spack ? 30 : 32));
  int bx = (tasks + 7) / 8;
  if (bx > 64) bx = 64;
  dim3 grid(bx, B);
  launch_pdl(preprocess_u8_kernel, dim3(grid), dim3(256), 0, st, x, theta, reinterpret_cast<__nv_bfloat16*>(X), B, H, W,
             aug_seed, step, spack);
  hefl::cuda::note_launch();
}

// Gradient of (ReLU -> 2x2 max-pool): scatter the pooled gradient to the arg-max position on the
// conv-input grid, zero elsewhere (including the invalid border). Output [P, Co] bf16 feeds both
// dgrad (K-major A operand) and wgrad (MN-major B operand). One thread = one pixel x 8 channels.
__global__ void unpool_relu_kernel(const __nv_bfloat16* __restrict__ g, const uint8_t* __restrict__ amax,
                                   const __nv_bfloat16* __restrict__ ypool, __nv_bfloat16* __restrict__ dY, int B,
                                   int H, int W, int Hp, int Wp, int Co) {
  pdl_prologue();
  // One thread = one 2x2 window x 8 channels: reads the pooled data once, writes all four
  // positions (windows outside the pooled grid write zeros so the whole H x W grid is defined).
  const int groups = Co >> 3;
  const int Wh = (W + 1) >> 1;
  const int rowp = blockIdx.y;                // b*Hh + hh, Hh = ceil(H/2)
  const int Hh = (H + 1) >> 1;
  const int b = rowp / Hh, hh = rowp - b * Hh;
  const int per_row = Wh * groups;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_row; t += gridDim.x * blockDim.x) {
    const int ww = t / groups;
    const int cg = t - ww * groups;
    uint32_t ow[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) ow[q][i] = 0u;
    if (hh < Hp && ww < Wp) {
      const int64_t o = (((int64_t)b * Hp + hh) * Wp + ww) * Co + cg * 8;
      const uint4 gv = *reinterpret_cast<const uint4*>(g + o);
      const uint4 yv = *reinterpret_cast<const uint4*>(ypool + o);
      const uint2 av = *reinterpret_cast<const uint2*>(amax + o);
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
      const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
      const uint32_t aw[2] = {av.x, av.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = i * 2 + j;
          const uint32_t a8 = (aw[c >> 2] >> ((c & 3) * 8)) & 3u;
          const uint32_t yb = (yw[i] >> (16 * j)) & 0xFFFFu;
          const bool on = yb != 0u && (yb & 0x8000u) == 0u;             // pooled > 0
          const uint32_t val = on ? (((gw[i] >> (16 * j)) & 0xFFFFu) << (16 * j)) : 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) ow[q][i] |= (a8 == (uint32_t)q) ? val : 0u;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int h = 2 * hh + (q >> 1), w = 2 * ww + (q & 1);
      if (h < H && w < W)
        *reinterpret_cast<uint4*>(dY + (((int64_t)b * H + h) * W + w) * Co + cg * 8) =
            make_uint4(ow[q][0], ow[q][1], ow[q][2], ow[q][3]);
    }
  }
}

void unpool_relu(const void* g, const uint8_t* amax, const void* ypool, void* dY, int B, int H, int W, int Hp,
                 int Wp, int Co, cudaStream_t st) {
  const int per_row = ((W + 1) / 2) * (Co / 8);
  dim3 grid((per_row + 127) / 128, B * ((H + 1) / 2));
  launch_pdl(unpool_relu_kernel, dim3(grid), dim3(128), 0, st, reinterpret_cast<const __nv_bfloat16*>(g), amax,
                                           reinterpret_cast<const __nv_bfloat16*>(ypool),
                                           reinterpret_cast<__nv_bfloat16*>(dY), B, H, W, Hp, Wp, Co);
  hefl::cuda::note_launch();
}

// bf16 shadow of the flat parameters ([Co][Ci][3][3] per conv) -> Wf [tap][Co][CK] (forward B
// operand, channel-padded) and Wd [tap][Ci][Co] (dgrad B operand).
__global__ void conv_weight_relayout_kernel(const __nv_bfloat16* __restrict__ shadow, const ConvLayerTable t,
                                            __nv_bfloat16* __restrict__ Wf, __nv_bfloat16* __restrict__ Wd, int l0,
                                            int spack0) {
  pdl_wait();   // parameter writer: no early trigger (see launch.cuh)
  const int l = blockIdx.y + l0;
  const int Ci = t.Ci[l], CK = t.CK[l], Co = t.Co[l];
  const __nv_bfloat16* w = shadow + t.w_off[l];
  if (l == 0 && spack0) {
    // s-packed layer 1: Wf [3 filter rows][Co][CK], k = s*Ci + ci (the input carries pixels w, w+1, w+2)
    const int nf = 3 * Co * CK;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
      const int k = i % CK, co = (i / CK) % Co, r = i / (CK * Co);
      const int sx = k / Ci, ci = k - sx * Ci;
      Wf[t.wf_off[l] + i] = k < 3 * Ci ? w[(co * Ci + ci) * 9 + r * 3 + sx] : __float2bfloat16(0.f);
    }
    return;
  }
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

void conv_weight_relayout(const void* shadow, const ConvLayerTable& t, void* Wf, void* Wd, int l0, int l1,
                          int spack0, cudaStream_t st) {
  if (l1 <= l0) return;
  dim3 grid(32, l1 - l0);
  launch_pdl(conv_weight_relayout_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(shadow), t,
                                                    reinterpret_cast<__nv_bfloat16*>(Wf),
                                                    reinterpret_cast<__nv_bfloat16*>(Wd), l0, spack0);
  hefl::cuda::note_launch();
}

// dW32 [9*CK+1][Co] (wgrad output, row 9*CK = bias gradient) -> flat fp32 gradient in the
// parameter layout ([Co][Ci][3][3], then bias); clears dW32 for the next step.
__global__ void conv_grad_finalize_kernel(float* __restrict__ dW32, const ConvLayerTable t,
                                          float* __restrict__ grad, int l0) {
  pdl_prologue();
  const int l = blockIdx.y + l0;
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
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

void conv_grad_finalize(float* dW32, const ConvLayerTable& t, float* grad, int l0, int l1, cudaStream_t st) {
  if (l1 <= l0) return;
  dim3 grid(32, l1 - l0);
  launch_pdl(conv_grad_finalize_kernel, dim3(grid), dim3(256), 0, st, dW32, t, grad, l0);
  const int l = l1 - 1;
  const int64_t begin = t.dw_off[l0];
  const int64_t total = t.dw_off[l] + (int64_t)(9 * t.CK[l] + 1) * t.Co[l] - begin;
  launch_pdl(zero_f32_kernel, dim3(64), dim3(256), 0, st, dW32 + begin, total);
  hefl::cuda::note_launch(2);
}

}  // namespace nn
}  // namespace hefl

namespace hefl {
namespace nn {

// One launch for the whole parameter update of the tcgen05 engine: consumes the wgrad buffers
// (dW32, rows = tap*CK + ci, last row = bias gradient), applies Adam to the matching entries of the
// flat fp32 parameters, writes the bf16 shadow and both tensor-core weight layouts (Wf [tap][Co][CK],
// Wd [tap][Ci][Co]) and clears dW32; the last grid row updates the dense-head parameters from the
// flat gradient. Replaces finalize + zero + step increment + Adam + relayout (5 launches).
struct AdamHyper {
  float lr, decay, beta1, beta2, eps;
};

__device__ __forceinline__ float adam_one(float p, float g, float& mi, float& vi, float alpha, float b1, float b2,
                                          float eps) {
  mi = b1 * mi + (1.f - b1) * g;
  vi = b2 * vi + (1.f - b2) * g * g;
  return p - alpha * mi / (sqrtf(vi) + eps);
}

__global__ void fused_update_kernel(float* __restrict__ dW32, const ConvLayerTable t, float* __restrict__ flat,
                                    float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                    __nv_bfloat16* __restrict__ shadow, __nv_bfloat16* __restrict__ Wf,
                                    __nv_bfloat16* __restrict__ Wd, const int64_t* __restrict__ step,
                                    const float* __restrict__ lr_scale, AdamHyper h, int64_t dense_off,
                                    int64_t n_trainable) {
  pdl_wait();   // parameter writer: no early trigger (see launch.cuh)
  const float tt = (float)(*step);
  const float lr_t = h.lr * (lr_scale ? *lr_scale : 1.f) / (1.f + h.decay * (tt - 1.f));
  const float alpha = lr_t * sqrtf(1.f - powf(h.beta2, tt)) / (1.f - powf(h.beta1, tt));
  const int l = blockIdx.y;
  if (l < t.n) {
    const int Ci = t.Ci[l], CK = t.CK[l], Co = t.Co[l];
    float* src = dW32 + t.dw_off[l];
    const int total = (9 * CK + 1) * Co;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int R = i / Co, co = i - R * Co;
      const float g = src[i];
      src[i] = 0.f;
      int64_t pi;
      int tap = 0, ci = 0;
      if (R == 9 * CK) {
        pi = t.b_off[l] + co;
      } else {
        tap = R / CK;
        ci = R - tap * CK;
        if (ci >= Ci) continue;                      // zero-padded input channels of layer 1
        pi = t.w_off[l] + ((int64_t)co * Ci + ci) * 9 + tap;
      }
      float mi = m[pi], vi = v[pi];
      const float p = adam_one(flat[pi], g, mi, vi, alpha, h.beta1, h.beta2, h.eps);
      flat[pi] = p; m[pi] = mi; v[pi] = vi;
      const __nv_bfloat16 pb = __float2bfloat16(p);
      shadow[pi] = pb;
      if (R != 9 * CK) {
        Wf[t.wf_off[l] + ((int64_t)tap * Co + co) * CK + ci] = pb;
        if (l > 0) Wd[t.wd_off[l] + ((int64_t)tap * Ci + ci) * Co + co] = pb;
      }
    }
  } else {
    for (int64_t i = dense_off + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_trainable;
         i += (int64_t)gridDim.x * blockDim.x) {
      float mi = m[i], vi = v[i];
      const float p = adam_one(flat[i], grad[i], mi, vi, alpha, h.beta1, h.beta2, h.eps);
      flat[i] = p; m[i] = mi; v[i] = vi;
      grad[i] = 0.f;
      shadow[i] = __float2bfloat16(p);
    }
  }
}

void fused_update(float* dW32, const ConvLayerTable& t, float* flat, float* grad, float* m, float* v, void* shadow,
                  void* Wf, void* Wd, const int64_t* step, const float* lr_scale, float lr, float decay, float beta1,
                  float beta2, float eps, int64_t dense_off, int64_t n_trainable, cudaStream_t st) {
  dim3 grid(48, t.n + 1);
  AdamHyper h{lr, decay, beta1, beta2, eps};
  launch_pdl(fused_update_kernel, dim3(grid), dim3(256), 0, st, dW32, t, flat, grad, m, v, reinterpret_cast<__nv_bfloat16*>(shadow),
                                            reinterpret_cast<__nv_bfloat16*>(Wf), reinterpret_cast<__nv_bfloat16*>(Wd),
                                            step, lr_scale, h, dense_off, n_trainable);
  hefl::cuda::note_launch();
}

}  // namespace nn
}  // namespace hefl

// ------------------------------------------------------------------------------------------
// Dense head of the sequential CNNs: Flatten -> Dense(H1, ReLU) -> Dense(H2, ReLU) -> Dense(C)
// -> softmax cross-entropy (FLPyfhelin.py:133-136, :141), forward AND backward in four launches
// (K16/K17). ~13 MFLOP at batch 32: latency- and shared-memory-bound, so the work is spread over
// many CTAs per phase, operands are staged with all loads in flight, and every thread reuses
// each shared-memory operand for several FMAs.
//   k1  h1 = relu(feat W1^T + b1)                                  grid H1/4
//   k2  h2, logits, loss, dlogits, dW3, db3, dh2                   1 CTA (0.3 MFLOP)
//   k3  dW2, db2, dh1 (column slices of 16)                        grid H1/16
//   k4  dW1, db1, dfeat (column slices of 16)                      grid F/16
// ------------------------------------------------------------------------------------------
namespace hefl {
namespace nn {

// k1: one CTA = 4 neurons; thread = (sample b, neuron jj); feat and the 4 weight rows in smem.
__global__ void __launch_bounds__(128)
head_fc1_fwd_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ W1,
                    const float* __restrict__ b1, float* __restrict__ h1, int B, int F, int H1) {
  pdl_trigger();
  extern __shared__ float sm1[];
  const int FP = F + 1;
  float* fs = sm1;                 // [B][F+1]
  float* ws = fs + B * FP;         // [4][F+1]
  const int j0 = blockIdx.x * 4;
  const int F8 = F >> 3, F4 = F >> 2;
  const uint4* src = reinterpret_cast<const uint4*>(feat);
  // weights first: they do not depend on the previous kernel, so this overlaps its tail (PDL)
#pragma unroll 4
  for (int i = threadIdx.x; i < 4 * F4; i += blockDim.x) {
    const int jj = i / F4, k4 = i - jj * F4;
    const float4 v = (j0 + jj) < H1 ? reinterpret_cast<const float4*>(W1 + (size_t)(j0 + jj) * F)[k4]
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    float* dst = ws + jj * FP + k4 * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  pdl_wait();
#pragma unroll 8
  for (int i = threadIdx.x; i < B * F8; i += blockDim.x) {   // B*F/8 16-byte loads, 8+ in flight per thread
    const uint4 v = src[i];
    const int b = i / F8, k8 = i - b * F8;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float* dst = fs + b * FP + k8 * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dst[2 * q] = __uint_as_float(w[q] << 16);
      dst[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
    }
  }
  __syncthreads();
  const int b = threadIdx.x >> 2, jj = threadIdx.x & 3;
  if (b < B && j0 + jj < H1) {
    const float* fr = fs + b * FP;
    const float* wr = ws + jj * FP;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
#pragma unroll 4
    for (; k + 4 <= F; k += 4) {
      a0 = fmaf(fr[k], wr[k], a0);
      a1 = fmaf(fr[k + 1], wr[k + 1], a1);
      a2 = fmaf(fr[k + 2], wr[k + 2], a2);
      a3 = fmaf(fr[k + 3], wr[k + 3], a3);
    }
    for (; k < F; ++k) a0 = fmaf(fr[k], wr[k], a0);
    const float acc = (a0 + a1) + (a2 + a3) + b1[j0 + jj];
    h1[b * H1 + j0 + jj] = acc > 0.f ? acc : 0.f;
  }
}

struct HeadMidArgs {
  const float* h1;      // [B][H1]
  const float *W2, *b2, *W3, *b3;
  const int64_t* y;     // [B]
  float *gW3, *gb3;
  float* dh2;           // [B][H2] out
  float* out;           // [2]: loss, ncorrect
  int64_t* step;        // optimiser step counter, incremented here once per training step (may be null)
  int B, H1, H2, C, train;
};

// k2: fc2 (each thread: 2 samples x 4 neurons, operands reused from registers), fc3, loss, dW3, dh2.
__global__ void __launch_bounds__(256)
head_mid_kernel(const HeadMidArgs a) {
  pdl_trigger();
  extern __shared__ float sm[];
  const int B = a.B, H1 = a.H1, H2 = a.H2, C = a.C;
  float* h1 = sm;                       // [B][H1+1]
  float* W2 = h1 + B * (H1 + 1);        // [H2][H1+1]
  float* h2 = W2 + H2 * (H1 + 1);       // [B][H2+1]
  float* lg = h2 + B * (H2 + 1);        // [B][C]  logits -> dlogits
  float* red = lg + B * C;              // [2*B]
  float* W3 = red + 2 * B;              // [C][H2]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int H14 = H1 >> 2;
#pragma unroll 8
  for (int i = tid; i < H2 * H14; i += nt) {
    const float4 v = reinterpret_cast<const float4*>(a.W2)[i];
    float* d = W2 + (i / H14) * (H1 + 1) + (i % H14) * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  for (int i = tid; i < C * H2; i += nt) W3[i] = a.W3[i];
  pdl_wait();          // weights staged above overlap the previous kernel; activations only from here
#pragma unroll 4
  for (int i = tid; i < B * H14; i += nt) {
    const float4 v = reinterpret_cast<const float4*>(a.h1)[i];
    float* d = h1 + (i / H14) * (H1 + 1) + (i % H14) * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  {  // fc2: tiles of 2 samples x 4 neurons
    const int tiles_j = H2 >> 2, tiles = (B >> 1) * tiles_j;
    for (int t = tid; t < tiles; t += nt) {
      const int b0 = (t / tiles_j) * 2, j0 = (t % tiles_j) * 4;
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int i = 0; i < H1; ++i) {
        const float x0 = h1[b0 * (H1 + 1) + i], x1 = h1[(b0 + 1) * (H1 + 1) + i];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float w = W2[(j0 + q) * (H1 + 1) + i];
          acc[0][q] = fmaf(x0, w, acc[0][q]);
          acc[1][q] = fmaf(x1, w, acc[1][q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float bb = a.b2[j0 + q];
        const float y0 = acc[0][q] + bb, y1 = acc[1][q] + bb;
        h2[b0 * (H2 + 1) + j0 + q] = y0 > 0.f ? y0 : 0.f;
        h2[(b0 + 1) * (H2 + 1) + j0 + q] = y1 > 0.f ? y1 : 0.f;
      }
    }
  }
  __syncthreads();
  for (int o = tid; o < B * C; o += nt) {                      // fc3
    const int b = o / C, c = o % C;
    float a0 = a.b3[c], a1 = 0.f;
    int j = 0;
    for (; j + 2 <= H2; j += 2) {
      a0 = fmaf(h2[b * (H2 + 1) + j], W3[c * H2 + j], a0);
      a1 = fmaf(h2[b * (H2 + 1) + j + 1], W3[c * H2 + j + 1], a1);
    }
    for (; j < H2; ++j) a0 = fmaf(h2[b * (H2 + 1) + j], W3[c * H2 + j], a0);
    lg[o] = a0 + a1;
  }
  __syncthreads();
  if (tid < 32) {                                              // warp 0: softmax cross-entropy, one lane per sample
    float l = 0.f, nc = 0.f;
    if (tid < B) {
      float mx = -1e30f;
      int am = 0;
      for (int c = 0; c < C; ++c) if (lg[tid * C + c] > mx) { mx = lg[tid * C + c]; am = c; }
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += __expf(lg[tid * C + c] - mx);
      const int yy = (int)a.y[tid];
      l = -(lg[tid * C + yy] - mx - __logf(s));
      nc = am == yy ? 1.f : 0.f;
      const float invB = 1.f / B;
      for (int c = 0; c < C; ++c) {
        const float p = __expf(lg[tid * C + c] - mx) / s;
        lg[tid * C + c] = (p - (c == yy ? 1.f : 0.f)) * invB;   // dlogits
      }
    }
    for (int off = 16; off; off >>= 1) {                       // all 32 lanes take part (B <= 32)
      l += __shfl_xor_sync(0xffffffffu, l, off);
      nc += __shfl_xor_sync(0xffffffffu, nc, off);
    }
    if (tid == 0) {
      a.out[0] = l / B;
      a.out[1] = nc;
      if (a.train && a.step) *a.step += 1;     // read by the update kernel after this one; no other reader now
    }
  }
  __syncthreads();
  if (!a.train) return;
  for (int o = tid; o < C * H2 + C; o += nt) {                 // dW3, db3
    if (o < C * H2) {
      const int c = o / H2, j = o % H2;
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc = fmaf(lg[b * C + c], h2[b * (H2 + 1) + j], acc);
      a.gW3[o] = acc;
    } else {
      const int c = o - C * H2;
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc += lg[b * C + c];
      a.gb3[c] = acc;
    }
  }
  for (int o = tid; o < B * H2; o += nt) {                     // dh2 -> global for k3
    const int b = o / H2, j = o % H2;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(lg[b * C + c], W3[c * H2 + j], acc);
    a.dh2[o] = h2[b * (H2 + 1) + j] > 0.f ? acc : 0.f;
  }
}

// k3: one CTA owns 16 columns i of layer 2's input: dW2[:, i], dh1[:, i]; CTA 0 also db2.
__global__ void __launch_bounds__(256)
head_fc2_bwd_kernel(const float* __restrict__ dh2g, const float* __restrict__ h1g, const float* __restrict__ W2,
                    float* __restrict__ gW2, float* __restrict__ gb2, float* __restrict__ dh1, int B, int H1, int H2) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* d = sm;                       // [B][H2+1]
  float* hs = d + B * (H2 + 1);        // [B][16]
  float* ws = hs + B * 16;             // [H2][17]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int i0 = blockIdx.x * 16;
#pragma unroll 4
  for (int i = tid; i < H2 * 16; i += nt) ws[(i >> 4) * 17 + (i & 15)] = (i0 + (i & 15)) < H1 ? W2[(size_t)(i >> 4) * H1 + i0 + (i & 15)] : 0.f;
  pdl_wait();
#pragma unroll 4
  for (int i = tid; i < B * H2; i += nt) d[(i / H2) * (H2 + 1) + i % H2] = dh2g[i];
#pragma unroll 2
  for (int i = tid; i < B * 16; i += nt) hs[i] = (i0 + (i & 15)) < H1 ? h1g[(i >> 4) * H1 + i0 + (i & 15)] : 0.f;
  __syncthreads();
  for (int o = tid; o < H2 * 16; o += nt) {                    // dW2[j][i0+k]
    const int j = o >> 4, k = o & 15;
    float a0 = 0.f, a1 = 0.f;
    int b = 0;
    for (; b + 2 <= B; b += 2) {
      a0 = fmaf(d[b * (H2 + 1) + j], hs[b * 16 + k], a0);
      a1 = fmaf(d[(b + 1) * (H2 + 1) + j], hs[(b + 1) * 16 + k], a1);
    }
    for (; b < B; ++b) a0 = fmaf(d[b * (H2 + 1) + j], hs[b * 16 + k], a0);
    if (i0 + k < H1) gW2[(size_t)j * H1 + i0 + k] = a0 + a1;
  }
  for (int o = tid; o < B * 16; o += nt) {                     // dh1[b][i0+k]
    const int b = o >> 4, k = o & 15;
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 2 <= H2; j += 2) {
      a0 = fmaf(d[b * (H2 + 1) + j], ws[j * 17 + k], a0);
      a1 = fmaf(d[b * (H2 + 1) + j + 1], ws[(j + 1) * 17 + k], a1);
    }
    for (; j < H2; ++j) a0 = fmaf(d[b * (H2 + 1) + j], ws[j * 17 + k], a0);
    if (i0 + k < H1) dh1[b * H1 + i0 + k] = hs[b * 16 + k] > 0.f ? (a0 + a1) : 0.f;
  }
  if (blockIdx.x == 0) {
    for (int j = tid; j < H2; j += nt) {
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc += d[b * (H2 + 1) + j];
      gb2[j] = acc;
    }
  }
}

// k4: dW1[j][k] = sum_b dh1[b][j] feat[b][k]; db1[j] = sum_b dh1[b][j]; dfeat[b][k] = sum_j dh1[b][j] W1[j][k].
// One CTA owns 16 input columns k.
__global__ void __launch_bounds__(256)
head_fc1_bwd_kernel(const float* __restrict__ dh1, const __nv_bfloat16* __restrict__ feat,
                    const float* __restrict__ W1, float* __restrict__ gW1, float* __restrict__ gb1,
                    __nv_bfloat16* __restrict__ dfeat, int B, int F, int H1) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* d = sm;                       // [B][H1+1]
  float* fs = d + B * (H1 + 1);        // [B][16]
  float* ws = fs + B * 16;             // [H1][17]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int k0 = blockIdx.x * 16;
#pragma unroll 8
  for (int i = tid; i < H1 * 16; i += nt) {
    const int j = i >> 4, k = i & 15;
    ws[j * 17 + k] = k0 + k < F ? W1[(size_t)j * F + k0 + k] : 0.f;
  }
  pdl_wait();
#pragma unroll 8
  for (int i = tid; i < B * H1; i += nt) d[(i / H1) * (H1 + 1) + i % H1] = dh1[i];
#pragma unroll 2
  for (int i = tid; i < B * 16; i += nt) {
    const int b = i >> 4, k = i & 15;
    fs[i] = k0 + k < F ? __bfloat162float(feat[b * F + k0 + k]) : 0.f;
  }
  __syncthreads();
  for (int o = tid; o < H1 * 16; o += nt) {
    const int j = o >> 4, k = o & 15;
    float a0 = 0.f, a1 = 0.f;
    int b = 0;
    for (; b + 2 <= B; b += 2) {
      a0 = fmaf(d[b * (H1 + 1) + j], fs[b * 16 + k], a0);
      a1 = fmaf(d[(b + 1) * (H1 + 1) + j], fs[(b + 1) * 16 + k], a1);
    }
    for (; b < B; ++b) a0 = fmaf(d[b * (H1 + 1) + j], fs[b * 16 + k], a0);
    if (k0 + k < F) gW1[(size_t)j * F + k0 + k] = a0 + a1;
  }
  for (int o = tid; o < B * 16; o += nt) {
    const int b = o >> 4, k = o & 15;
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 2 <= H1; j += 2) {
      a0 = fmaf(d[b * (H1 + 1) + j], ws[j * 17 + k], a0);
      a1 = fmaf(d[b * (H1 + 1) + j + 1], ws[(j + 1) * 17 + k], a1);
    }
    for (; j < H1; ++j) a0 = fmaf(d[b * (H1 + 1) + j], ws[j * 17 + k], a0);
    if (k0 + k < F) dfeat[b * F + k0 + k] = __float2bfloat16(a0 + a1);
  }
  if (blockIdx.x == 0) {
    for (int j = tid; j < H1; j += nt) {
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc += d[b * (H1 + 1) + j];
      gb1[j] = acc;
    }
  }
}

void head_forward_backward(const void* feat, const float* W1, const float* b1, const float* W2, const float* b2,
                           const float* W3, const float* b3, const int64_t* y, float* gW1, float* gb1, float* gW2,
                           float* gb2, float* gW3, float* gb3, void* dfeat, float* h1_buf, float* dh1_buf,
                           float* out, int64_t* step, int B, int F, int H1, int H2, int C, int train, cudaStream_t st) {
  const auto* f = reinterpret_cast<const __nv_bfloat16*>(feat);
  float* dh2_buf = dh1_buf + (size_t)B * H1;                  // scratch tail: [B][H2]
  const int smem_a = (B + 4) * (F + 1) * 4;
  cudaFuncSetAttribute(head_fc1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
  launch_pdl(head_fc1_fwd_kernel, dim3((H1 + 3) / 4), dim3(128), smem_a, st, f, W1, b1, h1_buf, B, F, H1);
  HeadMidArgs a{h1_buf, W2, b2, W3, b3, y, gW3, gb3, dh2_buf, out, step, B, H1, H2, C, train};
  const int smem_b = (B * (H1 + 1) + H2 * (H1 + 1) + B * (H2 + 1) + B * C + 2 * B + C * H2) * 4;
  cudaFuncSetAttribute(head_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
  launch_pdl(head_mid_kernel, dim3(1), dim3(256), smem_b, st, a);
  hefl::cuda::note_launch(2);
  if (train) {
    const int smem_c = (B * (H2 + 1) + B * 16 + H2 * 17) * 4;
    launch_pdl(head_fc2_bwd_kernel, dim3((H1 + 15) / 16), dim3(256), smem_c, st, dh2_buf, h1_buf, W2, gW2, gb2, dh1_buf, B, H1, H2);
    const int smem_d = (B * (H1 + 1) + B * 16 + H1 * 17) * 4;
    cudaFuncSetAttribute(head_fc1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
    launch_pdl(head_fc1_bwd_kernel, dim3((F + 15) / 16), dim3(256), smem_d, st, dh1_buf, f, W1, gW1, gb1,
                                                            reinterpret_cast<__nv_bfloat16*>(dfeat), B, F, H1);
    hefl::cuda::note_launch(2);
  }
}

}  // namespace nn
}  // namespace hefl
