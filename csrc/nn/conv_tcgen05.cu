// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Activations are NHWC bf16 viewed as a 2-D matrix [P = B*H*W pixels, C channels]. A 3x3 valid
// convolution is then 9 "tap" GEMMs whose A operand is the SAME matrix shifted by
// r*W + s rows:   Y[m, :] = sum_{r,s} X[m + r*W + s, :] * W[r,s]^T   (m on the input grid; rows
// with h >= H-2 or w >= W-2 are garbage and masked). Each tap's A tile is therefore one plain
// 2-D TMA box — no im2col buffer, no gather.
//
//   conv_fwd_pool : per CTA tile = 2 image rows x 64 columns (128 GEMM rows) so that the
//                   epilogue can do bias + ReLU + 2x2 max-pool + argmax straight out of TMEM and
//                   write only the pooled tensor (4x less HBM traffic than the conv output).
//   conv_dgrad    : dX[m,:] = sum_taps dY[m - off, :] * W[r,s]   (same kernel, negative offsets,
//                   plain bf16 store).
//   conv_wgrad    : dW[r,s] = sum_m X[m+off,:]^T dY[m,:]. Both operands are consumed in their
//                   native [P, C] layout as MN-major UMMA operands (pixels = K), so no transposed
//                   copies exist; taps are stacked along M (128/C tap atoms per MMA group), split-K
//                   over pixels, fp32 vector-RED into a [9*C+1, Co] buffer whose extra row is the
//                   bias gradient (an all-ones tap atom).
//
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> global). smem ring of NSTAGE tap tiles with
// full/empty mbarriers; two TMEM accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1. Weights for all 9 taps stay resident in shared memory for the CTA's lifetime.
#include <cuda_bf16.h>

#include <cstdio>
#include <stdexcept>

#include "../he/kernels.h"
#include "nn.h"
#include "tc_common.cuh"

namespace hefl {
namespace nn {

using namespace hefl::tc;

// ------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor [outer rows][inner cols], row pitch in bytes; box = box_inner x box_outer.
static CUtensorMap make_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                            uint32_t box_inner, uint32_t box_outer) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const uint32_t inner_bytes = box_inner * 2;
  CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u",
             (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes,
             box_inner, box_outer);
    throw std::runtime_error(buf);
  }
  return m;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// G1: tap-GEMM kernel (forward + pool epilogue, or dgrad with plain store)
// ------------------------------------------------------------------------------------------
struct TapGemmArgs {
  int B, H, W;          // input grid of the A matrix
  int Hp, Wp;           // pooled output grid (POOL)
  int tiles_w;          // column tiles per row pair (POOL)
  int num_tiles;
  int P;                // B*H*W
  int sign;             // +1 forward (m + off), -1 dgrad (m - off)
  const float* bias;    // [CO] (POOL)
  __nv_bfloat16* out;   // POOL: [B,Hp,Wp,CO]; else [P,CO]
  uint8_t* argmax;      // POOL, may be null
};

template <int CK, int CO, bool POOL>
struct TapGemmCfg {
  static constexpr int KB = CK < 64 ? CK : 64;            // channels per k-block (one swizzle atom)
  static constexpr int NKB = CK / KB;
  static constexpr int ROW_BYTES = KB * 2;
  static constexpr int A_SUB = 128 * ROW_BYTES;           // one k-block of one tap tile
  static constexpr int A_STAGE = A_SUB * NKB;
  static constexpr int W_SUB = CO * ROW_BYTES;
  static constexpr int W_TAP = W_SUB * NKB;
  static constexpr int W_BYTES = 9 * W_TAP;
  static constexpr int EXCH = POOL ? (2 * 32 * 16 * 4 + 2 * 32 * 16) : 0;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 224 * 1024;
  static constexpr int NSTAGE_RAW = (BUDGET - W_BYTES - EXCH - BAR_BYTES - 1024) / A_STAGE;
  static constexpr int NSTAGE = NSTAGE_RAW > 6 ? 6 : NSTAGE_RAW;
  static constexpr int SMEM = W_BYTES + NSTAGE * A_STAGE + EXCH + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * CO <= 32 ? 32 : (2 * CO <= 64 ? 64 : (2 * CO <= 128 ? 128 : 256));
  static_assert(NSTAGE >= 2, "not enough shared memory for a 2-stage pipeline");
  static_assert(CO % 32 == 0 && CO <= 128, "CO must be 32, 64, 96 or 128");
};

template <int CK, int CO, bool POOL>
__global__ void __launch_bounds__(192, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const TapGemmArgs a) {
  using Cfg = TapGemmCfg<CK, CO, POOL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::W_BYTES;
  float* exch_v = reinterpret_cast<float*>(sA + Cfg::NSTAGE * Cfg::A_STAGE);
  uint8_t* exch_i = reinterpret_cast<uint8_t*>(exch_v) + 2 * 32 * 16 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + Cfg::NSTAGE * Cfg::A_STAGE + Cfg::EXCH);
  uint64_t* full = bars;                       // [NSTAGE]
  uint64_t* empty = bars + Cfg::NSTAGE;        // [NSTAGE]
  uint64_t* wfull = bars + 2 * Cfg::NSTAGE;    // [1]
  uint64_t* tfull = wfull + 1;                 // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(wfull, Cfg::W_BYTES);
      for (int tap = 0; tap < 9; ++tap)
        for (int kb = 0; kb < Cfg::NKB; ++kb)
          tma_load_2d(sW + tap * Cfg::W_TAP + kb * Cfg::W_SUB, &tmW, kb * Cfg::KB, tap * CO, wfull);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        int base0, base1 = 0;
        if (POOL) {
          const int per_img = a.Hp * a.tiles_w;
          const int b = t / per_img;
          const int rem = t - b * per_img;
          const int hp = rem / a.tiles_w;
          const int tw = rem - hp * a.tiles_w;
          base0 = (b * a.H + 2 * hp) * a.W + tw * 64;
          base1 = base0 + a.W;
        } else {
          base0 = t * 128;
        }
        for (int tap = 0; tap < 9; ++tap) {
          const int off = a.sign * ((tap / 3) * a.W + (tap % 3));
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::A_STAGE);
          uint8_t* dst = sA + stage * Cfg::A_STAGE;
          for (int kb = 0; kb < Cfg::NKB; ++kb) {
            if (POOL) {
              tma_load_2d(dst + kb * Cfg::A_SUB, &tmA, kb * Cfg::KB, base0 + off, &full[stage]);
              tma_load_2d(dst + kb * Cfg::A_SUB + 64 * Cfg::ROW_BYTES, &tmA, kb * Cfg::KB, base1 + off, &full[stage]);
            } else {
              tma_load_2d(dst + kb * Cfg::A_SUB, &tmA, kb * Cfg::KB, base0 + off, &full[stage]);
            }
          }
          if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(128, CO);
    mbar_wait(wfull, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_STAGE);
          const uint32_t w_addr = smem_u32(sW + tap * Cfg::W_TAP);
#pragma unroll
          for (int kb = 0; kb < Cfg::NKB; ++kb) {
#pragma unroll
            for (int k = 0; k < Cfg::KB / 16; ++k) {
              const uint64_t ad = make_kmajor_desc(a_addr + kb * Cfg::A_SUB + k * 32, Cfg::ROW_BYTES);
              const uint64_t bd = make_kmajor_desc(w_addr + kb * Cfg::W_SUB + k * 32, Cfg::ROW_BYTES);
              umma_bf16(tmem_base + acc * CO, ad, bd, idesc, (tap | kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (tap == 8) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===== epilogue warps (2..5) =====
    const int qd = warp & 3;                       // TMEM lane quadrant this warp may read
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (POOL) {
        const int per_img = a.Hp * a.tiles_w;
        const int b = t / per_img;
        const int rem = t - b * per_img;
        const int hp = rem / a.tiles_w;
        const int tw = rem - hp * a.tiles_w;
        const int col = tw * 64 + (qd & 1) * 32 + lane;      // conv-output column of this thread
        const int wp = col >> 1;
        const bool lower = qd >= 2;                          // image row h+1
        const bool even = (lane & 1) == 0;
        const int slot = lane >> 1;
        const int half = qd & 1;
#pragma unroll 1
        for (int ch = 0; ch < CO / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + acc * CO + ch * 32, v);
          uint32_t hbits = 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float x = v[c] + __ldg(a.bias + ch * 32 + c);
            x = x > 0.f ? x : 0.f;
            const float o = __shfl_xor_sync(0xffffffffu, x, 1);
            if (o > x) { x = o; hbits |= 1u << c; }        // meaningful on even lanes only
            v[c] = x;
          }
          if (lower && even) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              exch_v[(half * 32 + c) * 16 + slot] = v[c];
              exch_i[(half * 32 + c) * 16 + slot] = (uint8_t)((hbits >> c) & 1u);
            }
          }
          named_barrier_sync(1, 128);
          if (!lower && even && wp < a.Wp) {
            const size_t o = ((size_t)(b * a.Hp + hp) * a.Wp + wp) * CO + ch * 32;
            uint32_t packed[16];
            uint32_t idx4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) idx4[i] = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              float x = v[c];
              uint32_t id = (hbits >> c) & 1u;
              const float pv = exch_v[(half * 32 + c) * 16 + slot];
              if (pv > x) { x = pv; id = 2u + exch_i[(half * 32 + c) * 16 + slot]; }
              const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(x));
              if (c & 1) packed[c >> 1] |= hb << 16; else packed[c >> 1] = hb;
              idx4[c >> 2] |= id << ((c & 3) * 8);
            }
            uint4* dst = reinterpret_cast<uint4*>(a.out + o);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            if (a.argmax) {
              uint4* di = reinterpret_cast<uint4*>(a.argmax + o);
              di[0] = make_uint4(idx4[0], idx4[1], idx4[2], idx4[3]);
              di[1] = make_uint4(idx4[4], idx4[5], idx4[6], idx4[7]);
            }
          }
          named_barrier_sync(1, 128);
        }
      } else {
        const int m = t * 128 + qd * 32 + lane;
#pragma unroll 1
        for (int ch = 0; ch < CO / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + acc * CO + ch * 32, v);
          if (m < a.P) {
            uint32_t packed[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const uint32_t lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(v[c]));
              const uint32_t hi = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(v[c + 1]));
              packed[c >> 1] = lo | (hi << 16);
            }
            uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)m * CO + ch * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int CK, int CO, bool POOL>
static void launch_tap_gemm(const __nv_bfloat16* A, const __nv_bfloat16* Wt, TapGemmArgs a, cudaStream_t st) {
  using Cfg = TapGemmCfg<CK, CO, POOL>;
  const CUtensorMap tmA = make_map(A, CK, (uint64_t)a.P, (uint64_t)CK * 2, Cfg::KB, POOL ? 64 : 128);
  const CUtensorMap tmW = make_map(Wt, CK, (uint64_t)9 * CO, (uint64_t)CK * 2, Cfg::KB, CO);
  auto kern = tap_gemm_kernel<CK, CO, POOL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  if (grid < 1) grid = 1;
  kern<<<grid, 192, Cfg::SMEM, st>>>(tmA, tmW, a);
  hefl::cuda::note_launch();
}

void conv_fwd_pool(const void* X, const void* Wf, const float* bias, void* out, uint8_t* argmax, int B, int H,
                   int W, int CK, int CO, cudaStream_t st) {
  TapGemmArgs a{};
  a.B = B; a.H = H; a.W = W;
  a.Hp = (H - 2) / 2; a.Wp = (W - 2) / 2;
  a.tiles_w = (2 * a.Wp + 63) / 64;
  a.num_tiles = B * a.Hp * a.tiles_w;
  a.P = B * H * W;
  a.sign = 1;
  a.bias = bias;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.argmax = argmax;
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(X);
  const auto* w = reinterpret_cast<const __nv_bfloat16*>(Wf);
  if (CK == 16 && CO == 32) launch_tap_gemm<16, 32, true>(x, w, a, st);
  else if (CK == 32 && CO == 32) launch_tap_gemm<32, 32, true>(x, w, a, st);
  else if (CK == 32 && CO == 64) launch_tap_gemm<32, 64, true>(x, w, a, st);
  else if (CK == 64 && CO == 64) launch_tap_gemm<64, 64, true>(x, w, a, st);
  else if (CK == 64 && CO == 128) launch_tap_gemm<64, 128, true>(x, w, a, st);
  else throw std::runtime_error("conv_fwd_pool: unsupported (CK, CO)");
}

void conv_dgrad(const void* dY, const void* Wd, void* dX, int B, int H, int W, int CK, int CO, cudaStream_t st) {
  TapGemmArgs a{};
  a.B = B; a.H = H; a.W = W;
  a.P = B * H * W;
  a.num_tiles = (a.P + 127) / 128;
  a.sign = -1;
  a.out = reinterpret_cast<__nv_bfloat16*>(dX);
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(dY);
  const auto* w = reinterpret_cast<const __nv_bfloat16*>(Wd);
  if (CK == 32 && CO == 32) launch_tap_gemm<32, 32, false>(x, w, a, st);
  else if (CK == 64 && CO == 32) launch_tap_gemm<64, 32, false>(x, w, a, st);
  else if (CK == 64 && CO == 64) launch_tap_gemm<64, 64, false>(x, w, a, st);
  else if (CK == 128 && CO == 64) launch_tap_gemm<128, 64, false>(x, w, a, st);
  else throw std::runtime_error("conv_dgrad: unsupported (CK, CO)");
}

// ------------------------------------------------------------------------------------------
// G2: weight-gradient kernel (MN-major operands: consumes X [P,CK] and dY [P,Co] as they are)
// ------------------------------------------------------------------------------------------
struct WgradArgs {
  int W;            // image width of the layer input grid (tap offset = r*W + s)
  int P;            // pixels
  int nchunks;      // ceil(P / 64)
  int Co;           // total output channels (row pitch of dW32)
  float* dW32;      // [9*CK + 1][Co] fp32, accumulated with vector RED
};

template <int CK, int COT>
struct WgradCfg {
  static constexpr int ATOM_A = CK * 2;                        // bytes of one k-row of one tap atom
  static constexpr int ATOM_B = COT * 2;
  static constexpr int TPG = 128 / CK;                         // tap atoms per MMA group (M = 128)
  static constexpr int NG = (10 + TPG - 1) / TPG;              // 9 taps + the all-ones atom (bias grad)
  static constexpr int A_TAP = 64 * ATOM_A;                    // [64 px][CK] bf16
  static constexpr int A_STAGE = NG * 128 * 128;               // NG groups x 16 KB
  static constexpr int B_STAGE = 64 * ATOM_B;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int TX = 9 * A_TAP + B_STAGE;
  static constexpr int NSTAGE_RAW = (220 * 1024) / STAGE;
  static constexpr int NSTAGE = NSTAGE_RAW > 4 ? 4 : NSTAGE_RAW;
  static constexpr int SMEM = NSTAGE * STAGE + 256 + 1024;
  static constexpr int COLS = NG * COT;
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : (COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512)));
  static_assert(NSTAGE >= 2, "wgrad pipeline needs two stages");
  static_assert(COLS <= 512, "accumulators exceed TMEM");
  static_assert(CK == 16 || CK == 32 || CK == 64, "CK must be one swizzle atom");
  static_assert(COT == 32 || COT == 64, "COT must be one swizzle atom");
};

template <int CK, int COT>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
             const WgradArgs a) {
  using Cfg = WgradCfg<CK, COT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::NSTAGE;
  uint64_t* done = bars + 2 * Cfg::NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * COT;

  // tap atom #9 is all ones: its GEMM rows are the column sums of dY = the bias gradient. It is
  // never touched by TMA, so it is written once per stage buffer (uniform => swizzle-agnostic).
  for (int s = 0; s < Cfg::NSTAGE; ++s) {
    uint32_t* atom = reinterpret_cast<uint32_t*>(smem + s * Cfg::STAGE + 9 * Cfg::A_TAP);
    for (int i = threadIdx.x; i < Cfg::A_TAP / 4; i += blockDim.x) atom[i] = 0x3F803F80u;  // bf16 1.0 x2
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();   // generic-proxy writes (ones atom) visible to the tensor-core async proxy
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x;
  const int step = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = first; c < a.nchunks; c += step) {
        const int k0 = c * 64;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::TX);
        uint8_t* sa = smem + stage * Cfg::STAGE;
        for (int tap = 0; tap < 9; ++tap)
          tma_load_2d(sa + tap * Cfg::A_TAP, &tmX, 0, k0 + (tap / 3) * a.W + (tap % 3), &full[stage]);
        tma_load_2d(sa + Cfg::A_STAGE, &tmDY, co0, k0, &full[stage]);
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, COT, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    bool firstc = true;
    for (int c = first; c < a.nchunks; c += step) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t b_addr = a_addr + Cfg::A_STAGE;
#pragma unroll
        for (int g = 0; g < Cfg::NG; ++g) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x UMMA_K(16) pixels per 64-pixel chunk
            const uint64_t ad = make_mnmajor_desc(a_addr + g * 128 * 128 + k * 16 * Cfg::ATOM_A, Cfg::ATOM_A, Cfg::A_TAP);
            const uint64_t bd = make_mnmajor_desc(b_addr + k * 16 * Cfg::ATOM_B, Cfg::ATOM_B, Cfg::B_STAGE);
            umma_bf16(tmem_base + g * COT, ad, bd, idesc, (firstc && k == 0) ? 0u : 1u);
          }
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      firstc = false;
      if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) umma_commit(done);
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    if (first < a.nchunks) {
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < Cfg::NG; ++g) {
        const int R = g * 128 + qd * 32 + lane;      // global row: tap*CK + ci, or 9*CK = bias row
#pragma unroll 1
        for (int ch = 0; ch < COT / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + g * COT + ch * 32, v);
          if (R <= 9 * CK) {
            float* dst = a.dW32 + (size_t)R * a.Co + co0 + ch * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[c]), "f"(v[c + 1]),
                           "f"(v[c + 2]), "f"(v[c + 3])
                           : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int CK, int COT>
static void launch_wgrad(const __nv_bfloat16* X, const __nv_bfloat16* DY, WgradArgs a, int Co, cudaStream_t st) {
  using Cfg = WgradCfg<CK, COT>;
  const CUtensorMap tmX = make_map(X, CK, (uint64_t)a.P, (uint64_t)CK * 2, CK, 64);
  const CUtensorMap tmD = make_map(DY, Co, (uint64_t)a.P, (uint64_t)Co * 2, COT, 64);
  auto kern = wgrad_kernel<CK, COT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  int split = a.nchunks / 4;
  if (split < 1) split = 1;
  const int cot = Co / COT;
  int cap = num_sms() / cot;
  if (split > cap) split = cap;
  dim3 grid(split, cot);
  kern<<<grid, 192, Cfg::SMEM, st>>>(tmX, tmD, a);
  hefl::cuda::note_launch();
}

void conv_wgrad(const void* X, const void* DY, float* dW32, int P, int W, int CK, int Co, cudaStream_t st) {
  WgradArgs a{};
  a.W = W;
  a.P = P;
  a.nchunks = (P + 63) / 64;
  a.Co = Co;
  a.dW32 = dW32;
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(X);
  const auto* d = reinterpret_cast<const __nv_bfloat16*>(DY);
  if (CK == 16 && Co == 32) launch_wgrad<16, 32>(x, d, a, Co, st);
  else if (CK == 32 && Co == 32) launch_wgrad<32, 32>(x, d, a, Co, st);
  else if (CK == 32 && Co == 64) launch_wgrad<32, 64>(x, d, a, Co, st);
  else if (CK == 64 && Co == 64) launch_wgrad<64, 64>(x, d, a, Co, st);
  else if (CK == 64 && Co == 128) launch_wgrad<64, 64>(x, d, a, Co, st);
  else throw std::runtime_error("conv_wgrad: unsupported (CK, Co)");
}

}  // namespace nn
}  // namespace hefl
