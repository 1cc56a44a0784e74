// General tcgen05 / TMEM / TMA GEMM kernels for the ResNet convolutions (BASELINE.json
// configs[2], [4]). On NHWC activations
//
//   * a 1x1 convolution is the GEMM  Y[pixels, Co] = X[pixels, Ci] . W[Co, Ci]^T,
//   * a 3x3 / pad-1 / stride-1 convolution is NINE such GEMMs accumulated in TMEM: tap (r, s) multiplies the same
//     activation matrix shifted by (r-1)*Wp + (s-1) rows, when the activations live on a zero-padded
//     (H+2) x (W+2) grid -- the shift is just the row coordinate of the TMA copy (no im2col),
//   * dgrad is the same kernel with the transposed / flipped filter,
//   * wgrad is  dW[tap][Co, Ci] = sum_pixels dY[p, Co] . X[p + shift(tap), Ci]: both operands are consumed as
//     stored (MN-major UMMA descriptors, the pixel axis is K), split over pixel ranges, fp32 vector RED.
//
// kmajor_gemm_kernel : C[M, N] = sum_taps A[M + shift_t, K] . B_t[N, K]^T, bf16 (kind::f16) or e4m3 (kind::f8f6f4,
//                      per-tensor scales folded into the epilogue), fp32 accumulation in TMEM, bf16 out.
//                      Warp-specialised and persistent: 1 TMA producer warp, 1 MMA-issuing warp, 4 epilogue warps;
//                      5-stage smem ring (128 B swizzle), two TMEM accumulator buffers so the epilogue of tile i
//                      overlaps the main loop of tile i+1.
// wgrad_mn_kernel    : up to three taps (one filter row) per CTA share the dY stage.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <stdexcept>

#include "../he/kernels.h"
#include "nn.h"
#include "tc_common.cuh"

namespace hefl {
namespace nn {

using namespace hefl::tc;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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

// 2-D tensor [rows][cols] of `esize`-byte elements, row pitch in bytes; box = box_cols x box_rows, 128 B swizzle.
CUtensorMap map2d(const void* ptr, int esize, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_cols,
                  uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                 const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(gemm) failed (%d): cols=%llu rows=%llu pitch=%llu box=%ux%u", (int)r,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_bytes, box_cols, box_rows);
    throw std::runtime_error(buf);
  }
  return m;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// e4m3 x e4m3 -> fp32 (kind::f8f6f4): same descriptor layout as kind::f16 with a/b format 0 (E4M3)
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f8_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// K-major GEMM with row-shifted taps
// ------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  int m_tiles, n_tiles;   // 128-row tiles of the A grid, BN-column tiles
  int kblocks;            // K / (128 bytes of K)
  int taps;
  int shift[9];           // row shift of A for tap t
  int n_total;            // N (rows of B per tap, row pitch of C in elements)
  int64_t m_rows;         // rows of the A grid (dense: valid output rows)
  // padded-grid mode (3x3 convolutions): A rows index a zero-padded [B][Hp][Wp] grid, only interior pixels are
  // written, to row (b*H + hp-1)*W + wp-1 of C
  int padded, Bn, H, W, Hp, Wp;
  __nv_bfloat16* C;
  const float* scale_a;   // fp8: de-quantisation factors (device scalars) or null
  const float* scale_b;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = 128 * 128;          // 128 rows x 128 bytes of K
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = BN == 128 ? 5 : 6;
  static constexpr int SMEM = NSTAGE * STAGE + 256 + 1024;
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 256;
};

template <int BN, bool FP8>
__global__ void __launch_bounds__(192, 1)
kmajor_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE);
  uint64_t* full = bars;                          // [NSTAGE]
  uint64_t* empty = bars + Cfg::NSTAGE;           // [NSTAGE]
  uint64_t* tfull = empty + Cfg::NSTAGE;          // [2]
  uint64_t* tempty = tfull + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int KE = FP8 ? 128 : 64;              // K elements per 128-byte k-block

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = a.m_tiles * a.n_tiles;
  const int ksteps = a.taps * a.kblocks;

  if (warp == 4) {
    // ===== TMA producer =====
    if (elect_one()) {
      int st = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mb = t / a.n_tiles, nb = t - mb * a.n_tiles;
        for (int tap = 0; tap < a.taps; ++tap) {
          const int arow = mb * 128 + a.shift[tap];
          const int brow = tap * a.n_total + nb * BN;
          for (int kb = 0; kb < a.kblocks; ++kb) {
            mbar_wait(&empty[st], phase ^ 1u);
            uint8_t* sA = smem + st * Cfg::STAGE;
            mbar_expect_tx(&full[st], Cfg::STAGE);
            tma_load_2d(sA, &tmA, kb * KE, arow, &full[st]);
            tma_load_2d(sA + Cfg::A_BYTES, &tmB, kb * KE, brow, &full[st]);
            if (++st == Cfg::NSTAGE) { st = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(128, BN) : make_idesc_bf16(128, BN);
    constexpr uint32_t hi = desc_hi(8 * 128, 128);
    int st = 0;
    uint32_t phase = 0;
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[tb], tb_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + tb * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full[st], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = desc_lo(smem_u32(smem + st * Cfg::STAGE));
          const uint32_t b_lo = desc_lo(smem_u32(smem + st * Cfg::STAGE + Cfg::A_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 4 x 32 bytes of K per stage (UMMA_K = 16 bf16 / 32 e4m3)
            const uint32_t acc = (ks | k) != 0 ? 1u : 0u;
            if (FP8) umma_f8_lh(d_tmem, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, acc);
            else umma_bf16_lh(d_tmem, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, acc);
          }
          umma_commit(&empty[st]);
          if (ks == ksteps - 1) umma_commit(&tfull[tb]);
        }
        __syncwarp();
        if (++st == Cfg::NSTAGE) { st = 0; phase ^= 1u; }
      }
      if (++tb == 2) { tb = 0; tb_phase ^= 1u; }
    }
  } else {
    // ===== epilogue warps 0..3: warp w owns TMEM lanes 32w .. 32w+31 (= tile rows) =====
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float alpha = 1.0f;
    if (FP8 && a.scale_a) alpha = (*a.scale_a) * (*a.scale_b);
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mb = t / a.n_tiles, nb = t - mb * a.n_tiles;
      const int64_t m = (int64_t)mb * 128 + warp * 32 + lane;
      bool valid;
      int64_t orow;
      if (a.padded) {
        const int per = a.Hp * a.Wp;
        const int b = (int)(m / per);
        const int rem = (int)(m - (int64_t)b * per);
        const int hp = rem / a.Wp, wp = rem - hp * a.Wp;
        valid = b < a.Bn && hp >= 1 && hp <= a.H && wp >= 1 && wp <= a.W;
        orow = ((int64_t)b * a.H + (hp - 1)) * a.W + (wp - 1);
      } else {
        valid = m < a.m_rows;
        orow = m;
      }
      mbar_wait(&tfull[tb], tb_phase);
      tc_fence_after();
      __nv_bfloat16* dst = a.C + orow * a.n_total + nb * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + tb * BN + ch * 32, v);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            const uint4 pk = make_uint4(pack_bf16x2(v[c] * alpha, v[c + 1] * alpha), pack_bf16x2(v[c + 2] * alpha, v[c + 3] * alpha),
                                        pack_bf16x2(v[c + 4] * alpha, v[c + 5] * alpha), pack_bf16x2(v[c + 6] * alpha, v[c + 7] * alpha));
            *reinterpret_cast<uint4*>(dst + ch * 32 + c) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[tb]);
      if (++tb == 2) { tb = 0; tb_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool FP8>
void launch_kmajor(const void* A, const void* Bm, const GemmArgs& a, int64_t a_rows, int K, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  const int es = FP8 ? 1 : 2;
  const int ke = FP8 ? 128 : 64;
  const CUtensorMap tmA = map2d(A, es, (uint64_t)K, (uint64_t)a_rows, (uint64_t)K * es, ke, 128);
  const CUtensorMap tmB = map2d(Bm, es, (uint64_t)K, (uint64_t)a.taps * a.n_total, (uint64_t)K * es, ke, BN);
  auto kern = kmajor_gemm_kernel<BN, FP8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, 192, Cfg::SMEM, st>>>(tmA, tmB, a);
  hefl::cuda::note_launch();
}

// ------------------------------------------------------------------------------------------------------------
// Block-scaled e4m3 GEMM (kind::mxf8f6f4.block_scale): one UE8M0 scale per row and per 32 elements of K
// ------------------------------------------------------------------------------------------------------------
// C[M, N] = sum_k (A[m, k] 2^(sa[m, k/32] - 127)) (B[n, k] 2^(sb[n, k/32] - 127)). A, B: e4m3 bytes, K-major.
// Scale factors travel in the layout the tensor core wants: for every (128-row block, group of four 32-element
// k-blocks) one 512-byte tile [lane l = 0..31][row quarter i = 0..3][k-block kb = 0..3] holding the scale of row
// 32*i + l -- a stage brings its two tiles in with a bulk copy, `tcgen05.cp.32x128b.warpx4` replicates each into
// four TMEM columns, and the instruction descriptor of MMA kb selects byte kb of those columns (a_sf_id / b_sf_id).
struct MxArgs {
  int m_tiles, n_tiles, kgroups;   // 128-row tiles, BN-column tiles, K / 128
  int n_total;
  int64_t m_rows;
  __nv_bfloat16* C;
  const uint8_t* sfa;              // [m_tiles][kgroups][512]
  const uint8_t* sfb;              // [N / 128][kgroups][512]
  uint32_t idesc;                  // instruction descriptor without the scale-factor ids
};

constexpr int kMxStage = 128 * 128 + 128 * 128 + 1024;    // A tile, B tile, two scale tiles
constexpr int kMxStages = 5;
constexpr int kMxSmem = kMxStages * kMxStage + 256 + 1024;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 32 rows x 16 bytes, no swizzle: four 8-row core matrices 128 bytes apart
__device__ __forceinline__ uint64_t sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;     // SBO
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void utccp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                          uint32_t idesc, uint32_t tsfa, uint32_t tsfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], da, db, %5, [%6], [%7], p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(tsfa), "r"(tsfb), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(192, 1)
mx_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MxArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMxStages * kMxStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMxStages;
  uint64_t* tfull = empty + kMxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int BN = 128;
  constexpr uint32_t SFA_COL = 2 * BN, SFB_COL = 2 * BN + 4;
  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kMxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = a.m_tiles * a.n_tiles;

  if (warp == 4) {
    if (elect_one()) {
      int st = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mb = t / a.n_tiles, nb = t - mb * a.n_tiles;
        for (int kg = 0; kg < a.kgroups; ++kg) {
          mbar_wait(&empty[st], phase ^ 1u);
          uint8_t* s = smem + st * kMxStage;
          mbar_expect_tx(&full[st], kMxStage);
          tma_load_2d(s, &tmA, kg * 128, mb * 128, &full[st]);
          tma_load_2d(s + 16384, &tmB, kg * 128, nb * BN, &full[st]);
          bulk_load_1d(s + 32768, a.sfa + ((size_t)mb * a.kgroups + kg) * 512, 512, &full[st]);
          bulk_load_1d(s + 32768 + 512, a.sfb + ((size_t)nb * a.kgroups + kg) * 512, 512, &full[st]);
          if (++st == kMxStages) { st = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t hi = desc_hi(8 * 128, 128);
    int st = 0;
    uint32_t phase = 0;
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[tb], tb_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + tb * BN;
      for (int kg = 0; kg < a.kgroups; ++kg) {
        mbar_wait(&full[st], phase);
        tc_fence_after();
        if (elect_one()) {
          uint8_t* s = smem + st * kMxStage;
          const uint32_t a_lo = desc_lo(smem_u32(s));
          const uint32_t b_lo = desc_lo(smem_u32(s + 16384));
          utccp_32x128b_warpx4(tmem_base + SFA_COL, sf_desc(smem_u32(s + 32768)));
          utccp_32x128b_warpx4(tmem_base + SFB_COL, sf_desc(smem_u32(s + 32768 + 512)));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t idesc = a.idesc | ((uint32_t)k << 4) | ((uint32_t)k << 29);      // b_sf_id, a_sf_id
            umma_mxf8(d_tmem, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, tmem_base + SFA_COL, tmem_base + SFB_COL,
                      (kg | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          if (kg == a.kgroups - 1) umma_commit(&tfull[tb]);
        }
        __syncwarp();
        if (++st == kMxStages) { st = 0; phase ^= 1u; }
      }
      if (++tb == 2) { tb = 0; tb_phase ^= 1u; }
    }
  } else {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mb = t / a.n_tiles, nb = t - mb * a.n_tiles;
      const int64_t m = (int64_t)mb * 128 + warp * 32 + lane;
      const bool valid = m < a.m_rows;
      mbar_wait(&tfull[tb], tb_phase);
      tc_fence_after();
      __nv_bfloat16* dst = a.C + m * a.n_total + nb * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + tb * BN + ch * 32, v);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8)
            *reinterpret_cast<uint4*>(dst + ch * 32 + c) = make_uint4(pack_bf16x2(v[c], v[c + 1]), pack_bf16x2(v[c + 2], v[c + 3]),
                                                                       pack_bf16x2(v[c + 4], v[c + 5]), pack_bf16x2(v[c + 6], v[c + 7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[tb]);
      if (++tb == 2) { tb = 0; tb_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: MN-major operands (pixel axis = K), up to 3 taps per CTA, split over pixel ranges
// ------------------------------------------------------------------------------------------------------------
struct WgArgs {
  int co_tiles, ci_tiles;     // 128-wide tiles (64-wide when CT = 64)
  int groups;                 // tap groups (1 for a 1x1 convolution, 3 filter rows for 3x3)
  int taps_per;               // taps in a group (1 or 3)
  int gshift[3];              // row shift of X for the FIRST tap of group g (the others follow at +1, +2)
  int ksplit, kchunks;        // pixel-range splits, 64-row chunks in total
  int Co, Ci;
  float* dW;                  // [taps][Co][Ci] fp32, accumulated with RED
};

// The dY operand (MMA "A", M = output channels) is always 128 channels = two 64-channel atoms; when Co = 64 the
// second atom lies outside the tensor and TMA fills it with zeros. CT = input-channel tile (MMA N): 64 or 128.
template <int CT>
struct WgCfg {
  static constexpr int ATOMS = CT / 64;                       // 64-channel (128-byte) MN atoms of the X operand
  static constexpr int A_ATOM = 64 * 128;                     // dY: 64 k-rows x 128 B
  static constexpr int B_ROWS = 72;                           // X: 64 + 2 shifted rows, rounded to 8
  static constexpr int B_ATOM = B_ROWS * 128;
  static constexpr int A_BYTES = 2 * A_ATOM;
  static constexpr int B_BYTES = ATOMS * B_ATOM;
  static constexpr int STAGE = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int NSTAGE = CT == 128 ? 5 : 7;
  static constexpr int SMEM = NSTAGE * STAGE + 256 + 1024;
  static constexpr int TMEM_COLS = 3 * CT <= 256 ? 256 : 512;
};

template <int CT>
__global__ void __launch_bounds__(192, 1)
wgrad_mn_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgArgs a) {
  using Cfg = WgCfg<CT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::NSTAGE;
  uint64_t* tfull = empty + Cfg::NSTAGE;          // [1]
  uint64_t* tempty = tfull + 1;                   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = a.co_tiles * a.ci_tiles * a.groups * a.ksplit;
  const int per_split = (a.kchunks + a.ksplit - 1) / a.ksplit;

  auto decode = [&](int u, int& co0, int& ci0, int& g, int& c0, int& c1) {
    const int ks = u % a.ksplit;
    int r = u / a.ksplit;
    g = r % a.groups;
    r /= a.groups;
    ci0 = (r % a.ci_tiles) * CT;
    co0 = (r / a.ci_tiles) * 128;
    c0 = ks * per_split;
    c1 = min(a.kchunks, c0 + per_split);
  };

  if (warp == 4) {
    if (elect_one()) {
      int st = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int co0, ci0, g, c0, c1;
        decode(u, co0, ci0, g, c0, c1);
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&empty[st], phase ^ 1u);
          uint8_t* s = smem + st * Cfg::STAGE;
          mbar_expect_tx(&full[st], Cfg::A_BYTES + Cfg::B_BYTES);
          for (int at = 0; at < 2; ++at) tma_load_2d(s + at * Cfg::A_ATOM, &tmDY, co0 + at * 64, c * 64, &full[st]);
          for (int at = 0; at < Cfg::ATOMS; ++at)
            tma_load_2d(s + Cfg::A_BYTES + at * Cfg::B_ATOM, &tmX, ci0 + at * 64, c * 64 + a.gshift[g], &full[st]);
          if (++st == Cfg::NSTAGE) { st = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = make_idesc_bf16(128, CT, 1, 1);
    constexpr uint32_t hi = desc_hi(8 * 128, 128);
    constexpr uint32_t a_lbo = (uint32_t)(Cfg::A_ATOM >> 4) << 16;
    constexpr uint32_t b_lbo = (uint32_t)(Cfg::B_ATOM >> 4) << 16;
    int st = 0;
    uint32_t phase = 0, tphase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int co0, ci0, g, c0, c1;
      decode(u, co0, ci0, g, c0, c1);
      mbar_wait(tempty, tphase ^ 1u);
      tc_fence_after();
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&full[st], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = desc_lo(smem_u32(smem + st * Cfg::STAGE)) + a_lbo;
          const uint32_t b_lo = desc_lo(smem_u32(smem + st * Cfg::STAGE + Cfg::A_BYTES)) + b_lbo;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // 4 x 16 pixel rows per chunk
            const uint32_t acc = (c != c0 || k != 0) ? 1u : 0u;
            for (int j = 0; j < a.taps_per; ++j)
              umma_bf16_lh(tmem_base + j * CT, a_lo + ((k * 16 * 128) >> 4), hi, b_lo + (((k * 16 + j) * 128) >> 4), hi,
                           idesc, acc);
          }
          umma_commit(&empty[st]);
          if (c == c1 - 1) umma_commit(tfull);
        }
        __syncwarp();
        if (++st == Cfg::NSTAGE) { st = 0; phase ^= 1u; }
      }
      if (c1 > c0) tphase ^= 1u;
    }
  } else {
    // epilogue: TMEM lane = output channel (row of dW), 4 warps x 32 lanes = 128 rows
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t tphase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int co0, ci0, g, c0, c1;
      decode(u, co0, ci0, g, c0, c1);
      if (c1 <= c0) continue;
      mbar_wait(tfull, tphase);
      tc_fence_after();
      const int row = warp * 32 + lane;
      const bool valid = co0 + row < a.Co;
#pragma unroll 1
      for (int j = 0; j < a.taps_per; ++j) {
        const int tap = g * a.taps_per + j;
#pragma unroll 1
        for (int ch = 0; ch < CT / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + j * CT + ch * 32, v);
          if (valid) {
            float* dst = a.dW + ((size_t)tap * a.Co + co0 + row) * a.Ci + ci0 + ch * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[c]), "f"(v[c + 1]),
                           "f"(v[c + 2]), "f"(v[c + 3])
                           : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      tphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int CT>
void launch_wgrad_mn(const void* DY, const void* X, const WgArgs& a, int64_t rows, cudaStream_t st) {
  using Cfg = WgCfg<CT>;
  const CUtensorMap tmDY = map2d(DY, 2, (uint64_t)a.Co, (uint64_t)rows, (uint64_t)a.Co * 2, 64, 64);
  const CUtensorMap tmX = map2d(X, 2, (uint64_t)a.Ci, (uint64_t)rows, (uint64_t)a.Ci * 2, 64, Cfg::B_ROWS);
  auto kern = wgrad_mn_kernel<CT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  const int units = a.co_tiles * a.ci_tiles * a.groups * a.ksplit;
  const int grid = units < sm_count() ? units : sm_count();
  kern<<<grid, 192, Cfg::SMEM, st>>>(tmDY, tmX, a);
  hefl::cuda::note_launch();
}

// ------------------------------------------------------------------------------------------------------------
// data movement around the GEMMs: one pass each (ATen needs a fill + a strided copy, resp. cast + permute + flip)
// ------------------------------------------------------------------------------------------------------------
// x [B][H][W][C] -> xp [B][H+2][W+2][C] with a zero border; 16-byte vectors, C % 8 == 0
__global__ void pad_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ xp, int B, int H, int W, int C8) {
  const int Hp = H + 2, Wp = W + 2;
  const int64_t total = (int64_t)B * Hp * Wp * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    int64_t p = i / C8;
    const int wp = (int)(p % Wp);
    p /= Wp;
    const int hp = (int)(p % Hp);
    const int b = (int)(p / Hp);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hp >= 1 && hp <= H && wp >= 1 && wp <= W) v = x[(((int64_t)b * H + hp - 1) * W + wp - 1) * C8 + c];
    xp[i] = v;
  }
}

// w fp32 [Co][Ci][3][3] -> wt bf16 [tap][Co][Ci] (forward) and wd bf16 [tap][Ci][Co] with the filter rotated by 180
// degrees (dgrad); for k = 1: wt [Co][Ci] and wd = its transpose [Ci][Co]. One 32 x 32 (co, ci) tile per block and
// tap, transposed through shared memory so that both outputs are written with unit stride.
__global__ void __launch_bounds__(256)
conv_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ wd,
                        int Co, int Ci, int kk) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int co = co0 + ty + 8 * r, ci = ci0 + tx;
    if (co < Co && ci < Ci) {
      const __nv_bfloat16 v = __float2bfloat16(w[((int64_t)co * Ci + ci) * kk + tap]);
      tile[ty + 8 * r][tx] = v;
      wt[((int64_t)tap * Co + co) * Ci + ci] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ci = ci0 + ty + 8 * r, co = co0 + tx;
    if (co < Co && ci < Ci) wd[((int64_t)(kk - 1 - tap) * Ci + ci) * Co + co] = tile[tx][ty + 8 * r];
  }
}

// dW fp32 [tap][Co][Ci] -> grad fp32 [Co][Ci][kh][kw]
__global__ void conv_wgrad_unpack_kernel(const float* __restrict__ dw, float* __restrict__ g, int Co, int Ci, int kk) {
  const int64_t total = (int64_t)kk * Co * Ci;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk);
    const int64_t cc = i / kk;                     // co * Ci + ci
    g[i] = dw[(int64_t)tap * Co * Ci + cc];
  }
}

}  // namespace

void pad_nhwc(const void* x, void* xp, int B, int H, int W, int C, cudaStream_t st) {
  if (C % 8) throw std::runtime_error("pad_nhwc: C must be a multiple of 8");
  const int64_t total = (int64_t)B * (H + 2) * (W + 2) * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pad_nhwc_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(xp), B, H, W, C / 8);
  hefl::cuda::note_launch();
}

void conv_weight_prep(const float* w, void* wt, void* wd, int Co, int Ci, int kk, cudaStream_t st) {
  dim3 grid((Ci + 31) / 32, (Co + 31) / 32, kk);
  conv_weight_prep_kernel<<<grid, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(wt), reinterpret_cast<__nv_bfloat16*>(wd), Co, Ci, kk);
  hefl::cuda::note_launch();
}

void conv_wgrad_unpack(const float* dw, float* g, int Co, int Ci, int kk, cudaStream_t st) {
  const int64_t total = (int64_t)kk * Co * Ci;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv_wgrad_unpack_kernel<<<blocks, 256, 0, st>>>(dw, g, Co, Ci, kk);
  hefl::cuda::note_launch();
}

// Block-scaled e4m3 GEMM; sfa / sfb in the tiled layout described at MxArgs. idesc_variant selects the encoding of
// the M field of the block-scaled instruction descriptor (0: M >> 7 at bit 27, the documented layout; 1: M >> 4 at bit 24).
void gemm_mxfp8(const void* A, const void* Bm, const void* sfa, const void* sfb, void* C, int64_t M, int N, int K,
                int idesc_variant, cudaStream_t st) {
  if (K % 128 != 0 || N % 128 != 0) throw std::runtime_error("gemm_mxfp8: K and N must be multiples of 128");
  MxArgs a{};
  a.m_tiles = (int)((M + 127) / 128);
  a.n_tiles = N / 128;
  a.kgroups = K / 128;
  a.n_total = N;
  a.m_rows = M;
  a.C = reinterpret_cast<__nv_bfloat16*>(C);
  a.sfa = reinterpret_cast<const uint8_t*>(sfa);
  a.sfb = reinterpret_cast<const uint8_t*>(sfb);
  // e4m3 x e4m3 (formats 0), K-major both, N >> 3 at bit 17, UE8M0 scales (bit 23)
  uint32_t idesc = ((uint32_t)(128 >> 3) << 17) | (1u << 23);
  idesc |= idesc_variant == 0 ? ((uint32_t)(128 >> 7) << 27) : ((uint32_t)(128 >> 4) << 24);
  a.idesc = idesc;
  const CUtensorMap tmA = map2d(A, 1, (uint64_t)K, (uint64_t)M, (uint64_t)K, 128, 128);
  const CUtensorMap tmB = map2d(Bm, 1, (uint64_t)K, (uint64_t)N, (uint64_t)K, 128, 128);
  cudaFuncSetAttribute(mx_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmem);
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  mx_gemm_kernel<<<grid, 192, kMxSmem, st>>>(tmA, tmB, a);
  hefl::cuda::note_launch();
}

// C[rows_out, N] (bf16) = sum_taps A[m + shift_t, K] . B[t*N + n, K]; see GemmArgs. fp8: A, B are e4m3 bytes.
void gemm_taps(const void* A, const void* Bm, void* C, int64_t a_rows, int N, int K, int taps, const int* shifts, int padded,
               int Bn, int H, int W, bool fp8, const float* scale_a, const float* scale_b, cudaStream_t st) {
  const int kbytes = fp8 ? K : 2 * K;
  if (kbytes % 128 != 0) throw std::runtime_error("gemm_taps: K must cover whole 128-byte blocks");
  if (N % 64 != 0) throw std::runtime_error("gemm_taps: N must be a multiple of 64");
  if (taps < 1 || taps > 9) throw std::runtime_error("gemm_taps: 1..9 taps");
  GemmArgs a{};
  a.m_tiles = (int)((a_rows + 127) / 128);
  a.kblocks = kbytes / 128;
  a.taps = taps;
  for (int t = 0; t < taps; ++t) a.shift[t] = shifts ? shifts[t] : 0;
  a.n_total = N;
  a.m_rows = a_rows;
  a.padded = padded;
  a.Bn = Bn; a.H = H; a.W = W; a.Hp = H + 2; a.Wp = W + 2;
  a.C = reinterpret_cast<__nv_bfloat16*>(C);
  a.scale_a = scale_a;
  a.scale_b = scale_b;
  const bool wide = N % 128 == 0;
  a.n_tiles = wide ? N / 128 : N / 64;
  if (wide) {
    if (fp8) launch_kmajor<128, true>(A, Bm, a, a_rows, K, st);
    else launch_kmajor<128, false>(A, Bm, a, a_rows, K, st);
  } else {
    if (fp8) launch_kmajor<64, true>(A, Bm, a, a_rows, K, st);
    else launch_kmajor<64, false>(A, Bm, a, a_rows, K, st);
  }
}

// dW[taps][Co][Ci] (fp32, must be zeroed by the caller) += dY[rows, Co]^T . X[rows + shift, Ci]; taps = 1 or 9
// (3x3 on a padded grid of row length Wp).
void wgrad_taps(const void* DY, const void* X, float* dW, int64_t rows, int Co, int Ci, int taps, int Wp, cudaStream_t st) {
  if (Co % 64 != 0 || Ci % 64 != 0) throw std::runtime_error("wgrad_taps: channel counts must be multiples of 64");
  if (taps != 1 && taps != 9) throw std::runtime_error("wgrad_taps: 1 or 9 taps");
  WgArgs a{};
  const bool wide = Ci % 128 == 0;
  const int ct = wide ? 128 : 64;
  a.co_tiles = (Co + 127) / 128;
  a.ci_tiles = Ci / ct;
  a.groups = taps == 9 ? 3 : 1;
  a.taps_per = taps == 9 ? 3 : 1;
  for (int g = 0; g < a.groups; ++g) a.gshift[g] = taps == 9 ? (g - 1) * Wp - 1 : 0;
  a.kchunks = (int)((rows + 63) / 64);
  const int base_units = a.co_tiles * a.ci_tiles * a.groups;
  int ks = (sm_count() + base_units - 1) / base_units;          // about one wave of CTAs: longer K runs per unit,
                                                                // fewer REDs onto the same dW words
  if (ks < 1) ks = 1;
  if (ks > a.kchunks) ks = a.kchunks;
  a.ksplit = ks;
  a.Co = Co; a.Ci = Ci;
  a.dW = dW;
  if (wide) launch_wgrad_mn<128>(DY, X, a, rows, st);
  else launch_wgrad_mn<64>(DY, X, a, rows, st);
}

}  // namespace nn
}  // namespace hefl
