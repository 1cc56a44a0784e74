// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Activations are NHWC bf16 viewed as a 2-D matrix [P = B*H*W pixels, C channels]. A 3x3 valid
// convolution is then 9 "tap" GEMMs whose A operand is the SAME matrix shifted by
// r*W + s rows:   Y[m, :] = sum_{r,s} X[m + r*W + s, :] * W[r,s]^T   (m on the input grid; rows
// with h >= H-2 or w >= W-2 are garbage and masked). Each tap's A tile is therefore one plain
// 2-D TMA box — no im2col buffer, no gather.
//
//   conv_fwd_pool : per CTA tile = 2 image rows x 128 columns (two TMEM accumulators) so that the
//                   epilogue does bias + ReLU + 2x2 max-pool + argmax straight out of TMEM (vertical
//                   max in-thread across the two accumulators, horizontal max by one shuffle) and
//                   writes only the pooled tensor (4x less HBM traffic than the conv output).
//   conv_dgrad    : dX[m,:] = sum_taps dY[m - off, :] * W[r,s]   (same kernel, negative offsets,
//                   plain bf16 store).
//   conv_wgrad    : dW[r,s] = sum_m X[m+off,:]^T dY[m,:]. Both operands are consumed in their
//                   native [P, C] layout as MN-major UMMA operands (pixels = K), so no transposed
//                   copies exist; taps are stacked along M (128/C tap atoms per MMA group), split-K
//                   over pixels, fp32 vector-RED into a [9*C+1, Co] buffer whose extra row is the
//                   bias gradient (an all-ones tap atom).
//
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> global). smem ring of halo tiles with full/empty
// mbarriers; two TMEM tile buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
// Weights for all 9 taps stay resident in shared memory for the CTA's lifetime.
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
// G1: halo-tile tap-GEMM kernel (forward + pool epilogue, or dgrad with plain store)
//
// Measured on B200 (bench/probe_shift.py): a K-major swizzled tile written by TMA can be read by
// tcgen05.mma starting at ANY row (descriptor start address + rows*row_bytes, base_offset 0) —
// the swizzle is a function of absolute shared-memory address bits. So each input row segment
// is loaded ONCE per tile and all 9 taps address shifted windows of it:
//   forward : tile = 2 image rows x 128 columns; 4 halo segments (rows h..h+3) of 130 pixels;
//             accumulator j (row h+j), tap (r,s) reads segment j+r at row offset s.
//   dgrad   : tile = 128 consecutive pixels; 3 segments starting at m0 - r*W - 2; tap (r,s)
//             reads segment r at row offset 2 - s.
// L2->SM traffic drops from 9 tap tiles to ~2 (fwd) / ~3 (dgrad) tiles per output tile.
// ------------------------------------------------------------------------------------------
struct TapGemmArgs {
  int B, H, W;          // input grid of the A matrix
  int Hp, Wp;           // pooled output grid (POOL)
  int tiles_w;          // column tiles per row pair (POOL)
  int num_tiles;
  int P;                // B*H*W
  int co_total;         // output channels of the layer (CO per CTA column block = blockIdx.y)
  const float* bias;    // [CO] (POOL)
  __nv_bfloat16* out;   // POOL: [B,Hp,Wp,CO]; else [P,CO]
  uint8_t* argmax;      // POOL, may be null
};

template <int CK, int CO, bool POOL>
struct TapGemmCfg {
  static constexpr int KB = CK < 64 ? CK : 64;            // channels per k-block (one swizzle atom)
  static constexpr int NKB = CK / KB;
  static constexpr int ROW_BYTES = KB * 2;
  static constexpr int NSEG = POOL ? 4 : 3;
  static constexpr int NACC = POOL ? 2 : 1;
  static constexpr int SEG_ROWS = 136;                    // 130 used, multiple of 8
  static constexpr int SEG_BYTES = SEG_ROWS * ROW_BYTES;
  static constexpr int HALO = NKB * NSEG * SEG_BYTES;     // one tile's input
  static constexpr int W_SUB = CO * ROW_BYTES;
  static constexpr int W_TAP = W_SUB * NKB;
  static constexpr int W_BYTES = 9 * W_TAP;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 225 * 1024;
  static constexpr int NBUF_RAW = (BUDGET - W_BYTES - BAR_BYTES - 1024) / HALO;
  static constexpr int NBUF = NBUF_RAW > 3 ? 3 : NBUF_RAW;
  static constexpr int SMEM = W_BYTES + NBUF * HALO + BAR_BYTES + 1024;
  static constexpr int ACC_COLS = 2 * NACC * CO;          // two tile buffers
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  static_assert(NBUF >= 1, "halo tile does not fit in shared memory");
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + Cfg::NBUF * Cfg::HALO);
  uint64_t* full = bars;                       // [NBUF]
  uint64_t* empty = bars + Cfg::NBUF;          // [NBUF]
  uint64_t* wfull = bars + 2 * Cfg::NBUF;      // [1]
  uint64_t* tfull = wfull + 1;                 // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < Cfg::NBUF; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
          tma_load_2d(sW + tap * Cfg::W_TAP + kb * Cfg::W_SUB, &tmW, kb * Cfg::KB,
                      tap * a.co_total + blockIdx.y * CO, wfull);
      int buf = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        int base, seg_stride;
        if (POOL) {
          const int per_img = a.Hp * a.tiles_w;
          const int b = t / per_img;
          const int rem = t - b * per_img;
          const int hp = rem / a.tiles_w;
          const int tw = rem - hp * a.tiles_w;
          base = (b * a.H + 2 * hp) * a.W + tw * 128;
          seg_stride = a.W;                      // segment j = image row h + j
        } else {
          base = t * 128 - 2;
          seg_stride = -a.W;                     // segment r starts at m0 - r*W - 2
        }
        mbar_wait(&empty[buf], phase ^ 1);
        mbar_expect_tx(&full[buf], Cfg::HALO);
        uint8_t* dst = sA + buf * Cfg::HALO;
        for (int kb = 0; kb < Cfg::NKB; ++kb)
          for (int sg = 0; sg < Cfg::NSEG; ++sg)
            tma_load_2d(dst + (kb * Cfg::NSEG + sg) * Cfg::SEG_BYTES, &tmA, kb * Cfg::KB, base + sg * seg_stride,
                        &full[buf]);
        if (++buf == Cfg::NBUF) { buf = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(128, CO);
    mbar_wait(wfull, 0);
    tc_fence_after();
    int buf = 0;
    uint32_t phase = 0;
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[tb], tb_phase ^ 1);
      mbar_wait(&full[buf], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(sA + buf * Cfg::HALO);
        const uint32_t w_addr = smem_u32(sW);
#pragma unroll
        for (int j = 0; j < Cfg::NACC; ++j) {
          const uint32_t d_tmem = tmem_base + (tb * Cfg::NACC + j) * CO;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            const int seg = POOL ? (j + r) : r;
            const int shift = POOL ? s : (2 - s);
#pragma unroll
            for (int kb = 0; kb < Cfg::NKB; ++kb) {
#pragma unroll
              for (int k = 0; k < Cfg::KB / 16; ++k) {
                const uint64_t ad = make_kmajor_desc(
                    a_addr + (kb * Cfg::NSEG + seg) * Cfg::SEG_BYTES + shift * Cfg::ROW_BYTES + k * 32, Cfg::ROW_BYTES);
                const uint64_t bd = make_kmajor_desc(w_addr + tap * Cfg::W_TAP + kb * Cfg::W_SUB + k * 32, Cfg::ROW_BYTES);
                umma_bf16(d_tmem, ad, bd, idesc, (tap | kb | k) != 0 ? 1u : 0u);
              }
            }
          }
        }
        umma_commit(&empty[buf]);
        umma_commit(&tfull[tb]);
      }
      __syncwarp();
      if (++buf == Cfg::NBUF) { buf = 0; phase ^= 1; }
      tb ^= 1;
      if (tb == 0) tb_phase ^= 1;
    }
  } else {
    // ===== epilogue warps (2..5) =====
    const int qd = warp & 3;                       // TMEM lane quadrant this warp may read
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    int tb = 0;
    uint32_t tb_phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      mbar_wait(&tfull[tb], tb_phase);
      tc_fence_after();
      if (POOL) {
        const int per_img = a.Hp * a.tiles_w;
        const int b = t / per_img;
        const int rem = t - b * per_img;
        const int hp = rem / a.tiles_w;
        const int tw = rem - hp * a.tiles_w;
        const int col = tw * 128 + qd * 32 + lane;           // conv-output column of this thread
        const int wp = col >> 1;
        const bool writer = ((lane & 1) == 0) && wp < a.Wp;
#pragma unroll 1
        for (int ch = 0; ch < CO / 32; ++ch) {
          float v0[32], v1[32];
          tmem_ld32(tmem_base + lane_base + (tb * 2 + 0) * CO + ch * 32, v0);
          tmem_ld32(tmem_base + lane_base + (tb * 2 + 1) * CO + ch * 32, v1);
          uint32_t vbits = 0;                                 // 1 = lower image row wins
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float bsv = __ldg(a.bias + ch * 32 + c);
            float x0 = v0[c] + bsv, x1 = v1[c] + bsv;
            x0 = x0 > 0.f ? x0 : 0.f;
            x1 = x1 > 0.f ? x1 : 0.f;
            if (x1 > x0) { x0 = x1; vbits |= 1u << c; }
            v0[c] = x0;
          }
          const uint32_t pvbits = __shfl_xor_sync(0xffffffffu, vbits, 1);
          uint32_t packed[16];
          uint32_t idx4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) idx4[i] = 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float x = v0[c];
            const float o = __shfl_xor_sync(0xffffffffu, x, 1);
            uint32_t id = ((vbits >> c) & 1u) * 2u;
            if (o > x) { x = o; id = ((pvbits >> c) & 1u) * 2u + 1u; }
            const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(x));
            if (c & 1) packed[c >> 1] |= hb << 16; else packed[c >> 1] = hb;
            idx4[c >> 2] |= id << ((c & 3) * 8);
          }
          if (writer) {
            const size_t o = ((size_t)(b * a.Hp + hp) * a.Wp + wp) * CO + ch * 32;
            uint4* dst = reinterpret_cast<uint4*>(a.out + o);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            if (a.argmax) {
              uint4* di = reinterpret_cast<uint4*>(a.argmax + o);
              di[0] = make_uint4(idx4[0], idx4[1], idx4[2], idx4[3]);
              di[1] = make_uint4(idx4[4], idx4[5], idx4[6], idx4[7]);
            }
          }
        }
      } else {
        const int m = t * 128 + qd * 32 + lane;
#pragma unroll 1
        for (int ch = 0; ch < CO / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + tb * CO + ch * 32, v);
          if (m < a.P) {
            uint32_t packed[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const uint32_t lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(v[c]));
              const uint32_t hi = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(v[c + 1]));
              packed[c >> 1] = lo | (hi << 16);
            }
            uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)m * a.co_total + blockIdx.y * CO + ch * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[tb]);
      tb ^= 1;
      if (tb == 0) tb_phase ^= 1;
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
  const CUtensorMap tmA = make_map(A, CK, (uint64_t)a.P, (uint64_t)CK * 2, Cfg::KB, Cfg::SEG_ROWS);
  const CUtensorMap tmW = make_map(Wt, CK, (uint64_t)9 * a.co_total, (uint64_t)CK * 2, Cfg::KB, CO);
  auto kern = tap_gemm_kernel<CK, CO, POOL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  const int ny = a.co_total / CO;
  int gx = a.num_tiles < num_sms() / ny ? a.num_tiles : num_sms() / ny;
  if (gx < 1) gx = 1;
  dim3 grid(gx, ny);
  kern<<<grid, 192, Cfg::SMEM, st>>>(tmA, tmW, a);
  hefl::cuda::note_launch();
}

void conv_fwd_pool(const void* X, const void* Wf, const float* bias, void* out, uint8_t* argmax, int B, int H,
                   int W, int CK, int CO, cudaStream_t st) {
  TapGemmArgs a{};
  a.B = B; a.H = H; a.W = W;
  a.Hp = (H - 2) / 2; a.Wp = (W - 2) / 2;
  a.tiles_w = (2 * a.Wp + 127) / 128;
  a.num_tiles = B * a.Hp * a.tiles_w;
  a.P = B * H * W;
  a.co_total = CO;
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
  a.co_total = CO;
  a.out = reinterpret_cast<__nv_bfloat16*>(dX);
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(dY);
  const auto* w = reinterpret_cast<const __nv_bfloat16*>(Wd);
  if (CK == 32 && CO == 32) launch_tap_gemm<32, 32, false>(x, w, a, st);
  else if (CK == 64 && CO == 32) launch_tap_gemm<64, 32, false>(x, w, a, st);
  else if (CK == 64 && CO == 64) launch_tap_gemm<64, 64, false>(x, w, a, st);
  else if (CK == 128 && CO == 64) launch_tap_gemm<128, 32, false>(x, w, a, st);   // two column blocks
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

// ------------------------------------------------------------------------------------------
// Probe: can a K-major swizzled A tile be consumed starting at an arbitrary ROW offset (with the
// descriptor's base_offset field carrying the swizzle phase)? Used to validate halo-tile reuse.
// ------------------------------------------------------------------------------------------
namespace hefl {
namespace nn {
using namespace hefl::tc;

template <int CK>
__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        float* out, int shift_rows, int mode) {
  constexpr int RB = CK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                      // 144 rows
  uint8_t* sB = smem + 144 * 128;          // 32 rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 144 * RB + 32 * RB);
    tma_load_2d(sA, &tmA, 0, 0, bar);
    tma_load_2d(sB, &tmB, 0, 0, bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + shift_rows * RB;
    for (int k = 0; k < CK / 16; ++k) {
      uint64_t ad = make_kmajor_desc(a0 + k * 32, RB);
      if (mode == 1) ad |= (uint64_t)((a0 >> 7) & 7) << 49;   // base_offset = row phase in the 1024-B pattern
      const uint64_t bd = make_kmajor_desc(smem_u32(sB) + k * 32, RB);
      umma_bf16(tmem, ad, bd, make_idesc_bf16(128, 32), k ? 1u : 0u);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = v[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tmem); }
}

void umma_shift_probe(const void* A, const void* Bm, float* out, int CK, int shift_rows, int mode, cudaStream_t st) {
  const CUtensorMap tmA = make_map(A, CK, 144, (uint64_t)CK * 2, CK, 144);
  const CUtensorMap tmB = make_map(Bm, CK, 32, (uint64_t)CK * 2, CK, 32);
  const int smem = 144 * 128 + 32 * 128 + 64 + 1024;
  if (CK == 64) {
    cudaFuncSetAttribute(umma_shift_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_shift_probe_kernel<64><<<1, 128, smem, st>>>(tmA, tmB, out, shift_rows, mode);
  } else if (CK == 32) {
    cudaFuncSetAttribute(umma_shift_probe_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_shift_probe_kernel<32><<<1, 128, smem, st>>>(tmA, tmB, out, shift_rows, mode);
  } else {
    cudaFuncSetAttribute(umma_shift_probe_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_shift_probe_kernel<16><<<1, 128, smem, st>>>(tmA, tmB, out, shift_rows, mode);
  }
}

}  // namespace nn
}  // namespace hefl
