// Fused ciphertext all-reduce: coefficient-wise sum over ranks *modulo each RNS prime*,
// performed inside one kernel that reads/writes peer GPUs' memory over NVLink/NVSwitch.
// No NCCL call on this path (replaces the file-drop aggregation
// loop FLPyfhelin.py:372-381).
//
// Buffer layout on every rank: [C][2][L][N] u64 residues (< q_l < 2^61), symmetric address
// space (peer r's buffer is mapped at bufs[r]). Three algorithms:
//   two_shot  : rank r owns chunk r. It pulls chunk r from all P peers with 16-byte loads,
//               adds, reduces mod q_l in registers and pushes the result into chunk r of all
//               P buffers (in place).
//   one_shot  : every rank pulls everything and writes a private output (latency-optimal for
//               small messages).
//   multimem  : as two_shot but the P-way add is done by the NVSwitch
//               (multimem.ld_reduce.add.u64) and the result is broadcast with multimem.st.
// Cross-GPU synchronisation uses per-block flags in a symmetric signal pad with
// release/acquire CAS at system scope; flags return to 0 so no reset/epoch is needed.
// All spins are bounded: on timeout the kernel records a diagnostic in `status` and exits
// (failure detection).
#include <cstdint>
#include <cuda_runtime.h>

#include "../he/kernels.h"
#include "../he/modarith.h"
#include "comm.h"

namespace hefl {
namespace comm {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Returns false on timeout.
__device__ __forceinline__ bool put_signal(uint32_t* addr, uint64_t deadline) {
  uint32_t spins = 0;
  while (cas_release_sys(addr, 0u, 1u) != 0u) {
    if ((++spins & 0x3FFu) == 0 && globaltimer_ns() > deadline) return false;
  }
  return true;
}
__device__ __forceinline__ bool wait_signal(uint32_t* addr, uint64_t deadline) {
  uint32_t spins = 0;
  while (cas_acquire_sys(addr, 1u, 0u) != 1u) {
    if ((++spins & 0x3FFu) == 0 && globaltimer_ns() > deadline) return false;
  }
  return true;
}

// Block-level barrier across ranks: block b of every rank meets block b of every other rank.
// slot selects one of two flag banks so consecutive barriers in one kernel never alias.
__device__ __forceinline__ bool block_barrier(const AllReduceArgs& a, int slot, uint64_t deadline) {
  __shared__ int ok_flag;
  if (threadIdx.x == 0) ok_flag = 1;
  __syncthreads();
  if (threadIdx.x < (unsigned)a.world) {
    const int peer = threadIdx.x;
    const size_t base = ((size_t)slot * gridDim.x + blockIdx.x) * a.world;
    bool ok = put_signal(a.sigs[peer] + base + a.rank, deadline);
    ok = ok && wait_signal(a.sigs[a.rank] + base + peer, deadline);
    if (!ok) {
      ok_flag = 0;
      if (a.status) atomicExch(a.status, 1u + (uint32_t)peer + ((uint32_t)slot << 8) + ((uint32_t)blockIdx.x << 16));
    }
  }
  __syncthreads();
  return ok_flag != 0;
}

__device__ __forceinline__ ulonglong2 ld16(const uint64_t* p) {
  ulonglong2 v;
  asm volatile("ld.global.relaxed.sys.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st16(uint64_t* p, ulonglong2 v) {
  asm volatile("st.global.relaxed.sys.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(v.x), "l"(v.y) : "memory");
}

__device__ __forceinline__ uint64_t mod_sum(uint64_t s, uint64_t q, uint64_t ratio_hi) {
  const uint64_t qhat = __umul64hi(s, ratio_hi);
  uint64_t r = s - qhat * q;
  if (r >= q) r -= q;
  if (r >= q) r -= q;
  return r;
}

template <int ALGO>  // 0 two_shot, 1 one_shot, 2 multimem
__global__ void __launch_bounds__(512)
allreduce_modq_kernel(const __grid_constant__ AllReduceArgs a) {
  const uint64_t deadline = globaltimer_ns() + a.timeout_ns;
  if (!block_barrier(a, 0, deadline)) return;

  const int P = a.world;
  const int64_t pairs = a.numel >> 1;  // 16-byte units
  int64_t lo = 0, hi = pairs;
  if (ALGO != 1) {  // chunk ownership, aligned to 16 bytes
    const int64_t per = (pairs + P - 1) / P;
    lo = per * a.rank;
    hi = lo + per < pairs ? lo + per : pairs;
    if (lo > pairs) lo = pairs;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int logn = a.logn;

  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    const int64_t e = i << 1;
    const int l = (int)((e >> logn) % a.L);
    const uint64_t q = a.q[l], rh = a.ratio_hi[l];
    ulonglong2 acc;
    if (ALGO == 2) {
      const uint64_t* mc = a.mc + e;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(acc.x) : "l"(mc) : "memory");
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(acc.y) : "l"(mc + 1) : "memory");
    } else {
      ulonglong2 v[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < P) v[p] = ld16(a.bufs[(a.rank + p) % P] + e);  // stagger peers across ranks
      acc = v[0];
#pragma unroll
      for (int p = 1; p < kMaxWorld; ++p)
        if (p < P) { acc.x += v[p].x; acc.y += v[p].y; }
    }
    acc.x = mod_sum(acc.x, q, rh);
    acc.y = mod_sum(acc.y, q, rh);
    if (ALGO == 1) {
      *reinterpret_cast<ulonglong2*>(a.out + e) = acc;
    } else if (ALGO == 2) {
      uint64_t* mc = a.mc + e;
      asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" :: "l"(mc), "l"(acc.x) : "memory");
      asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" :: "l"(mc + 1), "l"(acc.y) : "memory");
    } else {
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < P) st16(a.bufs[(a.rank + p) % P] + e, acc);
    }
  }
  // make this block's peer stores visible before signalling completion
  __threadfence_system();
  block_barrier(a, 1, deadline);
}

void allreduce_modq(const AllReduceArgs& args, int algo, int blocks, int threads, cudaStream_t st) {
  switch (algo) {
    case 0: allreduce_modq_kernel<0><<<blocks, threads, 0, st>>>(args); break;
    case 1: allreduce_modq_kernel<1><<<blocks, threads, 0, st>>>(args); break;
    default: allreduce_modq_kernel<2><<<blocks, threads, 0, st>>>(args); break;
  }
  hefl::cuda::note_launch();
}

// Local K-way modular sum of K buffers living on ONE device (loopback transport on a single
// GPU, and the aggregation step of the file-based compat path).
__global__ void local_sum_modq_kernel(const uint64_t* const* __restrict__ srcs, int K,
                                      uint64_t* __restrict__ out, int64_t numel, int logn, int L,
                                      const uint64_t* __restrict__ consts) {
  const int64_t pairs = numel >> 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i << 1;
    const int l = (int)((e >> logn) % L);
    const uint64_t q = consts[l * 8], rh = consts[l * 8 + 2];
    ulonglong2 acc = make_ulonglong2(0, 0);
    for (int k = 0; k < K; ++k) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(srcs[k] + e);
      acc.x += v.x;
      acc.y += v.y;
      if ((k & 7) == 7) { acc.x = mod_sum(acc.x, q, rh); acc.y = mod_sum(acc.y, q, rh); }
    }
    acc.x = mod_sum(acc.x, q, rh);
    acc.y = mod_sum(acc.y, q, rh);
    *reinterpret_cast<ulonglong2*>(out + e) = acc;
  }
}

void local_sum_modq(const uint64_t* const* srcs_dev, int K, uint64_t* out, int64_t numel, int logn,
                    int L, const uint64_t* consts, cudaStream_t st) {
  int blocks = (int)(((numel >> 1) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  local_sum_modq_kernel<<<blocks, 256, 0, st>>>(srcs_dev, K, out, numel, logn, L, consts);
  hefl::cuda::note_launch();
}

}  // namespace comm
}  // namespace hefl
