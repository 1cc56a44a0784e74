// Fused ciphertext all-reduce launcher arguments (csrc/comm/allreduce_modq.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hefl {
namespace comm {

constexpr int kMaxWorld = 8;
constexpr int kMaxLimbs = 16;

struct AllReduceArgs {
  uint64_t* bufs[kMaxWorld];   // peer-mapped symmetric buffers (bufs[rank] is local)
  uint32_t* sigs[kMaxWorld];   // peer-mapped signal pads (u32 flags, zero at rest)
  uint64_t* mc;                // multicast address of the buffer (multimem algo) or null
  uint64_t* out;               // private output (one_shot algo)
  uint32_t* status;            // local diagnostic word (0 = ok)
  uint64_t q[kMaxLimbs];
  uint64_t ratio_hi[kMaxLimbs];  // floor(2^64 / q)
  int64_t numel;               // u64 words, multiple of 4
  uint64_t timeout_ns;
  int rank, world, L, logn;
  int no_owner;                // rank that owns no chunk (the secret-key holder) or -1
  uint32_t* stats;             // optional: += number of 32-byte peer-load steps issued by this rank
};

struct MaskArgs {
  uint64_t seed[kMaxWorld];    // pair seed shared with peer j
  int sign[kMaxWorld];         // +1 / -1
  int npeers;
  uint32_t round;
  uint64_t q[kMaxLimbs], ratio_lo[kMaxLimbs], ratio_hi[kMaxLimbs];
};
void pairwise_mask(uint64_t* data, int64_t numel, int L, int logn, const MaskArgs& m, cudaStream_t st);
void pairwise_mask_host(uint64_t* data, int64_t numel, int L, int logn, const MaskArgs& m);

// algo: 0 two_shot, 1 one_shot, 2 multimem. Needs 2 * blocks * world u32 flags per pad.
void allreduce_modq(const AllReduceArgs& args, int algo, int blocks, int threads, cudaStream_t st);

void local_sum_modq(const uint64_t* const* srcs_dev, int K, uint64_t* out, int64_t numel, int logn,
                    int L, const uint64_t* consts, cudaStream_t st);

}  // namespace comm
}  // namespace hefl
