// Counter-based per-env random stream (device side).
//
// Spec (must agree bit for bit with oracle/philox.py, which documents it):
//   Philox4x32-10; key = (seed_lo, seed_hi); counter = (block_lo, block_hi, env_id, tag);
//   draw n = word (n & 3) of block (n >> 2);  u32 -> [0,1): (u >> 8) * 2^-24;
//   u32 -> {lo..hi-1}: lo + ((u * (hi-lo)) >> 32).
// It replaces the reference's MT19937 draws (waterworld.py:141-170,360-374; hostage.py:149-177,
// 374-375; pursuit_evade.py:183-186; utils/agent_utils.py:39-45; utils/Controllers.py:16), which
// are consumed in exactly the reference's order.
#pragma once
#include <stdint.h>

namespace madrl {

struct Philox4 {
  uint32_t w[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  Philox4 o;
  o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

// Word `n` of the stream (stateless; use when draw indices are known up front).
__host__ __device__ __forceinline__ uint32_t stream_word(uint64_t seed, uint32_t env_id,
                                                         uint32_t tag, uint64_t n) {
  const uint64_t blk = n >> 2;
  const Philox4 b = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), env_id, tag,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t i = (uint32_t)n & 3u;
  return i == 0 ? b.w[0] : i == 1 ? b.w[1] : i == 2 ? b.w[2] : b.w[3];
}

// Sequential consumer with a one-block cache (serial draw chains: resets, respawn loops).
struct SeqStream {
  uint64_t seed;
  uint64_t counter;
  uint64_t cached_blk;
  uint32_t env_id, tag;
  Philox4 blk;
  __device__ __forceinline__ void init(uint64_t seed_, uint32_t env_id_, uint32_t tag_,
                                       uint64_t counter_) {
    seed = seed_; env_id = env_id_; tag = tag_; counter = counter_;
    cached_blk = ~0ull;
  }
  __device__ __forceinline__ uint32_t next_u32() {
    const uint64_t b = counter >> 2;
    if (b != cached_blk) {
      blk = philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), env_id, tag, (uint32_t)seed,
                          (uint32_t)(seed >> 32));
      cached_blk = b;
    }
    const uint32_t i = (uint32_t)counter & 3u;
    ++counter;
    return i == 0 ? blk.w[0] : i == 1 ? blk.w[1] : i == 2 ? blk.w[2] : blk.w[3];
  }
  template <typename real>
  __device__ __forceinline__ real next_unit() {
    return (real)(next_u32() >> 8) * (real)(1.0 / 16777216.0);
  }
  __device__ __forceinline__ int next_range(int lo, int hi) {
    return lo + (int)(((uint64_t)next_u32() * (uint64_t)(uint32_t)(hi - lo)) >> 32);
  }
};

__host__ __device__ __forceinline__ int u32_to_range(uint32_t u, int lo, int hi) {
  return lo + (int)(((uint64_t)u * (uint64_t)(uint32_t)(hi - lo)) >> 32);
}

}  // namespace madrl
