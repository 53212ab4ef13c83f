// jax.random (threefry2x32, non-partitionable scheme) as device functions -- bit-exact integer outputs.
// Replaces jax.random.split / uniform / randint / choice as reached from
// /root/reference/po_brax/envs/ant_tag.py:64,92-102,131-132, ant_heavenhell.py:88-99,
// ant_gather.py:110-117 and /root/reference/po_brax/more_jp.py:71-77.
#pragma once
#include <stdint.h>

namespace pobrax {

struct Key { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// One Threefry-2x32-20 block.
__host__ __device__ __forceinline__ void threefry2x32(Key key, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks0 = key.k0, ks1 = key.k1, ks2 = key.k0 ^ key.k1 ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define POBRAX_TF_ROUND(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
  POBRAX_TF_ROUND(13) POBRAX_TF_ROUND(15) POBRAX_TF_ROUND(26) POBRAX_TF_ROUND(6)
  x0 += ks1; x1 += ks2 + 1u;
  POBRAX_TF_ROUND(17) POBRAX_TF_ROUND(29) POBRAX_TF_ROUND(16) POBRAX_TF_ROUND(24)
  x0 += ks2; x1 += ks0 + 2u;
  POBRAX_TF_ROUND(13) POBRAX_TF_ROUND(15) POBRAX_TF_ROUND(26) POBRAX_TF_ROUND(6)
  x0 += ks0; x1 += ks1 + 3u;
  POBRAX_TF_ROUND(17) POBRAX_TF_ROUND(29) POBRAX_TF_ROUND(16) POBRAX_TF_ROUND(24)
  x0 += ks1; x1 += ks2 + 4u;
  POBRAX_TF_ROUND(13) POBRAX_TF_ROUND(15) POBRAX_TF_ROUND(26) POBRAX_TF_ROUND(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef POBRAX_TF_ROUND
}

// random_bits(key, n)[i]: counters 0..n-1 (padded with one 0 when n is odd) split in halves (x0 | x1).
__host__ __device__ __forceinline__ uint32_t random_bits_at(Key key, int n, int i) {
  const int m = n + (n & 1), h = m >> 1;
  const int j = (i < h) ? i : i - h;
  uint32_t x0 = (uint32_t)j;
  uint32_t x1 = (uint32_t)((j + h) < n ? (j + h) : 0);
  threefry2x32(key, x0, x1);
  return (i < h) ? x0 : x1;
}

// split(key, num)[j] = (flat[2j], flat[2j+1]) with flat = random_bits(key, 2*num)
__host__ __device__ __forceinline__ Key split_at(Key key, int num, int j) {
  // flat[i] for i < num comes from x0 of block i, for i >= num from x1 of block i-num
  Key out;
  const int i0 = 2 * j, i1 = 2 * j + 1;
  uint32_t a0 = (uint32_t)(i0 < num ? i0 : i0 - num), a1 = a0 + (uint32_t)num;
  threefry2x32(key, a0, a1);
  out.k0 = (i0 < num) ? a0 : a1;
  uint32_t b0 = (uint32_t)(i1 < num ? i1 : i1 - num), b1 = b0 + (uint32_t)num;
  threefry2x32(key, b0, b1);
  out.k1 = (i1 < num) ? b0 : b1;
  return out;
}

// split(key, 2): both children from two blocks.
__host__ __device__ __forceinline__ void split2(Key key, Key& a, Key& b) {
  uint32_t x0 = 0, x1 = 2, y0 = 1, y1 = 3;
  threefry2x32(key, x0, x1);
  threefry2x32(key, y0, y1);
  a.k0 = x0; a.k1 = y0; b.k0 = x1; b.k1 = y1;
}

#ifdef __CUDACC__
// split(key, 2) computed by a PAIR of lanes that hold the same key (lane parity picks the block: even lanes the
// counters (0, 2), odd lanes (1, 3)); the halves are exchanged with two xor-shuffles. Same bits as split2, one
// threefry block per lane instead of two. All 32 lanes must call it.
__device__ __forceinline__ void split2_pair(Key key, int lane, Key& a, Key& b) {
  const uint32_t odd = (uint32_t)(lane & 1);
  uint32_t x0 = odd, x1 = 2u + odd;
  threefry2x32(key, x0, x1);
  const uint32_t p0 = __shfl_xor_sync(0xffffffffu, x0, 1), p1 = __shfl_xor_sync(0xffffffffu, x1, 1);
  a.k0 = odd ? p0 : x0; a.k1 = odd ? x0 : p0;
  b.k0 = odd ? p1 : x1; b.k1 = odd ? x1 : p1;
}
// randint4 with its split shared by a lane pair
__device__ __forceinline__ int randint4_pair(Key key, int lane) {
  Key a, b;
  split2_pair(key, lane, a, b);
  uint32_t x0 = 0, x1 = 0;
  threefry2x32(b, x0, x1);
  return (int)(x0 & 3u);
}
#endif

// uniform bits -> float in [0,1): bitcast((bits >> 9) | 0x3F800000) - 1
__host__ __device__ __forceinline__ float bits_to_unit(uint32_t bits) {
  union { uint32_t u; float f; } c;
  c.u = (bits >> 9) | 0x3F800000u;
  return c.f - 1.0f;
}

// jax.random.randint(key, (), 0, 4): span 4 => multiplier (2^16 % 4)^2 % 4 = 0 => lower_bits & 3,
// lower_bits = random_bits(split(key)[1], 1)[0]
__host__ __device__ __forceinline__ int randint4(Key key) {
  Key a, b;
  split2(key, a, b);
  uint32_t x0 = 0, x1 = 0;
  threefry2x32(b, x0, x1);
  return (int)(x0 & 3u);
}

// General 32-bit jax.random.randint(key, (), lo, hi)
__host__ __device__ __forceinline__ int randint(Key key, int lo, int hi) {
  Key a, b;
  split2(key, a, b);
  uint32_t h0 = 0, h1 = 0, l0 = 0, l1 = 0;
  threefry2x32(a, h0, h1);
  threefry2x32(b, l0, l1);
  const uint32_t span = hi > lo ? (uint32_t)(hi - lo) : 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  const uint32_t off = ((h0 % span) * mult + (l0 % span)) % span;
  return lo + (int)off;
}

}  // namespace pobrax
