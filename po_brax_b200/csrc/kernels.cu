// Fused step / reset kernels for the po-brax Ant POMDP envs (sm_100a), plus the small layout kernels.
//
// step_kernel<KIND>  = brax.System.step (substeps) + task logic + obs + brax EpisodeWrapper + brax
//                      AutoResetWrapper in ONE launch:
//   /root/reference/po_brax/envs/ant_heavenhell.py:106-158, ant_gather.py:125-213, ant_tag.py:107-181,
//   brax.envs.ant.Ant.step, /root/reference/po_brax/envs/__init__.py:58-68 (wrapper stack).
// reset_kernel<KIND> = env.reset(rng) under vmap: ant_heavenhell.py:75-103, ant_gather.py:93-123,
//   ant_tag.py:63-105, brax.envs.ant.Ant.reset (threefry bit-exact), and with only_done=1 the gym-level
//   autoreset select of /root/reference/po_brax/envs/wrappers.py:245-262.
//
// Thread mapping: 4 lanes = 1 env (lane l owns leg l and a replica of the torso); a warp = 8 envs.
#include <cuda_runtime.h>
#include <cstdlib>
#include <math_constants.h>

#include "../../include/pobrax.h"
#include "ant_physics.cuh"
#include "threefry.cuh"

namespace pobrax {

#ifndef POBRAX_ANT_WARPS_PER_SMSP
#define POBRAX_ANT_WARPS_PER_SMSP 5
#endif
#ifndef POBRAX_WALL_WARPS_PER_SMSP
#define POBRAX_WALL_WARPS_PER_SMSP 4      // resident warps per SM sub-partition the wall variants are compiled for
#endif
constexpr int kThreads = 128;             // reset kernels
constexpr int kEnvsPerBlock = kThreads / 4;
// Step kernels use one-warp CTAs: warps never wait for each other, so finer CTAs let the hardware
// scheduler balance warps of different cost (wall contacts) and shorten the tail. Measured per env family.
template <int KIND> struct StepCfg {
#ifdef POBRAX_TUNE_THREADS   // tuning builds: one CTA size for every family
  static constexpr int threads = POBRAX_TUNE_THREADS;
#else
  static constexpr int threads = 32;   // re-swept on the final kernels (Gather ran 2-warp CTAs while its epilogue was inlined: -1.5 % now)
#endif
  static constexpr int envs = threads / 4;
  static constexpr int min_blocks = (KIND == POBRAX_ANT ? POBRAX_ANT_WARPS_PER_SMSP : POBRAX_WALL_WARPS_PER_SMSP) * (128 / threads);  // 96 / 128 registers
  // The small-batch instantiation (step_kernel<KIND, true>): compiled for 2 resident warps per sub-partition, i.e. up
  // to 255 registers -- no spills, a freer schedule. Below one warp per sub-partition a step is bound by ONE warp's
  // latency through the substeps, and that drops ~15 % (HeavenHell 128 envs 14.8 -> 12.4 us, Ant 4 096 10.2 -> 8.6 us);
  // from ~16 k envs on the throughput build wins. Same PTX for both; the inline code comes out of ptxas with the same
  // arithmetic and the out-of-line contact group is one shared function, so the two agree bit for bit
  // (tests/test_gpu_parity.py::test_small_batch_instantiation_is_bit_identical guards that).
  static constexpr int min_blocks_lat = 2 * (128 / threads);
};

// ------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ float clip1(float x) { return fminf(fmaxf(x, -1.0f), 1.0f); }
// A global load issued exactly here (volatile asm): the compilers otherwise sink the step's small scalar loads
// down to their first use behind the substep loop, where their DRAM latency is exposed.
__device__ __forceinline__ float ld_now(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// jp.norm of a 2-vector without FMA contraction (keeps <= radius tests identical to the CPU oracle's)
__device__ __forceinline__ float norm2_rn(float dx, float dy) {
  return sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
// jax.random.uniform sample: max(lo, f*(hi-lo)+lo), no contraction
__device__ __forceinline__ float uniform_rn(uint32_t bits, float lo, float hi) {
  const float f = bits_to_unit(bits);
  return fmaxf(lo, __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo));
}
__device__ __forceinline__ int quad_and(int x) {
  x &= __shfl_xor_sync(kFull, x, 1);
  x &= __shfl_xor_sync(kFull, x, 2);
  return x;
}
__device__ __forceinline__ int quad_isum(int x) {
  x += __shfl_xor_sync(kFull, x, 1);
  x += __shfl_xor_sync(kFull, x, 2);
  return x;
}
__device__ __forceinline__ float quad_min(float x) {
  x = fminf(x, __shfl_xor_sync(kFull, x, 1));
  x = fminf(x, __shfl_xor_sync(kFull, x, 2));
  return x;
}

template <int KIND>
struct ObsCols {
  static constexpr int P = (KIND == POBRAX_ANT) ? 1 : 3;  // plain Ant observes torso z only
  static constexpr int rot = P, ja = P + 4, vel = P + 12, ang = P + 15, jv = P + 18, cv = P + 26;
};

// The part of the observation every env shares: [pos0, rot0, joint angles, vel0, ang0, joint vels,
// clip(contact.vel), clip(contact.ang)] staged into this env's shared-memory row. The row was zeroed at kernel
// start and its torso / Aux contact slots already hold the accumulated impulses (ContactAcc): clip in place.
template <int KIND, bool HAVE_PSI>   // HAVE_PSI: acc.psi holds the joint angles (step kernel: the last substep's)
__device__ __forceinline__ void stage_common_obs(float* row, const Rig& r, const LegK& k, const ContactAcc& acc, int leg,
                                                 const DevConst& C) {
  using O = ObsCols<KIND>;
  // Revolute.angle_vel of the hip and the ankle, the two joints packed like in the substep:
  // hip psi = atan2(c0_A.c1_T, c0_A.c0_T), ankle psi = atan2(c2_B.(axA x c2_A), c2_A.c2_B); vel = (w_p - w_c).axis_p
  float jah, jaa, jvh, jva;
  if (HAVE_PSI) {
    // only the two parent-side axes are needed: c2_T and axA = axc c0_A + axs c1_A
    const float ts = r.T.qw, tx = r.T.qx, ty = r.T.qy, tz = r.T.qz, tz2 = tz + tz, ts2 = ts + ts;
    const V3 c2T = mk(tz2 * tx + ts2 * ty, tz2 * ty - ts2 * tx, tz2 * tz + (ts * ts - (tx * tx + ty * ty + tz * tz)));
    const Cols cA = rot_cols(r.A);
    const V3 axA = k.axc * cA.c0 + k.axs * cA.c1;
    jah = lo(acc.psi); jaa = hi(acc.psi);
    jvh = dot(r.T.w - r.A.w, c2T); jva = dot(r.A.w - r.B.w, axA);
  } else {
    const Cols cT = rot_cols(r.T);
    Body2 L;
    L.qw = pk(r.A.qw, r.B.qw); L.qx = pk(r.A.qx, r.B.qx); L.qy = pk(r.A.qy, r.B.qy); L.qz = pk(r.A.qz, r.B.qz);
    const Cols2 cL = rot_cols2(L);
    const V3 cA0 = lo3(cL.c0), cA1 = lo3(cL.c1), cA2 = lo3(cL.c2), cB2 = hi3(cL.c2);
    const V3 axA = k.axc * cA0 + k.axs * cA1;
    const V3 nA = k.axs * cA0 - k.axc * cA1;
    const F2 psi = atan2_fast2(pk(dot(cA0, cT.c1), dot(cB2, nA)), pk(dot(cA0, cT.c0), dot(cA2, cB2)));
    jah = lo(psi); jaa = hi(psi);
    jvh = dot(r.T.w - r.A.w, cT.c2); jva = dot(r.A.w - r.B.w, axA);
  }
  row[O::ja + 2 * leg] = jah; row[O::ja + 2 * leg + 1] = jaa;
  row[O::jv + 2 * leg] = jvh; row[O::jv + 2 * leg + 1] = jva;
  float* cv = acc.cv;
  float* ca = acc.ca;
  if (leg == 0) {
    if (KIND == POBRAX_ANT) {
      row[0] = r.T.p.z;
    } else {
      row[0] = r.T.p.x; row[1] = r.T.p.y; row[2] = r.T.p.z;
    }
    row[O::rot] = r.T.qw; row[O::rot + 1] = r.T.qx; row[O::rot + 2] = r.T.qy; row[O::rot + 3] = r.T.qz;
    row[O::vel] = r.T.v.x; row[O::vel + 1] = r.T.v.y; row[O::vel + 2] = r.T.v.z;
    row[O::ang] = r.T.w.x; row[O::ang + 1] = r.T.w.y; row[O::ang + 2] = r.T.w.z;
#pragma unroll
    for (int c = 0; c < 3; ++c) { cv[c] = clip1(cv[c]); ca[c] = clip1(ca[c]); }
  }
  const int a = 3 * (1 + 2 * leg), b = 3 * (2 + 2 * leg);
#pragma unroll
  for (int c = 0; c < 3; ++c) { cv[a + c] = clip1(cv[a + c]); ca[a + c] = clip1(ca[a + c]); }
  cv[b] = clip1(acc.Bv.x); cv[b + 1] = clip1(acc.Bv.y); cv[b + 2] = clip1(acc.Bv.z);
  ca[b] = clip1(acc.Bw.x); ca[b + 1] = clip1(acc.Bw.y); ca[b + 2] = clip1(acc.Bw.z);
}

// One object of AntGatherEnv._get_readings (ant_gather.py:152-181): its reading bin (n_read = "no write"; index -1
// wraps to the last reading, bombs are offset by n_apples) and intensity. Out of line, one copy: four inlined copies
// of the precise divisions and the atan2 cost the Gather kernel ~4 % (instruction fetch: ncu showed 0.76 "no
// instruction" stall cycles per issue against 0.13 for Tag), and this code runs once per env step.
__device__ __noinline__ float2 gather_bin(float ox, float oy, float dist, float ori, int is_bomb, int n_apples, int n_read,
                                          float half_span, float sensor_range, float bin_res) {
  const float ang = __fsub_rn(atan2_fast(ox, oy), ori);
  const bool valid = (fabsf(ang) <= half_span) && (dist <= sensor_range);
  int bin = valid ? (int)__fdiv_rn(__fadd_rn(ang, half_span), bin_res) : -1;
  if (is_bomb && bin >= 0) bin += n_apples;
  const float inten = bin >= 0 ? __fsub_rn(1.0f, __fdiv_rn(dist, sensor_range)) : 0.0f;
  return make_float2(__int_as_float(bin < 0 ? bin + n_read : bin), inten);
}

// AntGatherEnv._get_readings for this lane's objects (4 per lane), written in object order 0..n-1
// (sequential-scatter semantics: last writer wins).
__device__ __forceinline__ void gather_readings(float* readings, const Body& T, const float (*obj)[3], const float* dist,
                                                int leg, const DevConst& C) {
  // ori = atan2 of the torso x-axis projected on xy: (q (0,1,0,0) q^-1)[1:3]
  const float s = T.qw, x = T.qx, y = T.qy, z = T.qz;
  // quat_mul(rot, (0,1,0,0)) = (-x, s, z, -y); times quat_inv(rot) = (s,-x,-y,-z)
  const float aw = -x, ax = s, ay = z, az = -y;
  const float ox = aw * (-x) + ax * s + ay * (-z) - az * (-y);
  const float oy = aw * (-y) - ax * (-z) + ay * s + az * (-x);
  const float ori = atan2_fast(oy, ox);
  const int n_obj = C.n_apples + C.n_bombs, n_read = 2 * C.n_bins;
  int bins[4];
  float inten[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // all four lanes of the env in parallel
    const int kobj = 4 * leg + i;
    bins[i] = n_read;  // "no write"
    inten[i] = 0.0f;
    if (kobj < n_obj) {
      const float2 b = gather_bin(obj[i][0], obj[i][1], dist[i], ori, kobj >= C.n_apples, C.n_apples, n_read, C.half_span,
                                  C.sensor_range, C.bin_res);
      bins[i] = __float_as_int(b.x);
      inten[i] = b.y;
    }
  }
  for (int turn = 0; turn < 4; ++turn) {  // ordered scatter: object 0 first, object n-1 last
    if (turn == leg) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (bins[i] < n_read) readings[bins[i]] = inten[i];
    }
    __syncwarp();
  }
}

// Shared epilogue pieces ------------------------------------------------------------------------------
// The uncommon parts of write_obs_rows (below), out of line (one copy per library, keeps the step kernels' instruction footprint down): `copy` = the
// column-subset / skip-mask row copy, then the rewrite of the rows in `first_mask` from the cached first_obs.
__device__ __noinline__ void write_obs_rows_general(float* __restrict__ dst, const float* __restrict__ first,
                                                    const float* stage, int D, int lo, int Do, int rows, unsigned first_mask,
                                                    unsigned skip_mask, int lane, int copy) {
  if (copy) {
    for (int es = 0; es < rows; ++es)
      if (!((skip_mask >> es) & 1u))
#pragma unroll 1
        for (int c = lane; c < Do; c += 32) dst[es * Do + c] = stage[es * D + lo + c];
  }
  if (first_mask != 0u) {
    __syncwarp();
    for (int es = 0; es < rows; ++es)
      if ((first_mask >> es) & 1u)
#pragma unroll 1
        for (int c = lane; c < Do; c += 32) dst[es * Do + c] = first[es * Do + c];
  }
}

// Coalesced write of the warp's 8 staged observation rows (contiguous in obs[N][D]; 8*D floats start at a
// 16-byte boundary because env0 is a multiple of 8). Rows in `first_mask` are then overwritten from the cached
// first_obs (rare), rows in `skip_mask` are left untouched (reset_where_done).
template <bool OUTLINE>   // OUTLINE: the uncommon parts through write_obs_rows_general (Gather's step kernel)
__device__ __forceinline__ void write_obs_rows(float* __restrict__ obs, const float* __restrict__ first_obs,
                                               const float* stage, int D, int lo, int Do, long long env0, int n_envs,
                                               unsigned first_mask, unsigned skip_mask, int lane) {
  // D: staged row width, [lo, lo + Do): emitted columns (Do == D: the whole row, contiguous float4 path)
  const int rows = (int)min((long long)8, (long long)n_envs - env0);
  if (rows <= 0) return;
  float* dst = obs + env0 * Do;
  const bool fast = skip_mask == 0u && Do == D;
  if (fast) {
    const int n4 = (rows * D) >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    float4* d4 = reinterpret_cast<float4*>(dst);
    // warp-uniform trip counts (no per-lane predicates): blocks of 4 x 32 float4 with 4 LDS.128 in flight before the
    // first STG.128, then single rounds, then the partial round
    const int full = n4 >> 5;
    const float4* sp = s4 + lane;
    float4* dp = d4 + lane;
    int j = 0;
#pragma unroll 1
    for (; j + 4 <= full; j += 4, sp += 128, dp += 128) {
      const float4 a = sp[0], b = sp[32], c = sp[64], d = sp[96];
      dp[0] = a; dp[32] = b; dp[64] = c; dp[96] = d;
    }
#pragma unroll 1
    for (; j < full; ++j, sp += 32, dp += 32) dp[0] = sp[0];
    if (lane < (n4 & 31)) dp[0] = sp[0];
#pragma unroll 1
    for (int i = 4 * n4 + lane; i < rows * D; i += 32) dst[i] = stage[i];
  }
  if (OUTLINE) {
    if (!fast || first_mask != 0u)
      write_obs_rows_general(dst, first_obs ? first_obs + env0 * Do : nullptr, stage, D, lo, Do, rows, first_mask, skip_mask,
                             lane, fast ? 0 : 1);
    return;
  }
  if (!fast) {
    for (int es = 0; es < rows; ++es)
      if (!((skip_mask >> es) & 1u))
#pragma unroll 1
        for (int c = lane; c < Do; c += 32) dst[es * Do + c] = stage[es * D + lo + c];
  }
  if (first_mask != 0u) {
    __syncwarp();
    for (int es = 0; es < rows; ++es)
      if ((first_mask >> es) & 1u)
#pragma unroll 1
        for (int c = lane; c < Do; c += 32) dst[es * Do + c] = first_obs[(env0 + es) * Do + c];
  }
}

// --------------------------------------------------------------------------------------------- step
// SMALL = the small-batch instantiation (StepCfg::min_blocks_lat): same text, compiled for up to 255 registers.
template <int KIND, bool SMALL>
#ifdef POBRAX_TUNE_MAXNREG   // tuning builds: cap the wall variants' registers directly
__global__ void __launch_bounds__(StepCfg<KIND>::threads) __maxnreg__(KIND == POBRAX_ANT ? 96 : POBRAX_TUNE_MAXNREG)
#else
__global__ void __launch_bounds__(StepCfg<KIND>::threads, SMALL ? StepCfg<KIND>::min_blocks_lat : StepCfg<KIND>::min_blocks)
#endif
step_kernel(const __grid_constant__ DevConst C, const PobraxState S, const float* __restrict__ action) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int leg = lane & 3, es = lane >> 2;
  const int D = C.obs_dim;
  const size_t n = (size_t)C.n_envs;
  const long long env0 = ((long long)blockIdx.x * StepCfg<KIND>::envs) + warp * 8;
  const long long env_raw = env0 + es;
  const bool valid = env_raw < (long long)n;
  const size_t e = valid ? (size_t)env_raw : n - 1;  // out-of-range lanes shadow the last env (no stores)
  float* stage = smem + (size_t)warp * 8 * D;
  float* row = stage + es * D;

  // Issue every global load of the step first (state planes, action, episode scalars, task aux), then zero the
  // staging rows while they are in flight. The memory clobber / register pins keep the compiler from sinking the
  // loads down to their first use behind the substep loop.
  const LegK k = leg_consts(C, leg);
  Rig r;
  load_rig(reinterpret_cast<const float4*>(S.qp), n, e, leg, r);
  const float2 act = reinterpret_cast<const float2*>(action)[e * 4 + leg];
  // The episode scalars are only needed behind the substep loop: pull them into L2 now, load them there (their
  // L2 latency hides behind the observation math) instead of holding registers across the loop.
  if (leg == 0) {
    prefetch_l2(S.steps + e); prefetch_l2(S.done + e);
    if (C.track_metrics) prefetch_l2(S.ep_return + e);
    if (KIND == POBRAX_ANT_HEAVENHELL) prefetch_l2(S.aux + 2 * n + e);
    if (KIND == POBRAX_ANT_TAG) { prefetch_l2(S.aux + 2 * n + e); prefetch_l2(S.aux + 3 * n + e); }
  }
  if (KIND == POBRAX_ANT_GATHER) {  // the lane's 4 objects (read behind the loop)
    const int n_pl = min(12, 3 * (C.n_apples + C.n_bombs) - 12 * leg);
    const float* pa = S.aux + (size_t)(12 * leg) * n + e;
#pragma unroll 1
    for (int i = 0; i < n_pl; ++i, pa += n) prefetch_l2(pa);
  }
  // DRAM -> L2 prefetch of the state a CTA `prefetch_ctas` further on will load (CTAs start roughly in index
  // order): its prologue then waits for an L2 hit instead of a DRAM access.
  {
    const size_t ep = e + (size_t)C.prefetch_ctas * StepCfg<KIND>::envs;
    if (ep < n) {
      const float4* q = reinterpret_cast<const float4*>(S.qp);
      prefetch_l2(q + (size_t)leg * n + ep);                           // 4 torso planes, one per lane of the quad
      const float4* ql = q + (size_t)(4 + 7 * leg) * n + ep;
#pragma unroll
      for (int i = 0; i < 7; ++i) prefetch_l2(ql + (size_t)i * n);
      prefetch_l2(action + ep * 8 + 2 * leg);
    }
  }
  asm volatile("" ::: "memory");
  {
    float4* s4 = reinterpret_cast<float4*>(stage);  // 8 rows = 32*D bytes: a whole number of float4
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int i = lane; i < 2 * D; i += 32) s4[i] = z;
  }
  __syncwarp();
  const float x_before = r.T.p.x;

  // Tag: the opponent's move choice depends only on info['rng'] (ant_tag.py:131-132), not on the physics:
  // drawn here, while the state loads are still in flight, instead of serialised behind the substep loop.
  // Large batches (the throughput instantiation) get it from tag_rng_kernel, which has also advanced info['rng'].
  constexpr bool TAG_PRE = KIND == POBRAX_ANT_TAG && !SMALL;
  int tag_choice = 0;
  Key tag_knext; tag_knext.k0 = tag_knext.k1 = 0u;
  if (TAG_PRE) {
    tag_choice = (int)C.tag_choice[e];
  } else if (KIND == POBRAX_ANT_TAG) {
    Key key; key.k0 = S.rng[2 * e]; key.k1 = S.rng[2 * e + 1];
    Key k1;   // the 4 lanes of an env hold the same key: each pair of lanes shares the two blocks of a split
    split2_pair(key, lane, tag_knext, k1);
    tag_choice = randint4_pair(k1, lane);
  }
  ContactAcc acc;
  acc.Bv = acc.Bw = mk(0.f, 0.f, 0.f);
  acc.cv = row + ObsCols<KIND>::cv;
  acc.ca = acc.cv + 3 * C.nb;
  {
    // Rotated substep loop, one code copy of each half (the kernel must fit the instruction cache):
    // iteration s runs the dynamics of substep s-1 and then the kinetic update + torso / Aux wall-mask loads of
    // substep s, so those table loads are in flight during the next iteration's joint math (the lower leg's two
    // capsule-end lookups are issued inside substep2 as soon as its direction is known).
    // Inside the loop the Aux and the lower leg are the two halves of packed float32x2 registers (Rig2).
    constexpr bool W = KIND != POBRAX_ANT;
    unsigned mT = 0u, mA = 0u;
    const int nsub = C.substeps;
    const LegK2 k2 = leg_consts2(C, k, act.x, act.y);
    Rig2 p = pack_rig(r);
#pragma unroll 1
    for (int s = 0; s <= nsub; ++s) {
      if (s > 0) substep2<W>(p, k, k2, C, leg, mT, mA, acc);
      if (s < nsub) advance2<W>(p, C, mT, mA);
    }
    r = unpack_rig(p);
  }

  float steps = ld_now(S.steps + e);
  const float done_prev = ld_now(S.done + e);
  const float ep_ret = C.track_metrics ? ld_now(S.ep_return + e) : 0.0f;
  const float aux_side = (KIND == POBRAX_ANT_HEAVENHELL) ? ld_now(S.aux + 2 * n + e) : 0.0f;
  // Tag target / Gather objects: issued here so that their (L2) latency hides behind the observation math
  float tag_tx = 0.0f, tag_ty = 0.0f;
  if (KIND == POBRAX_ANT_TAG) { tag_tx = ld_now(S.aux + 2 * n + e); tag_ty = ld_now(S.aux + 3 * n + e); }
  float obj[4][3];
  if (KIND == POBRAX_ANT_GATHER) {
    const int n_obj = C.n_apples + C.n_bombs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ko = 4 * leg + i;
#pragma unroll
      for (int c = 0; c < 3; ++c) obj[i][c] = ko < n_obj ? ld_now(S.aux + (size_t)(3 * ko + c) * n + e) : 0.0f;
    }
  }
  __syncwarp();
  stage_common_obs<KIND, true>(row, r, k, acc, leg, C);
  if (C.auto_reset == POBRAX_AUTORESET_CACHED && done_prev != 0.0f) steps = 0.0f;  // AutoResetWrapper.step head
  const int extra = ObsCols<KIND>::cv + 6 * C.nb;

  // ---- task logic (computed redundantly by the 4 lanes from the replicated torso; lane 0 stores)
  const float tz = r.T.p.z;
  float dead = tz < 0.2f ? 1.0f : 0.0f;
  dead = tz > 1.0f ? 1.0f : dead;
  float reward = 0.0f, done = 0.0f, m0 = 0.0f, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
  double ev0 = 0.0, ev1 = 0.0, ev2 = 0.0;  // rare-event counters for acc[4..6]

  if (KIND == POBRAX_ANT) {
    const float forward = __fdiv_rn(__fsub_rn(r.T.p.x, x_before), C.dt);
    const float ctrl = 0.5f * quad_sum(act.x * act.x + act.y * act.y);
    // 0.5e-3 * sum over bodies of |clip(contact.vel)|^2, read back from the staged (clipped) row
    __syncwarp();
    auto sq3 = [](const float* v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; };
    const float contact = 0.5e-3f * (sq3(acc.cv) + quad_sum(sq3(acc.cv + 3 * (1 + 2 * leg)) + sq3(acc.cv + 3 * (2 + 2 * leg))));
    reward = ((forward - ctrl) - contact) + 1.0f;
    done = dead;
    m0 = ctrl; m1 = contact; m2 = forward; m3 = 1.0f;
  } else if (KIND == POBRAX_ANT_HEAVENHELL) {
    const int side = aux_side != 0.0f ? 1 : 0;
    const float hx = C.hh_xy[side][0], hy = C.hh_xy[side][1];
    const float lx = C.hh_xy[1 - side][0], ly = C.hh_xy[1 - side][1];
    const bool in_heaven = norm2_rn(hx - r.T.p.x, hy - r.T.p.y) <= C.visible_radius;
    const bool in_hell = norm2_rn(lx - r.T.p.x, ly - r.T.p.y) <= C.visible_radius;
    const bool in_priest = norm2_rn(C.priest_xy[0] - r.T.p.x, C.priest_xy[1] - r.T.p.y) <= C.visible_radius;
    reward = dead > 0.0f ? C.dying_cost : 0.0f;
    reward = in_heaven ? 1.0f : reward;
    reward = in_hell ? -1.0f : reward;
    done = reward != 0.0f ? 1.0f : 0.0f;
    m0 = done;
    ev0 = done; ev1 = (reward == 1.0f); ev2 = (reward == -1.0f);
    if (leg == 0) row[extra] = in_priest ? (hx > 0.0f ? 1.0f : (hx < 0.0f ? -1.0f : 0.0f)) : 0.0f;
  } else if (KIND == POBRAX_ANT_TAG) {
    // _step_target (ant_tag.py:129-146): rng, rng1 = split(rng); choice = randint(rng1, (), 0, 4)
    const int choice = tag_choice;
    const Key knext = tag_knext;
    const float tx = tag_tx, ty = tag_ty;
    float vx = __fsub_rn(r.T.p.x, tx), vy = __fsub_rn(r.T.p.y, ty);
    const float nv = norm2_rn(vx, vy);
    vx = __fdiv_rn(vx, nv); vy = __fdiv_rn(vy, nv);
    float cx, cy;
    if (choice == 0) { cx = vy; cy = -vx; }
    else if (choice == 1) { cx = -vy; cy = vx; }
    else if (choice == 2) { cx = -vx; cy = -vy; }
    else { cx = 0.0f; cy = 0.0f; }
    float nx = __fadd_rn(__fmul_rn(cx, C.target_step), tx), ny = __fadd_rn(__fmul_rn(cy, C.target_step), ty);
    // (|new| > cage).any() -> keep the old position (NaN compares false, like the reference)
    if (fabsf(nx) > C.cage_xy[0] || fabsf(ny) > C.cage_xy[1]) { nx = tx; ny = ty; }
    const float dist = norm2_rn(__fsub_rn(nx, r.T.p.x), __fsub_rn(ny, r.T.p.y));
    const bool vis = dist <= C.visible_radius;
    const float hit = norm2_rn(__fsub_rn(r.T.p.x, nx), __fsub_rn(r.T.p.y, ny)) <= C.tag_radius ? 1.0f : 0.0f;
    reward = dead > 0.0f ? C.dying_cost : 0.0f;
    reward = hit > 0.0f ? 1.0f : reward;
    done = (dead > 0.0f || hit > 0.0f) ? 1.0f : 0.0f;
    m0 = hit;
    ev0 = hit;
    if (leg == 0) {
      row[extra] = vis ? nx : 0.0f;
      row[extra + 1] = vis ? ny : 0.0f;
      if (valid) {
        if (!TAG_PRE) { S.rng[2 * e] = knext.k0; S.rng[2 * e + 1] = knext.k1; }
        S.aux[2 * n + e] = nx; S.aux[3 * n + e] = ny; S.aux[4 * n + e] = 1.0f;
      }
    }
  } else {  // POBRAX_ANT_GATHER
    const int n_obj = C.n_apples + C.n_bombs;
    float dist[4];
    int caught_a = 0, caught_b = 0, all_wait = 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ko = 4 * leg + i;
      if (ko < n_obj) {
        dist[i] = norm2_rn(__fsub_rn(r.T.p.x, obj[i][0]), __fsub_rn(r.T.p.y, obj[i][1]));
      } else {
        obj[i][0] = obj[i][1] = obj[i][2] = 0.0f; dist[i] = CUDART_INF_F;
      }
    }
    __syncwarp();
    gather_readings(row + extra, r.T, obj, dist, leg, C);  // obs uses the pre-pickup object positions
    // pickup (ant_gather.py:133-141): objects within catch_range move to the waiting area. Masks first, then a ROLLED
    // store loop -- pickups are rare and the unrolled form was 250 instructions of a kernel that has to fit the
    // instruction cache.
    unsigned in_mask = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ko = 4 * leg + i;
      if (ko < n_obj) {
        const bool in = dist[i] <= C.catch_range;
        in_mask |= in ? 1u << i : 0u;
        if (in) { if (ko < C.n_apples) ++caught_a; else ++caught_b; }
        all_wait &= (in || (obj[i][0] == C.waiting[0] && obj[i][1] == C.waiting[1] && obj[i][2] == C.waiting[2])) ? 1 : 0;
      }
    }
    if (valid && in_mask != 0u) {
#pragma unroll 1
      for (int i = 0; i < 4; ++i)
        if ((in_mask >> i) & 1u) {
          float* o = S.aux + (size_t)(3 * (4 * leg + i)) * n + e;
          o[0] = C.waiting[0]; o[n] = C.waiting[1]; o[2 * n] = C.waiting[2];
        }
    }
    caught_a = quad_isum(caught_a); caught_b = quad_isum(caught_b); all_wait = quad_and(all_wait);
    reward = dead > 0.0f ? C.dying_cost : 0.0f;
    reward = (caught_a > 0 && dead == 0.0f) ? 1.0f : reward;
    reward = (caught_b > 0 && dead == 0.0f) ? -1.0f : reward;
    done = all_wait ? 1.0f : dead;
    m0 = (float)caught_a; m1 = (float)caught_b;
    ev0 = caught_a; ev1 = caught_b;
  }

  // ---- brax EpisodeWrapper (episode_length, action_repeat 1): steps += 1; truncation; done at the limit
  float trunc = 0.0f;
  steps += 1.0f;
  if (C.episode_length > 0 && steps >= (float)C.episode_length) {
    trunc = 1.0f - done;
    done = 1.0f;
  }
  const bool reset_now = (C.auto_reset == POBRAX_AUTORESET_CACHED) && done != 0.0f;

  if (valid && leg == 0) {
    S.reward[e] = reward;
    S.done[e] = done;
    S.steps[e] = steps;
    S.truncation[e] = trunc;
    if (C.metrics_dim > 0) S.metrics[e] = m0;
    if (C.metrics_dim > 1) S.metrics[n + e] = m1;
    if (C.metrics_dim > 2) S.metrics[2 * n + e] = m2;
    if (C.metrics_dim > 3) S.metrics[3 * n + e] = m3;
    if (C.track_metrics) {
      const float ret = ep_ret + reward;
      S.ep_return[e] = done != 0.0f ? 0.0f : ret;
      if (done != 0.0f) {
        atomicAdd(S.acc + 0, 1.0);
        atomicAdd(S.acc + 1, (double)ret);
        atomicAdd(S.acc + 2, (double)steps);
        if (trunc != 0.0f) atomicAdd(S.acc + 3, 1.0);
      }
      if (ev0 != 0.0) atomicAdd(S.acc + 4, ev0);
      if (ev1 != 0.0) atomicAdd(S.acc + 5, ev1);
      if (ev2 != 0.0) atomicAdd(S.acc + 6, ev2);
      if (dead != 0.0f) atomicAdd(S.acc + 7, 1.0);
    }
  }

  // ---- brax AutoResetWrapper tail: qp/obs <- first_qp/first_obs where done
  if (reset_now) {
    load_rig(reinterpret_cast<const float4*>(S.first_qp), n, e, leg, r);
    if (valid) {
      for (int a = leg; a < C.aux_dim; a += 4) S.aux[(size_t)a * n + e] = S.first_aux[(size_t)a * n + e];
    }
  }
  if (valid) store_rig(reinterpret_cast<float4*>(S.qp), n, e, leg, r);
  const unsigned first_mask = __ballot_sync(kFull, reset_now && leg == 0);
  unsigned fm8 = first_mask & 0x11111111u;   // bit 4 es -> bit es (every quad's lane 0)
  fm8 = (fm8 | (fm8 >> 3)) & 0x03030303u;
  fm8 = (fm8 | (fm8 >> 6)) & 0x000f000fu;
  fm8 = (fm8 | (fm8 >> 12)) & 0xffu;
  __syncwarp();
  write_obs_rows<KIND == POBRAX_ANT_GATHER>(S.obs, S.first_obs, stage, D, C.obs_lo, C.obs_out, env0, C.n_envs, fm8, 0u, lane);
}

// -------------------------------------------------------------------------------------------- reset
// quat_mul (Hamilton), used only by default_qp
__device__ __forceinline__ void qmul(const float* u, const float* v, float* o) {
  o[0] = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  o[1] = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  o[2] = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  o[3] = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
}
__device__ __forceinline__ V3 qrot(V3 v, const float* q) {
  Body b; b.qw = q[0]; b.qx = q[1]; b.qy = q[2]; b.qz = q[3];
  const Cols c = rot_cols(b);
  return v.x * c.c0 + v.y * c.c1 + v.z * c.c2;
}

// System.default_qp for the lane's leg, torso at the origin, before the z-lift.
__device__ __forceinline__ void default_leg(Rig& r, const LegK& k, float qh, float qa, float vh, float va,
                                            const DevConst& C) {
  r.T.p = mk(0.f, 0.f, 0.f); r.T.qw = 1.f; r.T.qx = r.T.qy = r.T.qz = 0.f;
  r.T.v = r.T.w = mk(0.f, 0.f, 0.f);
  // hip: axis e_z, parent = identity torso
  float sh, ch; sincosf(0.5f * qh, &sh, &ch);
  const float qA[4] = {ch, 0.f, 0.f, sh};
  const V3 offp_h = mk(C.s_hip_p * k.ux, C.s_hip_p * k.uy, 0.f), offc_h = mk(C.s_hip_c * k.ux, C.s_hip_c * k.uy, 0.f);
  r.A.p = offp_h - qrot(offc_h, qA);
  r.A.qw = qA[0]; r.A.qx = qA[1]; r.A.qy = qA[2]; r.A.qz = qA[3];
  r.A.v = mk(0.f, 0.f, 0.f);
  r.A.w = mk(0.f, 0.f, vh);
  // ankle: axis (cos phi, sin phi, 0) in the Aux frame
  float sa, ca; sincosf(0.5f * qa, &sa, &ca);
  const float ql[4] = {ca, k.axc * sa, k.axs * sa, 0.f};
  float qB[4];
  qmul(qA, ql, qB);
  const V3 offp_a = mk(C.s_ank_p * k.ux, C.s_ank_p * k.uy, 0.f), offc_a = mk(C.s_ank_c * k.ux, C.s_ank_c * k.uy, 0.f);
  r.B.p = r.A.p + qrot(offp_a - qrot(offc_a, ql), qA);
  r.B.qw = qB[0]; r.B.qx = qB[1]; r.B.qy = qB[2]; r.B.qz = qB[3];
  r.B.v = mk(0.f, 0.f, 0.f);
  r.B.w = qrot(mk(k.axc * va, k.axs * va, 0.f), qA);
}

// The reset of ONE group of 8 consecutive envs by one warp (group = env0 / 8).
template <int KIND>
__device__ __forceinline__ void reset_group(const DevConst& C, const PobraxState& S, const uint32_t* __restrict__ keys,
                                            const float2* __restrict__ grid_xy, int only_done, uint32_t* __restrict__ chain,
                                            float* smem, long long group) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int leg = lane & 3, es = lane >> 2;
  const int D = C.obs_dim;
  const size_t n = (size_t)C.n_envs;
  const long long env0 = group * 8;
  const long long env_raw = env0 + es;
  const bool valid = env_raw < (long long)n;
  const size_t e = valid ? (size_t)env_raw : n - 1;
  float* stage = smem + (size_t)warp * 8 * D;
  float* row = stage + es * D;
  uint32_t* sort_keys = reinterpret_cast<uint32_t*>(smem + (size_t)(kThreads / 32) * 8 * D) +
                        (size_t)(warp * 8 + es) * (KIND == POBRAX_ANT_GATHER ? C.n_grid : 0);
  const bool active = !only_done || S.done[e] != 0.0f;  // uniform over the quad
  if (only_done && !__any_sync(kFull, active)) return;   // gym autoreset: most warps have no finished env
#pragma unroll 1
  for (int i = lane; i < 8 * D; i += 32) stage[i] = 0.0f;
  __syncwarp();
  const LegK k = leg_consts(C, leg);
  Key key;
  if (chain) {
    // gym key chain on the device (VmapGymWrapper._reset, wrappers.py:160-163): env i resets from
    // split(gym_key, N + 1)[i + 1]; chain[2] tells chain_advance_kernel that keys were drawn.
    Key gym; gym.k0 = chain[0]; gym.k1 = chain[1];
    key = split_at(gym, C.n_envs + 1, (int)e + 1);
    if (active) chain[2] = 1u;
  } else {
    key.k0 = keys[2 * e]; key.k1 = keys[2 * e + 1];
  }
  constexpr int NSPLIT = (KIND == POBRAX_ANT) ? 3 : (KIND == POBRAX_ANT_GATHER ? 4 : 5);
  const Key rng0 = split_at(key, NSPLIT, 0), r1 = split_at(key, NSPLIT, 1), r2 = split_at(key, NSPLIT, 2);
  Key r3 = rng0, r4 = rng0;
  if (KIND != POBRAX_ANT) r3 = split_at(key, NSPLIT, 3);
  if (KIND == POBRAX_ANT_TAG) r4 = split_at(key, NSPLIT, 4);

  // joint samples: default_angle + U(rng1, 8, -.1, .1), U(rng2, 8, -.1, .1)
  const float qh = __fadd_rn(C.hip_default, uniform_rn(random_bits_at(r1, 8, 2 * leg), -0.1f, 0.1f));
  const float qa = __fadd_rn(C.ank_default[leg], uniform_rn(random_bits_at(r1, 8, 2 * leg + 1), -0.1f, 0.1f));
  const float vh = uniform_rn(random_bits_at(r2, 8, 2 * leg), -0.1f, 0.1f);
  const float va = uniform_rn(random_bits_at(r2, 8, 2 * leg + 1), -0.1f, 0.1f);
  Rig r;
  default_leg(r, k, qh, qa, vh, va, C);
  // z-lift: lowest collider point of the tree (torso sphere, both Aux end spheres, the foot end) to z = 0
  {
    const Cols cA = rot_cols(r.A), cB = rot_cols(r.B);
    const V3 dA = k.ux * cA.c0 + k.uy * cA.c1, dB = k.ux * cB.c0 + k.uy * cB.c1;
    float mz = (0.0f + 0.0f) - C.r_torso;
    mz = fminf(mz, (r.A.p.z + C.s_aux * dA.z) - C.r_leg);
    mz = fminf(mz, (r.A.p.z - C.s_aux * dA.z) - C.r_leg);
    mz = fminf(mz, (r.B.p.z + C.s_foot * dB.z) - C.r_leg);
    mz = quad_min(mz);
    r.T.p.z -= mz; r.A.p.z -= mz; r.B.p.z -= mz;
  }

  // ---- task placement
  float gx = 0.0f, gy = 0.0f;  // ant xy offset (added to the ant parts AND the Ground body)
  float aux2 = 0.0f, aux3 = 0.0f, aux4 = 0.0f;
  Key info_rng = rng0;
  if (KIND == POBRAX_ANT_HEAVENHELL) {
    gx = uniform_rn(random_bits_at(r3, 2, 0), C.init_lo[0], C.init_hi[0]);
    gy = uniform_rn(random_bits_at(r3, 2, 1), C.init_lo[1], C.init_hi[1]);
    // choice(rng3, hhp[:2], (2,), replace=False): one shuffle round = stable sort by bits(split(rng3)[1], 2)
    Key a, b;
    split2(r3, a, b);
    const uint32_t s0 = random_bits_at(b, 2, 0), s1 = random_bits_at(b, 2, 1);
    aux2 = (s0 <= s1) ? 0.0f : 1.0f;
  } else if (KIND == POBRAX_ANT_TAG) {
    gx = uniform_rn(random_bits_at(r3, 2, 0), -C.cage_xy[0], C.cage_xy[0]);
    gy = uniform_rn(random_bits_at(r3, 2, 1), -C.cage_xy[1], C.cage_xy[1]);
    // _random_target (ant_tag.py:90-105): rejection loop on the key chain k <- split(k)[1]
    Key kk = r4;
    float tx = uniform_rn(random_bits_at(kk, 2, 0), -C.cage_xy[0], C.cage_xy[0]);
    float ty = uniform_rn(random_bits_at(kk, 2, 1), -C.cage_xy[1], C.cage_xy[1]);
    if (active) {
      while (norm2_rn(__fsub_rn(tx, gx), __fsub_rn(ty, gy)) <= C.min_spawn) {
        Key a, b;
        split2(kk, a, b);
        kk = b;
        tx = uniform_rn(random_bits_at(kk, 2, 0), -C.cage_xy[0], C.cage_xy[0]);
        ty = uniform_rn(random_bits_at(kk, 2, 1), -C.cage_xy[1], C.cage_xy[1]);
      }
    }
    aux2 = tx; aux3 = ty; aux4 = 0.5f;
  } else if (KIND == POBRAX_ANT_GATHER) {
    info_rng = key;  // ant_gather.py:106 keeps the un-split input key
  }
  r.T.p.x += gx; r.T.p.y += gy;
  r.A.p.x += gx; r.A.p.y += gy;
  r.B.p.x += gx; r.B.p.y += gy;

  float obj[4][3], dist[4];
  if (KIND == POBRAX_ANT_GATHER) {
    // choice(rng3, grid, (n_objects,), replace=False) = grid[stable_argsort(bits(split(rng3)[1], n_grid))[:n_objects]]
    const int n_obj = C.n_apples + C.n_bombs, ng = C.n_grid;
    Key a, b;
    split2(r3, a, b);
    for (int i = leg; i < ng; i += 4) sort_keys[i] = random_bits_at(b, ng, i);
    __syncwarp();
    // selection of the n_obj smallest (key, index) pairs in order; each lane scans a quarter
    uint32_t pk = 0; int pi = -1;
    for (int j = 0; j < n_obj; ++j) {
      uint32_t bk = 0xffffffffu; int bi = 0x7fffffff;
      for (int i = leg; i < ng; i += 4) {
        const uint32_t kv = sort_keys[i];
        const bool after_prev = (pi < 0) || kv > pk || (kv == pk && i > pi);
        const bool better = kv < bk || (kv == bk && i < bi);
        if (after_prev && better) { bk = kv; bi = i; }
      }
#pragma unroll
      for (int m = 1; m <= 2; m <<= 1) {
        const uint32_t ok = __shfl_xor_sync(kFull, bk, m);
        const int oi = __shfl_xor_sync(kFull, bi, m);
        if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
      }
      pk = bk; pi = bi;
      if ((j >> 2) == leg) {
        const float2 g = grid_xy[bi];
        const int i = j & 3;
        obj[i][0] = g.x; obj[i][1] = g.y; obj[i][2] = j < C.n_apples ? 1.0f : 0.0f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (4 * leg + i < n_obj) {
        dist[i] = norm2_rn(__fsub_rn(r.T.p.x, obj[i][0]), __fsub_rn(r.T.p.y, obj[i][1]));
      } else {
        obj[i][0] = obj[i][1] = obj[i][2] = 0.0f; dist[i] = CUDART_INF_F;
      }
    }
  }

  // ---- info = sys.info(qp): one collider evaluation, no integration (the velocity update is discarded)
  ContactAcc ct;
  ct.Bv = ct.Bw = mk(0.f, 0.f, 0.f);
  ct.cv = row + ObsCols<KIND>::cv;
  ct.ca = ct.cv + 3 * C.nb;
  {
    const Cols cA = rot_cols(r.A), cB = rot_cols(r.B);
    const V3 dA = k.ux * cA.c0 + k.uy * cA.c1, dB = k.ux * cB.c0 + k.uy * cB.c1;
    Rig2 tmp = pack_rig(r);
    unsigned mT = 0u, mA = 0u, mB = 0u;
    if (KIND != POBRAX_ANT && C.n_walls > 0) {
      mT = wall_mask_at(C, 0, r.T.p.x, r.T.p.y);
      mA = wall_mask_at(C, 1, r.A.p.x, r.A.p.y);
      mB = lower_leg_wall_mask<true>(tmp.L, dB, C);
    }
    contacts2<KIND != POBRAX_ANT>(tmp, C, dA, dB, mT, mA, mB, leg, ct);
  }
  __syncwarp();
  stage_common_obs<KIND, false>(row, r, k, ct, leg, C);
  const int extra = ObsCols<KIND>::cv + 6 * C.nb;
  if (KIND == POBRAX_ANT_TAG) {
    const bool vis = norm2_rn(__fsub_rn(aux2, r.T.p.x), __fsub_rn(aux3, r.T.p.y)) <= C.visible_radius;
    if (leg == 0) { row[extra] = vis ? aux2 : 0.0f; row[extra + 1] = vis ? aux3 : 0.0f; }
  } else if (KIND == POBRAX_ANT_GATHER) {
    __syncwarp();
    gather_readings(row + extra, r.T, obj, dist, leg, C);
  }  // HeavenHell: heaven_direction = 0 at reset (ant_heavenhell.py:78)

  // ---- stores
  const bool wr = valid && active;
  if (wr) {
    store_rig(reinterpret_cast<float4*>(S.qp), n, e, leg, r);
    if (!only_done && S.first_qp) store_rig(reinterpret_cast<float4*>(S.first_qp), n, e, leg, r);
    if (KIND == POBRAX_ANT_GATHER) {
      const int n_obj = C.n_apples + C.n_bombs;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ko = 4 * leg + i;
        if (ko < n_obj) {
          for (int c = 0; c < 3; ++c) {
            S.aux[(size_t)(3 * ko + c) * n + e] = obj[i][c];
            if (!only_done && S.first_aux) S.first_aux[(size_t)(3 * ko + c) * n + e] = obj[i][c];
          }
        }
      }
    }
    if (leg == 0) {
      if (KIND == POBRAX_ANT_HEAVENHELL || KIND == POBRAX_ANT_TAG) {
        const float av[5] = {gx, gy, aux2, aux3, aux4};
        for (int a = 0; a < C.aux_dim; ++a) {
          S.aux[(size_t)a * n + e] = av[a];
          if (!only_done && S.first_aux) S.first_aux[(size_t)a * n + e] = av[a];
        }
      }
      S.steps[e] = 0.0f;
      if (!only_done) {
        S.reward[e] = 0.0f; S.done[e] = 0.0f; S.truncation[e] = 0.0f;
        for (int m = 0; m < C.metrics_dim; ++m) S.metrics[(size_t)m * n + e] = 0.0f;
        if (C.has_rng) { S.rng[2 * e] = info_rng.k0; S.rng[2 * e + 1] = info_rng.k1; }
        if (S.ep_return) S.ep_return[e] = 0.0f;
      }
    }
  }
  const unsigned skip = __ballot_sync(kFull, !active && leg == 0);
  unsigned sk8 = 0;
  for (int i = 0; i < 8; ++i) sk8 |= ((skip >> (4 * i)) & 1u) << i;
  __syncwarp();
  write_obs_rows<false>(S.obs, S.obs, stage, D, C.obs_lo, C.obs_out, env0, C.n_envs, 0u, sk8, lane);
  if (!only_done && S.first_obs) write_obs_rows<false>(S.first_obs, S.obs, stage, D, C.obs_lo, C.obs_out, env0, C.n_envs, 0u, sk8, lane);
}

// only_done = 0: warp w of the grid resets group w. only_done = 1 (gym autoreset, wrappers.py:245-262): warp w SCANS the
// done flags of groups 32 w .. 32 w + 31 (one group per lane) and resets the few that hold a finished env -- with one CTA
// per 32 envs the launch spent 0.11 ms at 1 Mi envs dispatching 32 768 CTAs that had nothing to do.
template <int KIND>
__global__ void __launch_bounds__(kThreads, 2)
reset_kernel(const __grid_constant__ DevConst C, const PobraxState S, const uint32_t* __restrict__ keys,
             const float2* __restrict__ grid_xy, int only_done, uint32_t* __restrict__ chain) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const long long n_groups = ((long long)C.n_envs + 7) / 8;
  unsigned todo = (w < n_groups) ? 1u : 0u;   // only_done = 0: this warp's own group
  long long base = w;
  if (only_done) {
    const long long g = w * 32 + lane;
    bool any = false;
    if (g < n_groups) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long e = g * 8 + i;
        if (e < (long long)C.n_envs) any |= S.done[e] != 0.0f;
      }
    }
    todo = __ballot_sync(kFull, any);
    base = w * 32;
  }
#pragma unroll 1
  while (todo) {   // one code copy of the group reset for both modes
    const int j = __ffs(todo) - 1;
    todo &= todo - 1;
    reset_group<KIND>(C, S, keys, grid_xy, only_done, chain, smem, base + j);
    __syncwarp();   // the staging rows are reused by the next group
  }
}

// --------------------------------------------------------------------------- brax.QP <-> packed state
// One thread per (env, body): pos[N][nb][3], rot[N][nb][4], vel[N][nb][3], ang[N][nb][3].
__device__ __forceinline__ float packed_get(const float* qp, size_t n, size_t e, int body, int f) {
  // f: 0..12 = px py pz qw qx qy qz vx vy vz wx wy wz of ant body 0..8
  int plane, comp;
  if (body == 0) { plane = f >> 2; comp = f & 3; }
  else {
    const int leg = (body - 1) >> 1, lower = (body - 1) & 1;
    const int idx = lower ? 13 + f : f;
    plane = 4 + 7 * leg + (idx >> 2); comp = idx & 3;
  }
  return qp[((size_t)plane * n + e) * 4 + comp];
}
__device__ __forceinline__ void packed_set(float* qp, size_t n, size_t e, int body, int f, float v) {
  int plane, comp;
  if (body == 0) { plane = f >> 2; comp = f & 3; }
  else {
    const int leg = (body - 1) >> 1, lower = (body - 1) & 1;
    const int idx = lower ? 13 + f : f;
    plane = 4 + 7 * leg + (idx >> 2); comp = idx & 3;
  }
  qp[((size_t)plane * n + e) * 4 + comp] = v;
}

__global__ void unpack_qp_kernel(const __grid_constant__ DevConst C, const float* __restrict__ qp,
                                 const float* __restrict__ aux, float* __restrict__ pos, float* __restrict__ rot,
                                 float* __restrict__ vel, float* __restrict__ ang) {
  const size_t n = (size_t)C.n_envs;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C.nb) return;
  const size_t e = t / C.nb;
  const int b = (int)(t % C.nb);
  float f[13] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (b < 9) {
    for (int i = 0; i < 13; ++i) f[i] = packed_get(qp, n, e, b, i);
  } else if (b == 9) {  // Ground: xy follows the ant's spawn offset (HeavenHell / Tag)
    if (C.env_kind == POBRAX_ANT_HEAVENHELL || C.env_kind == POBRAX_ANT_TAG) { f[0] = aux[e]; f[1] = aux[n + e]; }
  } else if (C.env_kind == POBRAX_ANT_HEAVENHELL) {
    const int side = aux[2 * n + e] != 0.0f ? 1 : 0;
    if (b == 10) { f[0] = C.priest_xy[0]; f[1] = C.priest_xy[1]; f[2] = C.priest_z; }
    else if (b == 11) { f[0] = C.hh_xy[side][0]; f[1] = C.hh_xy[side][1]; f[2] = C.hh_z; }
    else if (b == 12) { f[0] = C.hh_xy[1 - side][0]; f[1] = C.hh_xy[1 - side][1]; f[2] = C.hh_z; }
    else { f[2] = C.arena_z; }
  } else if (C.env_kind == POBRAX_ANT_TAG) {
    if (b == 10) { f[0] = aux[2 * n + e]; f[1] = aux[3 * n + e]; f[2] = aux[4 * n + e]; }
    else { f[2] = C.arena_z; }
  } else if (C.env_kind == POBRAX_ANT_GATHER) {
    if (b == 10) { f[2] = C.arena_z; }
    else { const int ko = b - 11; for (int c = 0; c < 3; ++c) f[c] = aux[(size_t)(3 * ko + c) * n + e]; }
  }
  for (int c = 0; c < 3; ++c) { pos[t * 3 + c] = f[c]; vel[t * 3 + c] = f[7 + c]; ang[t * 3 + c] = f[10 + c]; }
  for (int c = 0; c < 4; ++c) rot[t * 4 + c] = f[3 + c];
}

__global__ void pack_qp_kernel(const __grid_constant__ DevConst C, const float* __restrict__ pos,
                               const float* __restrict__ rot, const float* __restrict__ vel,
                               const float* __restrict__ ang, float* __restrict__ qp, float* __restrict__ aux) {
  const size_t n = (size_t)C.n_envs;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C.nb) return;
  const size_t e = t / C.nb;
  const int b = (int)(t % C.nb);
  if (b < 9) {
    for (int c = 0; c < 3; ++c) {
      packed_set(qp, n, e, b, c, pos[t * 3 + c]);
      packed_set(qp, n, e, b, 7 + c, vel[t * 3 + c]);
      packed_set(qp, n, e, b, 10 + c, ang[t * 3 + c]);
    }
    for (int c = 0; c < 4; ++c) packed_set(qp, n, e, b, 3 + c, rot[t * 4 + c]);
    if (b == 0) {  // torso plane 3 padding
      qp[((size_t)3 * n + e) * 4 + 1] = 0.f; qp[((size_t)3 * n + e) * 4 + 2] = 0.f; qp[((size_t)3 * n + e) * 4 + 3] = 0.f;
    } else if (((b - 1) & 1) == 1) {
      const int leg = (b - 1) >> 1;
      qp[((size_t)(4 + 7 * leg + 6) * n + e) * 4 + 2] = 0.f; qp[((size_t)(4 + 7 * leg + 6) * n + e) * 4 + 3] = 0.f;
    }
  } else if (b == 9) {
    if (C.env_kind == POBRAX_ANT_HEAVENHELL || C.env_kind == POBRAX_ANT_TAG) { aux[e] = pos[t * 3]; aux[n + e] = pos[t * 3 + 1]; }
  } else if (C.env_kind == POBRAX_ANT_HEAVENHELL) {
    if (b == 11) aux[2 * n + e] = (pos[t * 3] == C.hh_xy[1][0] && pos[t * 3 + 1] == C.hh_xy[1][1] &&
                                   !(C.hh_xy[0][0] == C.hh_xy[1][0] && C.hh_xy[0][1] == C.hh_xy[1][1])) ? 1.0f : 0.0f;
  } else if (C.env_kind == POBRAX_ANT_TAG) {
    if (b == 10) { aux[2 * n + e] = pos[t * 3]; aux[3 * n + e] = pos[t * 3 + 1]; aux[4 * n + e] = pos[t * 3 + 2]; }
  } else if (C.env_kind == POBRAX_ANT_GATHER) {
    if (b >= 11) { const int ko = b - 11; for (int c = 0; c < 3; ++c) aux[(size_t)(3 * ko + c) * n + e] = pos[t * 3 + c]; }
  }
}

// jax.random.split(key, n)[first : first + count] (counter based: any slice is local work)
__global__ void split_keys_kernel(Key key, int n, int first, int count, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Key k = split_at(key, n, first + i);
  out[2 * i] = k.k0;
  out[2 * i + 1] = k.k1;
}

// vmapped jax.random.split(key, 2): keys[n][2] -> a[n][2] (split[0]), b[n][2] (split[1])
__global__ void split_pairs_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ a,
                                   uint32_t* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key k; k.k0 = keys[2 * i]; k.k1 = keys[2 * i + 1];
  Key ka, kb;
  split2(k, ka, kb);
  a[2 * i] = ka.k0; a[2 * i + 1] = ka.k1;
  b[2 * i] = kb.k0; b[2 * i + 1] = kb.k1;
}

// ------------------------------------------------------------------------------------------ launchers
static size_t obs_stage_bytes(const DevConst& C) { return (size_t)(kThreads / 32) * 8 * C.obs_dim * sizeof(float); }

template <int KIND>
static size_t step_smem_bytes(const DevConst& C) { return (size_t)(StepCfg<KIND>::threads / 32) * 8 * C.obs_dim * sizeof(float); }
template <int KIND>
static size_t reset_smem_bytes(const DevConst& C) {
  size_t smem = obs_stage_bytes(C);
  if (KIND == POBRAX_ANT_GATHER) smem += (size_t)kEnvsPerBlock * C.n_grid * sizeof(uint32_t);
  return smem;
}

// Per-handle, per-device kernel setup (called by pobrax_create with the handle's device current): the dynamic
// shared memory limits of this env family's step / reset kernels on THIS device (cudaFuncSetAttribute is per
// device) and the step kernel's L2 prefetch distance (one full wave of resident CTAs; POBRAX_PREFETCH_CTAS
// overrides, for tuning).
template <int KIND>
static cudaError_t setup_device_t(DevConst& C, size_t smem_limit, const char** what) {
  const size_t ss = step_smem_bytes<KIND>(C), rs = reset_smem_bytes<KIND>(C);
  if (ss > smem_limit) { *what = "step kernel: observation staging exceeds the device's shared memory per block"; return cudaErrorInvalidValue; }
  if (rs > smem_limit) { *what = "reset kernel: observation staging + object grid exceed the device's shared memory per block (smaller cage_xy?)"; return cudaErrorInvalidValue; }
  cudaError_t e;
  *what = "cudaFuncSetAttribute(step_kernel, MaxDynamicSharedMemorySize)";
  if ((e = cudaFuncSetAttribute(step_kernel<KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(step_kernel<KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss)) != cudaSuccess) return e;
  *what = "cudaFuncSetAttribute(reset_kernel, MaxDynamicSharedMemorySize)";
  if ((e = cudaFuncSetAttribute(reset_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs)) != cudaSuccess) return e;
  int dev = 0, sms = 148, per_sm = StepCfg<KIND>::min_blocks;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  *what = "cudaOccupancyMaxActiveBlocksPerMultiprocessor(step_kernel)";
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel<KIND, false>, StepCfg<KIND>::threads, ss)) != cudaSuccess) return e;
  const char* ov = getenv("POBRAX_PREFETCH_CTAS");
  C.prefetch_ctas = ov ? atoi(ov) : sms * per_sm;
  // batches up to 2 warps per sub-partition run the small-batch instantiation (POBRAX_SMALL_BATCH_ENVS overrides:
  // 0 = never, a large value = always; tests use it to compare the two instantiations bit for bit)
  const char* sv = getenv("POBRAX_SMALL_BATCH_ENVS");
  C.small_batch_envs = sv ? atoi(sv) : sms * 4 * 2 * 8;
  return cudaSuccess;
}

cudaError_t setup_device(DevConst& C, size_t smem_limit, const char** what) {
  switch (C.env_kind) {
    case POBRAX_ANT: return setup_device_t<POBRAX_ANT>(C, smem_limit, what);
    case POBRAX_ANT_HEAVENHELL: return setup_device_t<POBRAX_ANT_HEAVENHELL>(C, smem_limit, what);
    case POBRAX_ANT_GATHER: return setup_device_t<POBRAX_ANT_GATHER>(C, smem_limit, what);
    case POBRAX_ANT_TAG: return setup_device_t<POBRAX_ANT_TAG>(C, smem_limit, what);
  }
  *what = "unknown env_kind";
  return cudaErrorInvalidValue;
}

// AntTagEnv._step_target's random draw for the whole batch (ant_tag.py:131-132): rng, rng1 = split(rng);
// choice = randint(rng1, (), 0, 4). One thread per env, info['rng'] advanced in place, the choice left in the handle's
// scratch for the step kernel launched behind it on the same stream. Same bits as the in-kernel form (threefry.cuh).
__global__ void __launch_bounds__(256) tag_rng_kernel(uint32_t* __restrict__ rng, uint8_t* __restrict__ choice, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint2 k = reinterpret_cast<const uint2*>(rng)[e];
  Key key; key.k0 = k.x; key.k1 = k.y;
  Key knext, k1;
  split2(key, knext, k1);
  choice[e] = (uint8_t)randint4(k1);
  reinterpret_cast<uint2*>(rng)[e] = make_uint2(knext.k0, knext.k1);
}

template <int KIND>
static cudaError_t launch_step_t(const DevConst& C, const PobraxState& S, const float* action, cudaStream_t st) {
  const int blocks = (C.n_envs + StepCfg<KIND>::envs - 1) / StepCfg<KIND>::envs;
  if (C.n_envs <= C.small_batch_envs) {
    step_kernel<KIND, true><<<blocks, StepCfg<KIND>::threads, step_smem_bytes<KIND>(C), st>>>(C, S, action);
  } else {
    if (KIND == POBRAX_ANT_TAG) tag_rng_kernel<<<(C.n_envs + 255) / 256, 256, 0, st>>>(S.rng, C.tag_choice, C.n_envs);
    step_kernel<KIND, false><<<blocks, StepCfg<KIND>::threads, step_smem_bytes<KIND>(C), st>>>(C, S, action);
  }
  return cudaGetLastError();
}

// gym_key <- split(gym_key, N + 1)[0] if the reset kernel before it drew keys (some env was done)
__global__ void chain_advance_kernel(uint32_t* chain, int n_envs) {
  if (chain[2] != 0u) {
    Key gym; gym.k0 = chain[0]; gym.k1 = chain[1];
    const Key nxt = split_at(gym, n_envs + 1, 0);
    chain[0] = nxt.k0; chain[1] = nxt.k1; chain[2] = 0u;
  }
}

template <int KIND>
static cudaError_t launch_reset_t(const DevConst& C, const PobraxState& S, const uint32_t* keys, const float2* grid,
                                  int only_done, uint32_t* chain, cudaStream_t st) {
  const long long groups = ((long long)C.n_envs + 7) / 8, per_cta = (kThreads / 32) * (only_done ? 32 : 1);
  const int blocks = (int)((groups + per_cta - 1) / per_cta);
  reset_kernel<KIND><<<blocks, kThreads, reset_smem_bytes<KIND>(C), st>>>(C, S, keys, grid, only_done, chain);
  if (chain) chain_advance_kernel<<<1, 1, 0, st>>>(chain, C.n_envs);
  return cudaGetLastError();
}

cudaError_t launch_step(const DevConst& C, const PobraxState& S, const float* action, cudaStream_t st) {
  switch (C.env_kind) {
    case POBRAX_ANT: return launch_step_t<POBRAX_ANT>(C, S, action, st);
    case POBRAX_ANT_HEAVENHELL: return launch_step_t<POBRAX_ANT_HEAVENHELL>(C, S, action, st);
    case POBRAX_ANT_GATHER: return launch_step_t<POBRAX_ANT_GATHER>(C, S, action, st);
    case POBRAX_ANT_TAG: return launch_step_t<POBRAX_ANT_TAG>(C, S, action, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_reset(const DevConst& C, const PobraxState& S, const uint32_t* keys, const float2* grid,
                         int only_done, uint32_t* chain, cudaStream_t st) {
  switch (C.env_kind) {
    case POBRAX_ANT: return launch_reset_t<POBRAX_ANT>(C, S, keys, grid, only_done, chain, st);
    case POBRAX_ANT_HEAVENHELL: return launch_reset_t<POBRAX_ANT_HEAVENHELL>(C, S, keys, grid, only_done, chain, st);
    case POBRAX_ANT_GATHER: return launch_reset_t<POBRAX_ANT_GATHER>(C, S, keys, grid, only_done, chain, st);
    case POBRAX_ANT_TAG: return launch_reset_t<POBRAX_ANT_TAG>(C, S, keys, grid, only_done, chain, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_unpack(const DevConst& C, const float* qp, const float* aux, float* pos, float* rot, float* vel,
                          float* ang, cudaStream_t st) {
  const size_t total = (size_t)C.n_envs * C.nb;
  unpack_qp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(C, qp, aux, pos, rot, vel, ang);
  return cudaGetLastError();
}

cudaError_t launch_pack(const DevConst& C, const float* pos, const float* rot, const float* vel, const float* ang,
                        float* qp, float* aux, cudaStream_t st) {
  const size_t total = (size_t)C.n_envs * C.nb;
  pack_qp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(C, pos, rot, vel, ang, qp, aux);
  return cudaGetLastError();
}

cudaError_t launch_split_pairs(const uint32_t* keys, int n, uint32_t* a, uint32_t* b, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  split_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(keys, n, a, b);
  return cudaGetLastError();
}

cudaError_t launch_split_keys(const uint32_t key[2], int n, int first, int count, uint32_t* out, cudaStream_t st) {
  Key k; k.k0 = key[0]; k.k1 = key[1];
  if (count <= 0) return cudaSuccess;
  split_keys_kernel<<<(count + 255) / 256, 256, 0, st>>>(k, n, first, count, out);
  return cudaGetLastError();
}

}  // namespace pobrax

namespace pobrax {
// EvalGymWrapper.step (/root/reference/po_brax/envs/wrappers.py:202-219) as ONE launch: episode return, discounted
// return, length and running discount per env; finished episodes go into four double sums (count, return, discounted
// return, length -- what get_stats() averages) and their slots restart. The torch form was ~25 elementwise / reduction
// launches per gym step: 3x the gym step itself at scratch.py's 16 envs.
__global__ void __launch_bounds__(256) eval_update_kernel(const float* __restrict__ reward, const float* __restrict__ done,
                                                          float* __restrict__ ret, float* __restrict__ dret,
                                                          long long* __restrict__ len, float* __restrict__ disc,
                                                          double* __restrict__ sums, float discount, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0.0, sr = 0.0, sd = 0.0, sl = 0.0;
  if (e < n) {
    const float r = reward[e], g = disc[e];
    float R = __fadd_rn(ret[e], r), DR = __fadd_rn(dret[e], __fmul_rn(r, g)), G = __fmul_rn(g, discount);
    long long L = len[e] + 1;
    if (done[e] != 0.0f) {
      c = 1.0; sr = (double)R; sd = (double)DR; sl = (double)L;
      R = 0.0f; DR = 0.0f; L = 0; G = 1.0f;
    }
    ret[e] = R; dret[e] = DR; len[e] = L; disc[e] = G;
  }
  if (__any_sync(kFull, c != 0.0)) {   // finished episodes are rare: one set of atomics per warp that has any
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      c += __shfl_xor_sync(kFull, c, m); sr += __shfl_xor_sync(kFull, sr, m);
      sd += __shfl_xor_sync(kFull, sd, m); sl += __shfl_xor_sync(kFull, sl, m);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(sums + 0, c); atomicAdd(sums + 1, sr); atomicAdd(sums + 2, sd); atomicAdd(sums + 3, sl);
    }
  }
}

cudaError_t launch_eval_update(const float* reward, const float* done, float* ret, float* dret, long long* len, float* disc,
                               double* sums, float discount, int n, cudaStream_t st) {
  eval_update_kernel<<<(n + 255) / 256, 256, 0, st>>>(reward, done, ret, dret, len, disc, sums, discount, n);
  return cudaGetLastError();
}
}  // namespace pobrax

// ------------------------------------------------------------------------- FP32 FMA peak probe (bench)
// 8 independent FMA chains per thread; used by bench.py to measure the non-tensor FP32 roof on the box.
namespace pobrax {
__global__ void __launch_bounds__(256) fma_probe_kernel(float* __restrict__ out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the chains alive
}
cudaError_t launch_fma_probe(float* out, int blocks, int iters, cudaStream_t st) {
  fma_probe_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
  return cudaGetLastError();
}
}  // namespace pobrax
