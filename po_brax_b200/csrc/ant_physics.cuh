// Brax v1 "legacy spring" pipeline specialised to the Ant, leg-parallel: 4 lanes = 1 env.
// Lane l of an env's quad owns leg l (A = Aux l, B = lower leg l) and a bit-identical replica of the
// torso T; torso impulses are combined with two xor-shuffles per substep.
//
// What this replaces: brax.System.step / System.info as called from
//   /root/reference/po_brax/envs/ant_heavenhell.py:108,77  ant_gather.py:127,95  ant_tag.py:109,81
// (brax itself is un-vendored; the algorithm is the one restated and pinned in oracle/brax_v1.py and
// kinetic Euler -> revolute spring joints + torque actuators -> potential ->
// capsule-end/plane + capsule/box-wall one-way impulses -> collision).
#pragma once
#include <stdint.h>

#include "dev_const.h"
#include "vec.cuh"

namespace pobrax {

constexpr unsigned kFull = 0xffffffffu;

struct Cols { V3 c0, c1, c2; };  // columns of the rotation matrix R(q): rotate(v, q) = R v

// brax math.rotate written out per basis vector: 2(u.v)u + (s^2 - u.u)v + 2s(u x v)
__device__ __forceinline__ Cols rot_cols(const Body& b) {
  const float s = b.qw, x = b.qx, y = b.qy, z = b.qz;
  const float d = s * s - (x * x + y * y + z * z);
  const float x2 = x + x, y2 = y + y, z2 = z + z, s2 = s + s;
  Cols c;
  c.c0 = mk(x2 * x + d, x2 * y + s2 * z, x2 * z - s2 * y);
  c.c1 = mk(y2 * x - s2 * z, y2 * y + d, y2 * z + s2 * x);
  c.c2 = mk(z2 * x + s2 * y, z2 * y - s2 * x, z2 * z + d);
  return c;
}

// atan2 for the joint-limit logic of the substep loop: octant reduction + degree-7 minimax polynomial in t^2
// (max abs error 1.3e-7 rad incl. float32 rounding, i.e. the same 1-2 ulp class as atan2f, at ~40% of its
// instruction count; no special-case handling: the arguments are products of unit vectors, never both zero).
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = __fdividef(mn, mx);
  const float s = t * t;
  float p = -0.004054381512105465f;
  p = fmaf(p, s, 0.021862255409359932f);
  p = fmaf(p, s, -0.055911269038915634f);
  p = fmaf(p, s, 0.09642115980386734f);
  p = fmaf(p, s, -0.13908596336841583f);
  p = fmaf(p, s, 0.1994655877351761f);
  p = fmaf(p, s, -0.33329859375953674f);
  p = fmaf(p, s, 0.9999993443489075f);
  float r = p * t;
  r = ay > ax ? 1.57079632679489662f - r : r;
  r = x < 0.0f ? 3.14159265358979324f - r : r;
  return copysignf(r, y);
}

// Bare MUFU.RSQ / MUFU.RCP (no denormal-range fix-up code): the arguments here are never denormal -- quaternion
// norms, 1/m + lever^2, 1e-6 + |v|, max(|sin|, |cos|). sqrt_pos: sqrt to ~2 ulp for lengths (0 below 1e-15).
#ifndef POBRAX_HOST_EMU   // (tests/host_emu supplies correctly rounded stand-ins from its shim header)
__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
__device__ __forceinline__ float sqrt_pos(float x) { return x > 1e-30f ? x * rsqrt_ftz(x) : 0.0f; }

struct Imp { V3 dv, dw; float hit; };  // one contact's (dvel, dang) and whether it is non-zero

// OneWayCollider contact impulse on a unit-inertia body, general normal.
// rel = contact point - body pos, v = contact point velocity. Zero when pen <= 0.
__device__ __forceinline__ Imp impulse(V3 rel, V3 v, V3 n, float pen, float inv_m, float baumgarte, float friction,
                                       float elasticity) {
  Imp o;
  o.dv = o.dw = mk(0.f, 0.f, 0.f);
  o.hit = 0.0f;
  const float nv = dot(n, v);
  const V3 t1 = cross(rel, n);
  const float rden = rcp_ftz(inv_m + dot(n, cross(t1, rel)));
  const float J = (baumgarte * pen - (1.0f + elasticity) * nv) * rden;
  if (!((pen > 0.0f) && (nv < 0.0f) && (J > 0.0f))) return o;
  const V3 Jn = J * n;
  o.dv = inv_m * Jn;
  o.dw = cross(rel, Jn);
  const V3 vd = v - nv * n;
  const float nd = sqrt_pos(dot(vd, vd));
  if (nd > 0.01f) {
    const float Jd = fminf(nd * rden, friction * J);
    const V3 Jdv = (-Jd * rcp_ftz(1e-6f + nd)) * vd;
    o.dv += inv_m * Jdv;
    o.dw += cross(rel, Jdv);
  }
  o.hit = ((o.dv.x != 0.0f) || (o.dv.y != 0.0f) || (o.dv.z != 0.0f)) ? 1.0f : 0.0f;
  return o;
}

// Capsule-end vs ground plane (normal +z through the origin) for the foot, the one contact that is live on
// most substeps: `impulse` written out for n = e_z and branch-free (masked like the reference's where()s), so
// the substep loop has no divergent region for it. e = world-frame offset of the capsule end from the body.
__device__ __forceinline__ void foot_ground(const Body& b, V3 e, float r, float inv_m, const DevConst& C, V3& dv,
                                            V3& dw) {
  const float pen = r - (b.p.z + e.z);
  const float rx = e.x, ry = e.y, rz = e.z - r;
  const float vx = fmaf(b.w.y, rz, fmaf(-b.w.z, ry, b.v.x));
  const float vy = fmaf(b.w.z, rx, fmaf(-b.w.x, rz, b.v.y));
  const float nv = fmaf(b.w.x, ry, fmaf(-b.w.y, rx, b.v.z));
  const float rden = rcp_ftz(fmaf(rx, rx, fmaf(ry, ry, inv_m)));
  const float J = (C.baumgarte * pen - (1.0f + C.elasticity) * nv) * rden;
  const bool apply_n = (pen > 0.0f) && (nv < 0.0f) && (J > 0.0f);
  const float Jn = apply_n ? J : 0.0f;
  const float nd = sqrt_pos(fmaf(vx, vx, vy * vy));  // |v_d| to ~2 ulp (feeds a min() and a 0.01 threshold)
  const float cd = -fminf(nd * rden, C.friction * J) * rcp_ftz(1e-6f + nd);
  const float c = (apply_n && nd > 0.01f) ? cd : 0.0f;
  const float jx = c * vx, jy = c * vy;
  dv = mk(inv_m * jx, inv_m * jy, inv_m * Jn);
  dw = mk(ry * Jn - rz * jy, rz * jx - rx * Jn, rx * jy - ry * jx);
}

// The torso's capsule_plane candidate: a sphere of radius r centred on the body (rel = (0, 0, -r)) on the ground plane
// -- `impulse` for n = e_z with the lever's zero terms dropped (rel x n = 0: no angular term in the effective mass, the
// normal impulse carries no torque). Inline: an ant lying on its back keeps this contact live for the rest of its
// episode, in every substep.
__device__ __forceinline__ void torso_ground(const Body& b, float r, float inv_m, const DevConst& C, V3& dv, V3& dw) {
  dv = dw = mk(0.f, 0.f, 0.f);
  const float pen = r - b.p.z;
  const float vx = fmaf(-r, b.w.y, b.v.x), vy = fmaf(r, b.w.x, b.v.y), nv = b.v.z;   // v + w x rel
  const float rden = rcp_ftz(inv_m);
  const float J = (C.baumgarte * pen - (1.0f + C.elasticity) * nv) * rden;
  if (!((pen > 0.0f) && (nv < 0.0f) && (J > 0.0f))) return;
  dv.z = inv_m * J;
  const float nd = sqrt_pos(fmaf(vx, vx, vy * vy));
  if (nd > 0.01f) {
    const float cd = -fminf(nd * rden, C.friction * J) * rcp_ftz(1e-6f + nd);
    const float jx = cd * vx, jy = cd * vy;
    dv.x = inv_m * jx; dv.y = inv_m * jy;
    dw = mk(r * jy, -r * jx, 0.f);          // rel x (jx, jy, 0)
  }
}

__device__ __forceinline__ V3 clamp3(V3 p, V3 lo, V3 hi) {
  return mk(fminf(fmaxf(p.x, lo.x), hi.x), fminf(fmaxf(p.y, lo.y), hi.y), fminf(fmaxf(p.z, lo.z), hi.z));
}

// Closest point of segment a + t d (t in [0,1]) to an axis-aligned box: root of the monotone
// piecewise-linear g(t) = (p - clamp(p)).d by 16 bisections + one false-position step
// (same procedure as oracle/brax_v1.py:_closest_segment_box).
__device__ __forceinline__ float seg_box_t(V3 a, V3 d, V3 lo, V3 hi) {
  auto g = [&](float t) {
    const V3 p = a + t * d;
    return dot(p - clamp3(p, lo, hi), d);
  };
  const float g0 = g(0.0f), g1 = g(1.0f);
  if (g0 >= 0.0f) return 0.0f;
  if (g1 <= 0.0f) return 1.0f;
  float tl = 0.0f, tr = 1.0f, gl = g0, gr = g1;
#pragma unroll 1
  for (int i = 0; i < 16; ++i) {
    const float tm = 0.5f * (tl + tr);
    const float gm = g(tm);
    if (gm > 0.0f) { tr = tm; gr = gm; } else { tl = tm; gl = gm; }
  }
  const float den = gr - gl;
  return den > 0.0f ? tl - gl * (tr - tl) / den : tl;
}

// A rare contact (evaluated out of line, see rare_group): capsule (segment p + e .. p - e, radius rad) vs the axis-aligned
// Arena box [lo, hi]: one contact at the closest box point, normal (seg_pt - box_pt)/(1e-6 + d), penetration rad - d.
__device__ __forceinline__ Imp contact_general(V3 p, V3 e, V3 v, V3 w, float rad, float inv_m, V3 lo, V3 hi,
                                               float baumgarte, float friction, float elasticity) {
  const V3 a = p + e;
  const V3 d = (p - e) - a;
  const float t = seg_box_t(a, d, lo, hi);
  const V3 sp = a + t * d;
  const V3 bp = clamp3(sp, lo, hi);
  const V3 dvec = sp - bp;
  const float dist = sqrt_pos(dot(dvec, dvec));
  const float pen = rad - dist;
  if (!(pen > 0.0f)) {
    Imp o;
    o.dv = o.dw = mk(0.f, 0.f, 0.f);
    o.hit = 0.0f;
    return o;
  }
  const V3 n = rcp_ftz(1e-6f + dist) * dvec;
  const V3 rel = bp - p;
  return impulse(rel, v + cross(w, rel), n, pen, inv_m, baumgarte, friction, elasticity);
}

// Candidate walls of a body centred at (x, y): bit w of the byte set <=> wall w is within that body type's
// reach of the table cell (exact rectangle-rectangle distance in xy, so culling stays exact). One table per
// body type `kind` (0 torso sphere, 1 Aux capsule; the lower leg has its capsule-end table below), read through a
// layered 2D texture:
// cell = floor((p - origin) / cell_size) is one FMA per axis here, and the texture unit does the float -> cell
// conversion, the clamp onto the border cells (which list every wall, as does everything outside the table) and the
// addressing -- 3 instructions per lookup instead of 10 (30 lookups per lane and step).
#ifndef POBRAX_HOST_EMU   // (tests/host_emu reads the host copy of the tables from its shim header)
__device__ __forceinline__ unsigned wall_mask_at(const DevConst& C, int kind, float x, float y) {
  return tex2DLayered<unsigned char>((cudaTextureObject_t)C.wall_tex, fmaf(x, C.sdf_inv_cell, C.sdf_bx),
                                     fmaf(y, C.sdf_inv_cell, C.sdf_by), kind);
}
// Candidate walls of a lower-leg capsule END (tip or knee, a sphere of radius r_leg centred at (x, y)): bit w set <=>
// the cell is within r_leg of wall w's footprint, or within half the capsule's segment + r_leg of one of its four
// footprint vertices. That makes the two end lookups together exact for the whole capsule: against a flat face the
// distance along the segment is linear (smallest at an end), so an interior point can only be closest next to a
// vertex, and then the end nearer to it is within half a segment of that vertex. Own table at 1/32 m cells -- the
// band along a face is r_leg wide, not the capsule's reach -- read through a 2D texture like the body tables.
__device__ __forceinline__ unsigned tip_mask_at(const DevConst& C, float x, float y) {
  return tex2D<unsigned char>((cudaTextureObject_t)C.tip_tex, fmaf(x, C.tip_inv_cell, C.tip_bx),
                              fmaf(y, C.tip_inv_cell, C.tip_by));
}
#endif

// Per-lane constants of leg l.
struct LegK {
  float ux, uy;    // leg direction in the body frame (offsets and capsule ends are scalar multiples)
  float axc, axs;  // ankle joint axis (cos phi, sin phi, 0)
  float alo, ahi;  // ankle limits
};

__device__ __forceinline__ LegK leg_consts(const DevConst& C, int leg) {
  LegK k;
  k.ux = C.leg_u[leg][0]; k.uy = C.leg_u[leg][1];
  k.axc = C.ank_ax[leg][0]; k.axs = C.ank_ax[leg][1];
  k.alo = C.ank_lo[leg]; k.ahi = C.ank_hi[leg];
  return k;
}

struct Rig { Body T, A, B; };                 // one lane: torso replica + its leg
// Info.contact accumulators of the lane: the foot's (hot) live in registers; the torso's and the Aux body's
// (torso-ground and wall contacts only) accumulate straight into the env's staged observation row in shared
// memory (cv / ca = the row's contact.vel / contact.ang blocks, [nb][3] each), clipped when the row is finalised.
// psi: the (hip, ankle) joint angles of the substep just run -- the rotations do not change behind the joint math, so
// after the last substep these are the observation's joint angles (the step kernel does not compute them again).
struct ContactAcc { V3 Bv, Bw; float* cv; float* ca; F2 psi; };

__device__ __forceinline__ void row_add(float* base, int body, V3 a) {
  base[3 * body] += a.x; base[3 * body + 1] += a.y; base[3 * body + 2] += a.z;
}

__device__ __forceinline__ float quad_sum(float x) {
  x += __shfl_xor_sync(kFull, x, 1);
  x += __shfl_xor_sync(kFull, x, 2);
  return x;
}
__device__ __forceinline__ V3 quad_sum(V3 a) { return mk(quad_sum(a.x), quad_sum(a.y), quad_sum(a.z)); }
// Two V3 at once: 12 shuffles, the adds packed pairwise (same butterfly order: bit-identical in the 4 lanes)
__device__ __forceinline__ void quad_sum2(V3& a, V3& b) {
  F2 p0 = pk(a.x, a.y), p1 = pk(a.z, b.x), p2 = pk(b.y, b.z);
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
    p0 = p0 + pk(__shfl_xor_sync(kFull, lo(p0), m), __shfl_xor_sync(kFull, hi(p0), m));
    p1 = p1 + pk(__shfl_xor_sync(kFull, lo(p1), m), __shfl_xor_sync(kFull, hi(p1), m));
    p2 = p2 + pk(__shfl_xor_sync(kFull, lo(p2), m), __shfl_xor_sync(kFull, hi(p2), m));
  }
  a = mk(lo(p0), hi(p0), lo(p1)); b = mk(hi(p1), lo(p2), hi(p2));
}

// The Arena collider group of one body, out of line (one copy, register-passed arguments) so the substep loop
// stays small enough for the instruction cache: capsule (segment p + e .. p - e, radius rad) vs its candidate boxes
// (bit mask m != 0, boxes in global memory). Contacts are summed and divided by (1e-8 + #contacts with a non-zero dv).
// A box separated from the segment's bounding box by >= rad along some axis cannot touch: skipped exactly.
__device__ __noinline__ Imp rare_group(V3 p, V3 e, V3 v, V3 w, float rad, float inv_m, unsigned m,
                                       const float4* __restrict__ walls, float baumgarte, float friction,
                                       float elasticity) {
  const V3 zero = mk(0.f, 0.f, 0.f);
  Imp o;
  o.dv = o.dw = zero;
  o.hit = 0.0f;
  const V3 a = p + e, b = p - e;
  const V3 smin = mk(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
  const V3 smax = mk(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
  do {
    const int k = __ffs(m) - 1;
    m &= m - 1;
    const float4 l4 = __ldg(walls + 2 * k), h4 = __ldg(walls + 2 * k + 1);
    const V3 lo = mk(l4.x, l4.y, l4.z), hi = mk(h4.x, h4.y, h4.z);
    // exact cull: the distance between the segment and the box is at least their gap along every axis
    const float gap = fmaxf(fmaxf(fmaxf(lo.x - smax.x, smin.x - hi.x), fmaxf(lo.y - smax.y, smin.y - hi.y)),
                            fmaxf(lo.z - smax.z, smin.z - hi.z));
    if (gap < rad) {
      const Imp c = contact_general(p, e, v, w, rad, inv_m, lo, hi, baumgarte, friction, elasticity);
      o.dv += c.dv; o.dw += c.dw; o.hit += c.hit;
    }
  } while (m);
  if (o.hit > 1.0f) {  // (1e-8 + 1) == 1 in float32: only a multi-contact group divides
    const float inv = 1.0f / (1e-8f + o.hit);
    o.dv = inv * o.dv; o.dw = inv * o.dw;
  }
  return o;
}

__device__ __forceinline__ Imp wall_group(const Body& b, V3 e, float rad, float inv_m, unsigned m, const DevConst& C) {
  return rare_group(b.p, e, b.v, b.w, rad, inv_m, m, C.walls, C.baumgarte, C.friction, C.elasticity);
}

// Exact early-out of a body's Arena group with ONE candidate box k: the segment's bounding box is separated from
// the box by more than rad along some axis, so they cannot touch (the out-of-line group culls with the same gap).
__device__ __forceinline__ bool box_out_of_reach(const Body& b, V3 e, float rad, V3 lo, V3 hi) {
  const float ex = fabsf(e.x), ey = fabsf(e.y), ez = fabsf(e.z), rs = rad + 1e-5f;
  const float gap = fmaxf(fmaxf(fmaxf(lo.x - (b.p.x + ex), (b.p.x - ex) - hi.x), fmaxf(lo.y - (b.p.y + ey), (b.p.y - ey) - hi.y)),
                          fmaxf(lo.z - (b.p.z + ez), (b.p.z - ez) - hi.z));
  return gap > rs;
}
// The same as a guard in front of the out-of-line group (torso, Aux): true = single candidate and out of reach.
// An ant that lingers near a wall keeps these bodies flagged for many steps without touching; the guard costs the
// flagged lane ~20 instructions instead of the group's call + loop + global box loads.
__device__ __forceinline__ bool wall_far_single(const Body& b, V3 e, float rad, unsigned m, const DevConst& C) {
  if (m & (m - 1u)) return false;
  const int k = __ffs(m) - 1;
  const float4 l4 = C.wall_box[k][0], h4 = C.wall_box[k][1];
  return box_out_of_reach(b, e, rad, mk(l4.x, l4.y, l4.z), mk(h4.x, h4.y, h4.z));
}

// One-way contact impulse for a HORIZONTAL normal n = (nx, ny, 0): `impulse` with the zero terms dropped.
__device__ __forceinline__ Imp impulse_planar(V3 rel, V3 v, float nx, float ny, float pen, float inv_m, float baumgarte,
                                              float friction, float elasticity) {
  Imp o;
  o.dv = o.dw = mk(0.f, 0.f, 0.f);
  o.hit = 0.0f;
  const float nv = nx * v.x + ny * v.y;
  const V3 t1 = mk(-rel.z * ny, rel.z * nx, rel.x * ny - rel.y * nx);   // rel x n
  const float rden = rcp_ftz(inv_m + dot(t1, t1));                       // n.((rel x n) x rel) = |rel x n|^2
  const float J = (baumgarte * pen - (1.0f + elasticity) * nv) * rden;
  if (!((pen > 0.0f) && (nv < 0.0f) && (J > 0.0f))) return o;
  o.dv = mk(inv_m * (J * nx), inv_m * (J * ny), 0.f);
  o.dw = J * t1;
  const V3 vd = mk(v.x - nv * nx, v.y - nv * ny, v.z);
  const float nd = sqrt_pos(dot(vd, vd));
  if (nd > 0.01f) {
    const float Jd = fminf(nd * rden, friction * J);
    const V3 Jdv = (-Jd * rcp_ftz(1e-6f + nd)) * vd;
    o.dv += inv_m * Jdv;
    o.dw += cross(rel, Jdv);
  }
  o.hit = 1.0f;
  return o;
}

// Inline fast path of the lower leg's Arena group for what wall contacts are in practice (98-100 % of them over an
// episode, tools/wall_stats.py): the capsule's TIP (its t = 0 end, F = p + e) against the flat side of ONE box. Applies
// when the two end lookups name a single candidate box, the tip's closest box point lies in the tip's own height
// (dz = 0: the normal is horizontal, side face or vertical edge) and the tip is the capsule's closest point
// (g(0) >= 0 in seg_box_t <=> d.e <= 0). Same arithmetic as contact_general + impulse with the zero terms dropped.
// Returns false when the out-of-line group has to run instead (two candidates, top / bottom of the wall involved,
// knee or interior point closest).
__device__ __forceinline__ bool tip_wall(const Body& b, V3 e, float rad, float inv_m, unsigned m, const DevConst& C,
                                         Imp& c) {
  if (m & (m - 1u)) return false;
  const int k = __ffs(m) - 1;
  const float4 l4 = C.wall_box[k][0], h4 = C.wall_box[k][1];
  const V3 F = b.p + e;
  const V3 bp = clamp3(F, mk(l4.x, l4.y, l4.z), mk(h4.x, h4.y, h4.z));
  const float dx = F.x - bp.x, dy = F.y - bp.y;
  if (F.z != bp.z || dx * e.x + dy * e.y > 0.0f) return false;
  c.dv = c.dw = mk(0.f, 0.f, 0.f);
  c.hit = 0.0f;
  const float d2 = dx * dx + dy * dy, rs = rad + 1e-6f;
  if (d2 < rs * rs) {
    const float dist = sqrt_pos(d2);
    const float pen = rad - dist;
    const float inv = rcp_ftz(1e-6f + dist);
    const V3 rel = bp - b.p;
    c = impulse_planar(rel, b.v + cross(b.w, rel), inv * dx, inv * dy, pen, inv_m, C.baumgarte, C.friction, C.elasticity);
  }
  return true;
}

// =====================================================================================================
// Packed substep: the lane's Aux (A) and lower leg (B) travel as the two halves of float32x2 registers, so every
// operation the two bodies share -- kinetic update, rotation columns, lever arms, joint anchors, spring/damper
// forces, torques, the atan2 polynomial, the impulse application -- issues once (FFMA2 / FADD2 / FMUL2) instead
// of twice. The step kernels are issue-bound, so this is where their time goes. The torso stays scalar.
// Every packed operation rounds like its scalar counterpart (fma.rn / add.rn / mul.rn per half); parity bars in
// tests/_parity.py.
struct Rig2 { Body T; Body2 L; };  // L: lo = Aux, hi = lower leg

__device__ __forceinline__ Rig2 pack_rig(const Rig& r) {
  Rig2 o;
  o.T = r.T;
  o.L.p = pk3(r.A.p, r.B.p); o.L.v = pk3(r.A.v, r.B.v); o.L.w = pk3(r.A.w, r.B.w);
  o.L.qw = pk(r.A.qw, r.B.qw); o.L.qx = pk(r.A.qx, r.B.qx); o.L.qy = pk(r.A.qy, r.B.qy); o.L.qz = pk(r.A.qz, r.B.qz);
  return o;
}
__device__ __forceinline__ Rig unpack_rig(const Rig2& r) {
  Rig o;
  o.T = r.T;
  o.A.p = lo3(r.L.p); o.A.v = lo3(r.L.v); o.A.w = lo3(r.L.w);
  o.A.qw = lo(r.L.qw); o.A.qx = lo(r.L.qx); o.A.qy = lo(r.L.qy); o.A.qz = lo(r.L.qz);
  o.B.p = hi3(r.L.p); o.B.v = hi3(r.L.v); o.B.w = hi3(r.L.w);
  o.B.qw = hi(r.L.qw); o.B.qx = hi(r.L.qx); o.B.qy = hi(r.L.qy); o.B.qz = hi(r.L.qz);
  return o;
}


// integrators.kinetic: pos += vel*h; rot += quat_mul((0, ang*0.5*h), rot); rot /= |rot| (norm error <= 2 ulp, not
// accumulating: renormalised every substep)
__device__ __forceinline__ void kinetic_t(Body& b, float h) {
  b.p.x = fmaf(b.v.x, h, b.p.x);
  b.p.y = fmaf(b.v.y, h, b.p.y);
  b.p.z = fmaf(b.v.z, h, b.p.z);
  const float hh = 0.5f * h;
  const float ax = b.w.x * hh, ay = b.w.y * hh, az = b.w.z * hh;
  const float w = b.qw - ax * b.qx - ay * b.qy - az * b.qz;
  const float x = b.qx + ax * b.qw + ay * b.qz - az * b.qy;
  const float y = b.qy - ax * b.qz + ay * b.qw + az * b.qx;
  const float z = b.qz + ax * b.qy - ay * b.qx + az * b.qw;
  const float r = rsqrt_ftz(w * w + x * x + y * y + z * z);
  b.qw = w * r; b.qx = x * r; b.qy = y * r; b.qz = z * r;
}

__device__ __forceinline__ void kinetic2(Body2& b, float h) {
  b.p = fma3(h, b.v, b.p);
  const float hh = 0.5f * h;
  const F2 ax = hh * b.w.x, ay = hh * b.w.y, az = hh * b.w.z;
  const F2 w = fnma2(az, b.qz, fnma2(ay, b.qy, fnma2(ax, b.qx, b.qw)));
  const F2 x = fnma2(az, b.qy, fma2(ay, b.qz, fma2(ax, b.qw, b.qx)));
  const F2 y = fma2(az, b.qx, fma2(ay, b.qw, fnma2(ax, b.qz, b.qy)));
  const F2 z = fma2(az, b.qw, fnma2(ay, b.qx, fma2(ax, b.qy, b.qz)));
  const F2 n2 = fma2(z, z, fma2(y, y, fma2(x, x, w * w)));
  const F2 r = pk(rsqrt_ftz(lo(n2)), rsqrt_ftz(hi(n2)));
  b.qw = w * r; b.qx = x * r; b.qy = y * r; b.qz = z * r;
}

__device__ __forceinline__ Cols2 rot_cols2(const Body2& b) {
  const F2 s = b.qw, x = b.qx, y = b.qy, z = b.qz;
  const F2 x2 = x + x, y2 = y + y, z2 = z + z, s2 = s + s;
  const F2 d = fms2(s, s, fma2(z, z, fma2(y, y, x * x)));
  const F2 sx = s2 * x, sy = s2 * y, sz = s2 * z;
  Cols2 c;
  c.c0 = mk2(fma2(x2, x, d), fma2(x2, y, sz), fms2(x2, z, sy));
  c.c1 = mk2(fms2(y2, x, sz), fma2(y2, y, d), fma2(y2, z, sx));
  c.c2 = mk2(fma2(z2, x, sy), fms2(z2, y, sx), fma2(z2, z, d));
  return c;
}

// atan2_fast for two (y, x) pairs: the polynomial runs packed, the octant logic per half.
__device__ __forceinline__ F2 atan2_fast2(F2 y, F2 x) {
  const float xl = lo(x), xh = hi(x), yl = lo(y), yh = hi(y);
  const float axl = fabsf(xl), ayl = fabsf(yl), axh = fabsf(xh), ayh = fabsf(yh);
  const F2 mn = pk(fminf(axl, ayl), fminf(axh, ayh));
  const F2 t = mn * pk(rcp_ftz(fmaxf(axl, ayl)), rcp_ftz(fmaxf(axh, ayh)));
  const F2 s = t * t;
  F2 p = bc(-0.004054381512105465f);
  p = fma2(p, s, bc(0.021862255409359932f));
  p = fma2(p, s, bc(-0.055911269038915634f));
  p = fma2(p, s, bc(0.09642115980386734f));
  p = fma2(p, s, bc(-0.13908596336841583f));
  p = fma2(p, s, bc(0.1994655877351761f));
  p = fma2(p, s, bc(-0.33329859375953674f));
  p = fma2(p, s, bc(0.9999993443489075f));
  const F2 r = p * t;
  float rl = lo(r), rh = hi(r);
  rl = ayl > axl ? 1.57079632679489662f - rl : rl;
  rh = ayh > axh ? 1.57079632679489662f - rh : rh;
  rl = xl < 0.0f ? 3.14159265358979324f - rl : rl;
  rh = xh < 0.0f ? 3.14159265358979324f - rh : rh;
  return pk(copysignf(rl, yl), copysignf(rh, yh));
}

// limit_and_actuator for (hip, ankle): lim_lo / lim_hi = the two joints' limits, act_h = h * strength * action.
// dang = min(hi - psi, 0) + max(lo - psi, 0) is the limit violation (exactly 0 inside the limits), and the
// actuator torque is cut off whenever it is non-zero.
__device__ __forceinline__ F2 limit_and_actuator2(F2 sin_psi, F2 cos_psi, F2 lim_lo, F2 lim_hi, F2 act_h, float h_ls,
                                                  F2& psi) {
  psi = atan2_fast2(sin_psi, cos_psi);
  const F2 a = lim_hi - psi, b = lim_lo - psi;
  const F2 dang = pk(fminf(lo(a), 0.0f), fminf(hi(a), 0.0f)) + pk(fmaxf(lo(b), 0.0f), fmaxf(hi(b), 0.0f));
  const F2 t = pk(lo(dang) == 0.0f ? lo(act_h) : 0.0f, hi(dang) == 0.0f ? hi(act_h) : 0.0f);
  return fma2(h_ls, dang, t);
}

// Σ colliders.apply(qp) for the lane's bodies -- ground (torso sphere, foot end) + Arena walls (all three) --
// evaluated on the state in `r`; then integrators.collision (vel += dv, ang += dw) and the Info.contact sums.
// Every contact is evaluated on the same (pre-collision) state: a body's impulses are applied only after all
// of that body's contacts have been evaluated. A body has at most one ground candidate, so the ground
// group's "divide by the active count" is a no-op; the wall group divides per body.
template <bool WALLS>
__device__ __forceinline__ void contacts2(Rig2& r, const DevConst& C, V3 dA, V3 dB, unsigned mT, unsigned mA,
                                          unsigned mB, int leg, ContactAcc& acc) {   // mB: tip | knee end lookups
  Body A, B;  // scalar views of the two halves (only v / w are written back)
  A.p = lo3(r.L.p); A.v = lo3(r.L.v); A.w = lo3(r.L.w);
  B.p = hi3(r.L.p); B.v = hi3(r.L.v); B.w = hi3(r.L.w);
  V3 gv, gw;
  foot_ground(B, C.s_foot * dB, C.r_leg, C.inv_m_leg, C, gv, gw);
  const bool hitT = C.r_torso - r.T.p.z > 0.0f;
#ifdef POBRAX_TUNE_NO_RARE   // timing experiment only (WRONG physics): what the substep costs without the rare region,
  if (hitT || (WALLS && (mT | mA | mB) != 0u)) acc.Bw.z += 1e-30f;   // table lookups kept alive
  if (false) {
#else
  if (__builtin_expect(hitT || (WALLS && (mT | mA | mB) != 0u), 0)) {  // the one divergent region of the substep (rare)
#endif
    const V3 zero = mk(0.f, 0.f, 0.f);
    if (hitT || mT != 0u) {
      Imp t;
      t.dv = t.dw = zero;
      if (hitT) torso_ground(r.T, C.r_torso, C.inv_m_torso, C, t.dv, t.dw);
      if (WALLS && mT != 0u && !wall_far_single(r.T, zero, C.r_torso, mT, C)) {
        // torso and Aux in wall contact are rare (the lower legs reach furthest): no inline narrow phase, one code
        // copy out of line behind the inline out-of-reach guard
        const Imp c = wall_group(r.T, zero, C.r_torso, C.inv_m_torso, mT, C);
        t.dv += c.dv; t.dw += c.dw;
      }
      r.T.v += t.dv; r.T.w += t.dw;
      if (leg == 0) { row_add(acc.cv, 0, t.dv); row_add(acc.ca, 0, t.dw); }
    }
    if (WALLS && mA != 0u && !wall_far_single(A, C.s_aux * dA, C.r_leg, mA, C)) {
      const Imp c = wall_group(A, C.s_aux * dA, C.r_leg, C.inv_m_leg, mA, C);
      A.v += c.dv; A.w += c.dw;
      row_add(acc.cv, 1 + 2 * leg, c.dv); row_add(acc.ca, 1 + 2 * leg, c.dw);
    }
    if (WALLS && mB != 0u) {
      Imp c;
      if (!tip_wall(B, C.s_foot * dB, C.r_leg, C.inv_m_leg, mB, C, c))
        c = wall_group(B, C.s_foot * dB, C.r_leg, C.inv_m_leg, mB, C);
      B.v += c.dv; B.w += c.dw;
      acc.Bv += c.dv; acc.Bw += c.dw;
    }
  }
  B.v += gv; B.w += gw;
  acc.Bv += gv; acc.Bw += gw;
  r.L.v = pk3(A.v, B.v); r.L.w = pk3(A.w, B.w);
}

template <bool WALLS>
__device__ __forceinline__ void advance2(Rig2& r, const DevConst& C, unsigned& mT, unsigned& mA) {
  kinetic_t(r.T, C.h);
  kinetic2(r.L, C.h);
  mT = mA = 0u;
  if (WALLS && C.n_walls > 0) {
    mT = wall_mask_at(C, 0, r.T.p.x, r.T.p.y);
    mA = wall_mask_at(C, 1, lo(r.L.p.x), lo(r.L.p.y));
  }
}

// Candidate walls of the lower leg from its two capsule ends (tip_mask_at): issued as soon as the leg's rotated
// direction dB is known, consumed by the contact section behind the joint math.
template <bool WALLS>
__device__ __forceinline__ unsigned lower_leg_wall_mask(const Body2& L, V3 dB, const DevConst& C) {
  if (!WALLS || C.n_walls <= 0) return 0u;
  const float px = hi(L.p.x), py = hi(L.p.y), ex = C.s_foot * dB.x, ey = C.s_foot * dB.y;
  return tip_mask_at(C, px + ex, py + ey) | tip_mask_at(C, px - ex, py - ey);
}

// Loop-invariant packed constants of the lane.
struct LegK2 {
  F2 sc;              // child-side joint offset scales (s_hip_c, s_ank_c)
  F2 lim_lo, lim_hi;  // (hip, ankle) limits
  F2 act_h;           // h * strength * (hip action, ankle action)
};
__device__ __forceinline__ LegK2 leg_consts2(const DevConst& C, const LegK& k, float act_hip, float act_ank) {
  LegK2 q;
  q.sc = pk(C.s_hip_c, C.s_ank_c);
  q.lim_lo = pk(C.hip_lo, k.alo); q.lim_hi = pk(C.hip_hi, k.ahi);
  q.act_h = pk(act_hip * C.h_act, act_ank * C.h_act);
  return q;
}

// One physics substep for the lane's three bodies after `advance2` (integrators.kinetic) has run.
// All impulses carry the factor h (C.h_k = h*stiffness, ...), so `potential` is a plain add.
template <bool WALLS>
__device__ __forceinline__ void substep2(Rig2& r, const LegK& k, const LegK2& k2, const DevConst& C, int leg,
                                         unsigned mT, unsigned mA, ContactAcc& acc) {
  const Cols cT = rot_cols(r.T);
  const Cols2 cL = rot_cols2(r.L);
  // R u: every lever arm of the leg is a scalar times dT / dA / dB
  const V3 dT = k.ux * cT.c0 + k.uy * cT.c1;
  const V3x2 dL = fma3(k.uy, cL.c1, k.ux * cL.c0);
  const V3 xT = cross(r.T.w, dT);
  const V3x2 xL = cross(r.L.w, dL);
  const V3 dA = lo3(dL), xA = lo3(xL);
  const unsigned mB = lower_leg_wall_mask<WALLS>(r.L, hi3(dL), C);
  // ---- joint anchors: child side packed (A at the hip, B at the ankle), parent side (T at the hip, A at the
  // ankle) scalar into fresh pairs. G = h*F on the child, (hip, ankle).
  const V3x2 cp = fma3(k2.sc, dL, r.L.p), cv = fma3(k2.sc, xL, r.L.v);
  const V3 pph = fma3(C.s_hip_p, dT, r.T.p), ppa = fma3(C.s_ank_p, dA, lo3(r.L.p));
  const V3 pvh = fma3(C.s_hip_p, xT, r.T.v), pva = fma3(C.s_ank_p, xA, lo3(r.L.v));
  const V3x2 ep = pk3(pph, ppa) - cp, ev = pk3(pvh, pva) - cv;
  const V3x2 G = fma3(C.h_sd, ev, C.h_k * ep);
  // ---- joint angles. hip: axis e_z, ref -e_x: psi = atan2(c0_A.c1_T, c0_A.c0_T) (the triple product
  // (c0_T x c0_A).c2_T equals c0_A.(c2_T x c0_T)). ankle: axis (cos phi, sin phi, 0), ref e_z:
  // (c2_A x c2_B).axA = c2_B.nA with nA = axA x c2_A.
  const V3 cA0 = lo3(cL.c0), cA1 = lo3(cL.c1), cA2 = lo3(cL.c2), cB2 = hi3(cL.c2);
  const V3x2 ax = fma3(k.axs, cL.c1, k.axc * cL.c0);  // (axA, axB)
  const V3 axA = lo3(ax), axB = hi3(ax);
  const V3 nA = k.axs * cA0 - k.axc * cA1;
  const F2 s = limit_and_actuator2(pk(dot(cA0, cT.c1), dot(cB2, nA)), pk(dot(cA0, cT.c0), dot(cA2, cB2)), k2.lim_lo,
                                   k2.lim_hi, k2.act_h, C.h_ls, acc.psi);
  // ---- h*torque on the parent, (hip, ankle): -ad*(w_p - w_c) - s*axis_p + k*(axis_p x axis_c)
  const V3 wA = lo3(r.L.w);
  const V3x2 t = fma3(-C.h_ad, pk3(r.T.w - wA, wA - hi3(r.L.w)),
                      fma3(neg(s), pk3(cT.c2, axA), C.h_k * pk3(cross(cT.c2, cA2), cross(axA, axB))));
  // ---- joint impulses (x h): parent (-F/m, rp x -F + tau), child (F/m, rc x F - tau); torso summed over legs
  const V3 Gh = lo3(G), Ga = hi3(G), ta = hi3(t);
  V3 sG = Gh, dwT = fma3(-C.s_hip_p, cross(dT, Gh), lo3(t));
  quad_sum2(sG, dwT);
  // ---- integrators.potential: vel = exp(vdamp h) vel + (dv + g) h ; ang = exp(adamp h) ang + dw h
  if (C.vel_damp != 1.0f) { r.T.v = C.vel_damp * r.T.v; r.L.v = C.vel_damp * r.L.v; }
  r.T.v = fma3(-C.inv_m_torso, sG, r.T.v); r.T.v.z += C.h_g;
  r.T.w = fma3(C.ang_damp, r.T.w, dwT);
  V3x2 v = fma3(C.inv_m_leg, G, r.L.v);                    // A: +Gh/m (then -Ga/m below), B: +Ga/m
  v.z = v.z + bc(C.h_g);
  V3x2 arm = k2.sc * G;                                    // A: s_hip_c*Gh (then -s_ank_p*Ga), B: s_ank_c*Ga
  const V3 vA = fma3(-C.inv_m_leg, Ga, lo3(v)), armA = fma3(-C.s_ank_p, Ga, lo3(arm));
  v = pk3(vA, hi3(v)); arm = pk3(armA, hi3(arm));
  V3x2 dw = cross(dL, arm) - t;                            // A: ... - th (then + ta), B: ... - ta
  dw = pk3(lo3(dw) + ta, hi3(dw));
  r.L.v = v;
  r.L.w = fma3(C.ang_damp, r.L.w, dw);
  // ---- colliders on the post-potential state + integrators.collision; impulses accumulate into Info.contact
  contacts2<WALLS>(r, C, dA, hi3(dL), mT, mA, mB, leg, acc);
}

// ---- packed state load / store (layout in dev_const.h) -------------------------------------------
__device__ __forceinline__ void load_rig(const float4* __restrict__ qp, size_t n, size_t e, int leg, Rig& r) {
  const float4 t0 = qp[0 * n + e], t1 = qp[1 * n + e], t2 = qp[2 * n + e], t3 = qp[3 * n + e];
  r.T.p = mk(t0.x, t0.y, t0.z); r.T.qw = t0.w; r.T.qx = t1.x; r.T.qy = t1.y; r.T.qz = t1.z;
  r.T.v = mk(t1.w, t2.x, t2.y); r.T.w = mk(t2.z, t2.w, t3.x);
  const float4* q = qp + (size_t)(4 + 7 * leg) * n + e;
  const float4 a0 = q[0], a1 = q[n], a2 = q[2 * n], a3 = q[3 * n], a4 = q[4 * n], a5 = q[5 * n], a6 = q[6 * n];
  r.A.p = mk(a0.x, a0.y, a0.z); r.A.qw = a0.w; r.A.qx = a1.x; r.A.qy = a1.y; r.A.qz = a1.z;
  r.A.v = mk(a1.w, a2.x, a2.y); r.A.w = mk(a2.z, a2.w, a3.x);
  r.B.p = mk(a3.y, a3.z, a3.w); r.B.qw = a4.x; r.B.qx = a4.y; r.B.qy = a4.z; r.B.qz = a4.w;
  r.B.v = mk(a5.x, a5.y, a5.z); r.B.w = mk(a5.w, a6.x, a6.y);
}

__device__ __forceinline__ void store_rig(float4* __restrict__ qp, size_t n, size_t e, int leg, const Rig& r) {
  if (leg == 0) {
    qp[0 * n + e] = make_float4(r.T.p.x, r.T.p.y, r.T.p.z, r.T.qw);
    qp[1 * n + e] = make_float4(r.T.qx, r.T.qy, r.T.qz, r.T.v.x);
    qp[2 * n + e] = make_float4(r.T.v.y, r.T.v.z, r.T.w.x, r.T.w.y);
    qp[3 * n + e] = make_float4(r.T.w.z, 0.f, 0.f, 0.f);
  }
  float4* q = qp + (size_t)(4 + 7 * leg) * n + e;
  q[0] = make_float4(r.A.p.x, r.A.p.y, r.A.p.z, r.A.qw);
  q[n] = make_float4(r.A.qx, r.A.qy, r.A.qz, r.A.v.x);
  q[2 * n] = make_float4(r.A.v.y, r.A.v.z, r.A.w.x, r.A.w.y);
  q[3 * n] = make_float4(r.A.w.z, r.B.p.x, r.B.p.y, r.B.p.z);
  q[4 * n] = make_float4(r.B.qw, r.B.qx, r.B.qy, r.B.qz);
  q[5 * n] = make_float4(r.B.v.x, r.B.v.y, r.B.v.z, r.B.w.x);
  q[6 * n] = make_float4(r.B.w.y, r.B.w.z, 0.f, 0.f);
}

}  // namespace pobrax
