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

// integrators.kinetic: pos += vel*h; rot += quat_mul((0, ang*0.5*h), rot); rot /= |rot|
__device__ __forceinline__ void kinetic(Body& b, float h) {
  b.p.x = fmaf(b.v.x, h, b.p.x);
  b.p.y = fmaf(b.v.y, h, b.p.y);
  b.p.z = fmaf(b.v.z, h, b.p.z);
  const float hh = 0.5f * h;
  const float ax = b.w.x * hh, ay = b.w.y * hh, az = b.w.z * hh;
  const float w = b.qw - ax * b.qx - ay * b.qy - az * b.qz;
  const float x = b.qx + ax * b.qw + ay * b.qz - az * b.qy;
  const float y = b.qy - ax * b.qz + ay * b.qw + az * b.qx;
  const float z = b.qz + ax * b.qy - ay * b.qx + az * b.qw;
  const float n2 = w * w + x * x + y * y + z * z;
  float r = rsqrtf(n2);
  r = r * fmaf(-0.5f * n2, r * r, 1.5f);  // one Newton step: ~0.5 ulp reciprocal norm
  b.qw = w * r; b.qx = x * r; b.qy = y * r; b.qz = z * r;
}

// Revolute.apply + Torque.apply for one joint. rp/rc: world-frame lever arms (rotated offsets).
// Outputs F (force on the child; -F on the parent) and tau (torque on the parent; -tau on the child,
// actuator torque included). SURVEY App. A.3 "joints" + "actuators".
__device__ __forceinline__ void joint_force(const Body& P, const Body& Cb, V3 rp, V3 rc, V3 axis_p, V3 axis_c,
                                            V3 ref_p, V3 ref_c, float lo, float hi, float act,
                                            const DevConst& C, V3& F, V3& tau) {
  const V3 dpos = (P.p - Cb.p) + (rp - rc);
  const V3 dvel = (P.v - Cb.v) + (cross(P.w, rp) - cross(Cb.w, rc));
  F = C.k_joint * dpos + C.sd_joint * dvel;
  const float psi = atan2f(dot(cross(ref_p, ref_c), axis_p), dot(ref_p, ref_c));
  const bool below = psi < lo, above = psi > hi;
  const float dang = above ? hi - psi : (below ? lo - psi : 0.0f);
  const float t = (below || above) ? 0.0f : act * C.act_strength;
  // tau = k (axis_p x axis_c) - limitStrength*axis_p*dang - angularDamping*(w_p - w_c) - t*axis_p
  const float s = fmaf(C.ls_joint, dang, t);
  tau = C.k_joint * cross(axis_p, axis_c) - s * axis_p - C.ad_joint * (P.w - Cb.w);
}

// Joint angle and velocity for the observation (Revolute.angle_vel): psi as above, vel = (w_p - w_c).axis_p
__device__ __forceinline__ void joint_angle_vel(const Body& P, const Body& Cb, V3 axis_p, V3 ref_p, V3 ref_c,
                                                float& psi, float& vel) {
  psi = atan2f(dot(cross(ref_p, ref_c), axis_p), dot(ref_p, ref_c));
  vel = dot(P.w - Cb.w, axis_p);
}

// OneWayCollider contact impulse on a unit-inertia body, general normal.
// rel = contact point - body pos, v = contact point velocity. Adds nothing when pen <= 0.
__device__ __forceinline__ bool impulse(V3 rel, V3 v, V3 n, float pen, float inv_m, const DevConst& C, V3& dv,
                                        V3& dw) {
  const float bv = C.baumgarte * pen;
  const float nv = dot(n, v);
  const V3 t1 = cross(rel, n);
  const float ang = dot(n, cross(t1, rel));
  const float denom = inv_m + ang;
  const float J = (bv - (1.0f + C.elasticity) * nv) / denom;
  const bool apply_n = (pen > 0.0f) && (nv < 0.0f) && (J > 0.0f);
  dv = mk(0.f, 0.f, 0.f);
  dw = mk(0.f, 0.f, 0.f);
  if (!apply_n) return false;
  const V3 Jn = J * n;
  dv = inv_m * Jn;
  dw = cross(rel, Jn);
  const V3 vd = v - nv * n;
  const float nd = sqrtf(dot(vd, vd));
  if (nd > 0.01f) {
    const float Jd = fminf(nd / denom, C.friction * J);
    const V3 Jdv = (-Jd / (1e-6f + nd)) * vd;
    dv += inv_m * Jdv;
    dw += cross(rel, Jdv);
  }
  return (dv.x != 0.0f) || (dv.y != 0.0f) || (dv.z != 0.0f);
}

// Capsule-end vs ground plane (normal +z through the origin), SURVEY App. A.3 "colliders".
// e = world-frame offset of the capsule end from the body position.
__device__ __forceinline__ void ground_contact(const Body& b, V3 e, float r, float inv_m, const DevConst& C,
                                               V3& dv, V3& dw) {
  const float cz = (b.p.z + e.z) - r;
  const float pen = -cz;
  if (pen > 0.0f) {
    const V3 rel = mk(e.x, e.y, e.z - r);
    const V3 v = b.v + cross(b.w, rel);
    V3 a, c;
    impulse(rel, v, mk(0.f, 0.f, 1.f), pen, inv_m, C, a, c);
    dv += a;
    dw += c;
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

// Candidate walls of a body centred at (x, y): bit w set <=> wall w is within the largest capsule reach of the
// table cell (exact rectangle-rectangle distance, so culling stays exact); every wall outside the table.
__device__ __forceinline__ unsigned wall_mask_at(const DevConst& C, float x, float y) {
  const float fx = (x - C.sdf_x0) * C.sdf_inv_cell, fy = (y - C.sdf_y0) * C.sdf_inv_cell;
  const int ix = min(max((int)fx, 0), C.sdf_nx - 1);
  const int iy = min(max((int)fy, 0), C.sdf_ny - 1);
  return __ldg(C.wall_mask + iy * C.sdf_nx + ix);
}

// Capsule (segment p + e .. p - e, radius r) vs the candidate Arena boxes in `mask`: one contact per box at
// the closest box point; per body the contacts are summed and divided by (1e-8 + #contacts with a non-zero
// dv). Exact culling: a pair further apart than r contributes exactly zero (pen <= 0).
__device__ __forceinline__ void wall_contacts(const Body& b, V3 e, float r, float reach, float inv_m, unsigned mask,
                                              const DevConst& C, V3& dv, V3& dw) {
  V3 sv = mk(0.f, 0.f, 0.f), sw = mk(0.f, 0.f, 0.f);
  float cnt = 0.0f;
  const float reach2 = reach * reach;
  while (mask) {
    const int w = __ffs(mask) - 1;
    mask &= mask - 1;
    const V3 lo = mk(C.wall_lo[w][0], C.wall_lo[w][1], C.wall_lo[w][2]);
    const V3 hi = mk(C.wall_hi[w][0], C.wall_hi[w][1], C.wall_hi[w][2]);
    const V3 cd = b.p - clamp3(b.p, lo, hi);
    if (dot(cd, cd) > reach2) continue;
    const V3 a = b.p + e;
    const V3 d = (b.p - e) - a;
    const float t = seg_box_t(a, d, lo, hi);
    const V3 sp = a + t * d;
    const V3 bp = clamp3(sp, lo, hi);
    const V3 dvec = sp - bp;
    const float dist = sqrtf(dot(dvec, dvec));
    const float pen = r - dist;
    if (pen > 0.0f) {
      const V3 n = (1.0f / (1e-6f + dist)) * dvec;
      const V3 rel = bp - b.p;
      const V3 v = b.v + cross(b.w, rel);
      V3 a1, c1;
      if (impulse(rel, v, n, pen, inv_m, C, a1, c1)) cnt += 1.0f;
      sv += a1;
      sw += c1;
    }
  }
  const float inv = 1.0f / (1e-8f + cnt);
  dv += inv * sv;
  dw += inv * sw;
}

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
struct Contact { V3 Tv, Tw, Av, Aw, Bv, Bw; };  // contact impulses (dvel, dang) of the lane's three bodies

__device__ __forceinline__ float quad_sum(float x) {
  x += __shfl_xor_sync(kFull, x, 1);
  x += __shfl_xor_sync(kFull, x, 2);
  return x;
}
__device__ __forceinline__ V3 quad_sum(V3 a) { return mk(quad_sum(a.x), quad_sum(a.y), quad_sum(a.z)); }

// Σ colliders.apply(qp) for the lane's bodies: ground (torso sphere, foot end) + Arena walls (all three).
template <bool WALLS>
__device__ __forceinline__ void contacts(const Rig& r, const LegK& k, const DevConst& C, V3 dA, V3 dB,
                                         Contact& ct) {
  ct.Tv = ct.Tw = ct.Av = ct.Aw = ct.Bv = ct.Bw = mk(0.f, 0.f, 0.f);
  ground_contact(r.T, mk(0.f, 0.f, 0.f), C.r_torso, C.inv_m_torso, C, ct.Tv, ct.Tw);
  ground_contact(r.B, C.s_foot * dB, C.r_leg, C.inv_m_leg, C, ct.Bv, ct.Bw);
  if (WALLS) {
    const unsigned mT = wall_mask_at(C, r.T.p.x, r.T.p.y), mA = wall_mask_at(C, r.A.p.x, r.A.p.y),
                   mB = wall_mask_at(C, r.B.p.x, r.B.p.y);
    if ((mT | mA | mB) != 0u) {
      // one copy of the narrow phase: loop over the lane's bodies, selecting the operands
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        const unsigned m = i == 0 ? mT : (i == 1 ? mA : mB);
        if (m == 0u) continue;
        const Body X = i == 0 ? r.T : (i == 1 ? r.A : r.B);
        const float se = i == 0 ? 0.0f : (i == 1 ? C.s_aux : C.s_foot);
        const V3 d = i == 1 ? dA : dB;
        const float rad = i == 0 ? C.r_torso : C.r_leg;
        const float reach = (i == 0 ? C.r_torso : (i == 1 ? C.seg_aux + C.r_leg : C.seg_foot + C.r_leg)) + 1e-4f;
        V3 dv = mk(0.f, 0.f, 0.f), dw = mk(0.f, 0.f, 0.f);
        wall_contacts(X, se * d, rad, reach, i == 0 ? C.inv_m_torso : C.inv_m_leg, m, C, dv, dw);
        if (i == 0) { ct.Tv += dv; ct.Tw += dw; }
        else if (i == 1) { ct.Av += dv; ct.Aw += dw; }
        else { ct.Bv += dv; ct.Bw += dw; }
      }
    }
  }
}

// One physics substep for the lane's three bodies. act_h / act_a: hip / ankle actions.
template <bool WALLS>
__device__ __forceinline__ void substep(Rig& r, const LegK& k, float act_h, float act_a, const DevConst& C,
                                        Contact& acc) {
  const float h = C.h;
  kinetic(r.T, h);
  kinetic(r.A, h);
  kinetic(r.B, h);
  const Cols cT = rot_cols(r.T), cA = rot_cols(r.A), cB = rot_cols(r.B);
  const V3 dT = k.ux * cT.c0 + k.uy * cT.c1;  // R_T u
  const V3 dA = k.ux * cA.c0 + k.uy * cA.c1;
  const V3 dB = k.ux * cB.c0 + k.uy * cB.c1;
  // hip: Torso -> Aux. axis e_z, ref -e_x (the two minus signs cancel inside atan2)
  V3 Fh, th;
  joint_force(r.T, r.A, C.s_hip_p * dT, C.s_hip_c * dA, cT.c2, cA.c2, cT.c0, cA.c0, C.hip_lo, C.hip_hi, act_h, C,
              Fh, th);
  // ankle: Aux -> lower leg. axis (cos phi, sin phi, 0), ref e_z
  const V3 axA = k.axc * cA.c0 + k.axs * cA.c1;
  const V3 axB = k.axc * cB.c0 + k.axs * cB.c1;
  V3 Fa, ta;
  joint_force(r.A, r.B, C.s_ank_p * dA, C.s_ank_c * dB, axA, axB, cA.c2, cB.c2, k.alo, k.ahi, act_a, C, Fa, ta);
  // impulses: parent gets (-F/m, rp x -F + tau), child gets (F/m, rc x F - tau)
  V3 dvT = -C.inv_m_torso * Fh;
  V3 dwT = th - cross(C.s_hip_p * dT, Fh);
  dvT = quad_sum(dvT);
  dwT = quad_sum(dwT);
  const V3 dvA = C.inv_m_leg * (Fh - Fa);
  const V3 dwA = (cross(C.s_hip_c * dA, Fh) - th) + (ta - cross(C.s_ank_p * dA, Fa));
  const V3 dvB = C.inv_m_leg * Fa;
  const V3 dwB = cross(C.s_ank_c * dB, Fa) - ta;
  // integrators.potential: vel = exp(vdamp h) vel + (dv + g) h ; ang = exp(adamp h) ang + dw h
  const V3 g = mk(0.f, 0.f, C.gravity_z);
  r.T.v = C.vel_damp * r.T.v + h * (dvT + g);
  r.A.v = C.vel_damp * r.A.v + h * (dvA + g);
  r.B.v = C.vel_damp * r.B.v + h * (dvB + g);
  r.T.w = C.ang_damp * r.T.w + h * dwT;
  r.A.w = C.ang_damp * r.A.w + h * dwA;
  r.B.w = C.ang_damp * r.B.w + h * dwB;
  // colliders on the post-potential state, then integrators.collision
  Contact ct;
  contacts<WALLS>(r, k, C, dA, dB, ct);
  r.T.v += ct.Tv; r.T.w += ct.Tw;
  r.A.v += ct.Av; r.A.w += ct.Aw;
  r.B.v += ct.Bv; r.B.w += ct.Bw;
  acc.Tv += ct.Tv; acc.Tw += ct.Tw;
  acc.Av += ct.Av; acc.Aw += ct.Aw;
  acc.Bv += ct.Bv; acc.Bw += ct.Bw;
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
