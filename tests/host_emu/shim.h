// Host (g++) stand-ins for the device-only primitives of po_brax_b200/csrc/{vec.cuh, ant_physics.cuh}.
// TEST INFRASTRUCTURE ONLY: included by tests/host_emu/emu.cpp ahead of the product headers (which skip their own
// PTX / MUFU / texture definitions under POBRAX_HOST_EMU). Nothing under po_brax_b200/ includes this file.
//  * float32x2: the same operation per half in plain C++ (same rounding as fma.rn / add.rn / mul.rn.f32x2)
//  * MUFU.RSQ / MUFU.RCP: correctly rounded 1/sqrt(x), 1/x
//  * tex2DLayered on the wall candidate tables: floor + clamp by hand on the host copy (C.wall_tex = its address)
#pragma once
#include <cmath>
#include <cstring>

#include "../../po_brax_b200/csrc/dev_const.h"

namespace pobrax {

struct F2 { unsigned long long v; };
inline F2 pk(float lo, float hi) { unsigned a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); F2 r; r.v = (unsigned long long)a | ((unsigned long long)b << 32); return r; }
inline F2 bc(float s) { return pk(s, s); }
inline float lo(F2 a) { unsigned u = (unsigned)a.v; float f; memcpy(&f, &u, 4); return f; }
inline float hi(F2 a) { unsigned u = (unsigned)(a.v >> 32); float f; memcpy(&f, &u, 4); return f; }
inline F2 neg(F2 a) { return pk(-lo(a), -hi(a)); }
inline F2 operator+(F2 a, F2 b) { return pk(lo(a) + lo(b), hi(a) + hi(b)); }
inline F2 operator-(F2 a, F2 b) { return pk(lo(a) - lo(b), hi(a) - hi(b)); }
inline F2 operator*(F2 a, F2 b) { return pk(lo(a) * lo(b), hi(a) * hi(b)); }
inline F2 fma2(F2 a, F2 b, F2 c) { return pk(fmaf(lo(a), lo(b), lo(c)), fmaf(hi(a), hi(b), hi(c))); }

inline float rsqrt_ftz(float x) { return 1.0f / sqrtf(x); }
inline float rcp_ftz(float x) { return 1.0f / x; }

inline unsigned wall_mask_at(const DevConst& C, int kind, float x, float y) {
  const int ix = (int)fminf(fmaxf(floorf(fmaf(x, C.sdf_inv_cell, C.sdf_bx)), 0.0f), (float)(C.sdf_nx - 1));
  const int iy = (int)fminf(fmaxf(floorf(fmaf(y, C.sdf_inv_cell, C.sdf_by)), 0.0f), (float)(C.sdf_ny - 1));
  return reinterpret_cast<const unsigned char*>(C.wall_tex)[((size_t)kind * C.sdf_ny + iy) * C.sdf_nx + ix];
}

inline unsigned tip_mask_at(const DevConst& C, float x, float y) {
  const int ix = (int)fminf(fmaxf(floorf(fmaf(x, C.tip_inv_cell, C.tip_bx)), 0.0f), (float)(C.tip_nx - 1));
  const int iy = (int)fminf(fmaxf(floorf(fmaf(y, C.tip_inv_cell, C.tip_by)), 0.0f), (float)(C.tip_ny - 1));
  return reinterpret_cast<const unsigned char*>(C.tip_tex)[(size_t)iy * C.tip_nx + ix];
}

}  // namespace pobrax
