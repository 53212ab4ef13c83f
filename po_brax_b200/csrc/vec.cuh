// Small float3 helpers shared by the kernels.
#pragma once

namespace pobrax {

struct V3 { float x, y, z; };

__host__ __device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
// s*a + b as three FMAs
__host__ __device__ __forceinline__ V3 fma3(float s, V3 a, V3 b) { return mk(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// One rigid body of brax.QP: pos, rot (w,x,y,z), vel, ang.
struct Body { V3 p; float qw, qx, qy, qz; V3 v, w; };

// ---- packed float32x2 (Blackwell FFMA2 / FADD2 / FMUL2: one issue slot for two float32 lanes) -------------
// A pair lives in an aligned 64-bit register pair (lo, hi). ptxas folds negation (pk(-lo, -hi)), scalar broadcast
// (pk(s, s) -> the .F32 operand form) and half swaps into the instruction's operand modifiers, and packing two
// freshly computed scalars is free (they are allocated as a pair). Every operation rounds exactly like its scalar
// counterpart (fma.rn / add.rn / mul.rn per half).
#if defined(POBRAX_TUNE_SCALAR_F2)   // tuning builds: the same text on scalar FFMA / FADD / FMUL
struct F2 { float l, h; };
__device__ __forceinline__ F2 pk(float lo, float hi) { F2 r; r.l = lo; r.h = hi; return r; }
__device__ __forceinline__ F2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(F2 a) { return a.l; }
__device__ __forceinline__ float hi(F2 a) { return a.h; }
__device__ __forceinline__ F2 neg(F2 a) { return pk(-a.l, -a.h); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return pk(a.l + b.l, a.h + b.h); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return pk(a.l - b.l, a.h - b.h); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return pk(a.l * b.l, a.h * b.h); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return pk(fmaf(a.l, b.l, c.l), fmaf(a.h, b.h, c.h)); }
#elif !defined(POBRAX_HOST_EMU)   // (the g++ build of tests/host_emu supplies F2 and these primitives from its own shim header)
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 pk(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ F2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(F2 a) { return __uint_as_float((unsigned)a.v); }
__device__ __forceinline__ float hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ F2 neg(F2 a) { return pk(-lo(a), -hi(a)); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#endif
__device__ __forceinline__ F2 operator*(float s, F2 a) { return bc(s) * a; }
// a*b + c, a*b - c, c - a*b
__device__ __forceinline__ F2 fma2(float a, F2 b, F2 c) { return fma2(bc(a), b, c); }
__device__ __forceinline__ F2 fms2(F2 a, F2 b, F2 c) { return fma2(a, b, neg(c)); }
__device__ __forceinline__ F2 fnma2(F2 a, F2 b, F2 c) { return fma2(neg(a), b, c); }

struct V3x2 { F2 x, y, z; };  // two 3-vectors, component-wise packed
__device__ __forceinline__ V3x2 mk2(F2 x, F2 y, F2 z) { V3x2 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3x2 pk3(V3 l, V3 h) { return mk2(pk(l.x, h.x), pk(l.y, h.y), pk(l.z, h.z)); }
__device__ __forceinline__ V3 lo3(V3x2 a) { return mk(lo(a.x), lo(a.y), lo(a.z)); }
__device__ __forceinline__ V3 hi3(V3x2 a) { return mk(hi(a.x), hi(a.y), hi(a.z)); }
__device__ __forceinline__ V3x2 operator+(V3x2 a, V3x2 b) { return mk2(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3x2 operator-(V3x2 a, V3x2 b) { return mk2(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3x2 operator*(F2 s, V3x2 a) { return mk2(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3x2 operator*(float s, V3x2 a) { return bc(s) * a; }
__device__ __forceinline__ V3x2 fma3(F2 s, V3x2 a, V3x2 b) { return mk2(fma2(s, a.x, b.x), fma2(s, a.y, b.y), fma2(s, a.z, b.z)); }
__device__ __forceinline__ V3x2 fma3(float s, V3x2 a, V3x2 b) { return fma3(bc(s), a, b); }
__device__ __forceinline__ V3x2 cross(V3x2 a, V3x2 b) {
  return mk2(fms2(a.y, b.z, a.z * b.y), fms2(a.z, b.x, a.x * b.z), fms2(a.x, b.y, a.y * b.x));
}

// Two rigid bodies, field-wise packed (lo half = first body, hi half = second body).
struct Body2 { V3x2 p; F2 qw, qx, qy, qz; V3x2 v, w; };
struct Cols2 { V3x2 c0, c1, c2; };

}  // namespace pobrax
