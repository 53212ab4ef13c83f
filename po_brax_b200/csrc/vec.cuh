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
__host__ __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__host__ __device__ __forceinline__ V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
// s*a + b as three FMAs
__host__ __device__ __forceinline__ V3 fma3(float s, V3 a, V3 b) { return mk(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// One rigid body of brax.QP: pos, rot (w,x,y,z), vel, ang.
struct Body { V3 p; float qw, qx, qy, qz; V3 v, w; };
// Columns of the rotation matrix of a body.
struct Frame { V3 c0, c1, c2; };

}  // namespace pobrax
