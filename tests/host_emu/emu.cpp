// Host (g++, no GPU) compilation of the DEVICE physics source po_brax_b200/csrc/ant_physics.cuh.
//
// TEST INFRASTRUCTURE ONLY (tests/test_host_emu.py, `-m "not gpu"`): the product never loads this library. It lets
// the CPU suite check the text of the substep the step kernels run -- advance2 / substep2 / contacts2, the packed
// float32x2 arithmetic, the wall candidate tables, the out-of-line contact groups -- against the oracle without a
// GPU, and gives kernel work a GPU-free first gate. The real parity tests stay the `-m gpu` ones: this build has
// g++'s evaluation order (no FMA contraction), correctly rounded 1/x and 1/sqrt instead of MUFU, and a floor()
// instead of the texture unit.
//
// How the SIMT parts are emulated:
//  * 4 lanes = 1 env. The only cross-lane operation of the substep is quad_sum2 (two xor-shuffle rounds). The quad
//    is evaluated lane after lane, repeatedly from the saved pre-substep state: a shuffle returns the value its
//    partner lane recorded for the same call in the previous pass, and passes repeat until no recorded value
//    changes (3 passes for a two-round butterfly).
//  * tex2DLayered -> the host copy of the candidate tables (C.wall_tex holds its address), __ldg -> a load,
//    the f32x2 PTX -> the same operation per half: tests/host_emu/shim.h, included ahead of the product headers
//    (which only skip their own PTX / MUFU / texture definitions under POBRAX_HOST_EMU).
//  * DevConst comes from the product's own host code: api.cu is compiled in as host C++ (its launch_* entry points
//    are stubbed: nothing here can launch a kernel).
#define POBRAX_HOST_EMU 1
#include <cuda_runtime.h>   // host API + vector types only under g++

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// every standard header api.cu / the physics header need is in by now: libstdc++ spells its own attributes
// __noinline__, so the CUDA keyword can only become a macro behind them
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

#include "../../po_brax_b200/csrc/api.cu"   // host side of the C ABI: build_dev_const, defaults, arena builders

namespace pobrax {   // the launch entry points api.cu declares (defined in kernels.cu in the product)
cudaError_t launch_step(const DevConst&, const PobraxState&, const float*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_reset(const DevConst&, const PobraxState&, const uint32_t*, const float2*, int, uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_unpack(const DevConst&, const float*, const float*, float*, float*, float*, float*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_pack(const DevConst&, const float*, const float*, const float*, const float*, float*, float*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_split_keys(const uint32_t*, int, int, int, uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_fma_probe(float*, int, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_split_pairs(const uint32_t*, int, uint32_t*, uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_eval_update(const float*, const float*, float*, float*, long long*, float*, double*, float, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t setup_device(DevConst&, size_t, const char** what) { *what = "host emulator"; return cudaErrorNotSupported; }
}  // namespace pobrax

// ---- device intrinsics the physics header uses
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __ffs(unsigned m) { return __builtin_ffs((int)m); }
template <class T> static inline T __ldg(const T* p) { return *p; }

namespace emu {
constexpr int kMaxCalls = 64;
struct Quad {
  int lane = 0, call = 0;
  float cur[kMaxCalls][4], prev[kMaxCalls][4];
};
static thread_local Quad q;
}  // namespace emu

static inline float __shfl_xor_sync(unsigned, float v, int m) {
  emu::Quad& q = emu::q;
  const int c = q.call < emu::kMaxCalls ? q.call : emu::kMaxCalls - 1;   // step_env reports the overflow
  ++q.call;
  q.cur[c][q.lane] = v;
  return q.prev[c][q.lane ^ m];
}

#include "shim.h"                                // host stand-ins for the PTX / MUFU / texture primitives
#include "../../po_brax_b200/csrc/ant_physics.cuh"

namespace {
using namespace pobrax;

struct EmuHandle {
  DevConst C;
  std::vector<uint8_t> sdf, tip;
  std::vector<float2> grid;
  std::vector<float4> walls;
};

Body body_from(const float* pos, const float* rot, const float* vel, const float* ang, int b) {
  Body o;
  o.p = mk(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
  o.qw = rot[4 * b]; o.qx = rot[4 * b + 1]; o.qy = rot[4 * b + 2]; o.qz = rot[4 * b + 3];
  o.v = mk(vel[3 * b], vel[3 * b + 1], vel[3 * b + 2]);
  o.w = mk(ang[3 * b], ang[3 * b + 1], ang[3 * b + 2]);
  return o;
}
void body_to(const Body& o, float* pos, float* rot, float* vel, float* ang, int b) {
  pos[3 * b] = o.p.x; pos[3 * b + 1] = o.p.y; pos[3 * b + 2] = o.p.z;
  rot[4 * b] = o.qw; rot[4 * b + 1] = o.qx; rot[4 * b + 2] = o.qy; rot[4 * b + 3] = o.qz;
  vel[3 * b] = o.v.x; vel[3 * b + 1] = o.v.y; vel[3 * b + 2] = o.v.z;
  ang[3 * b] = o.w.x; ang[3 * b + 1] = o.w.y; ang[3 * b + 2] = o.w.z;
}

struct LaneState { Rig2 p; V3 Bv, Bw; unsigned mT, mA; };

// One env step of one env, the way step_kernel runs it (kernels.cu, the rotated substep loop), WALLS as a template.
template <bool W>
int step_env(const DevConst& C, float* pos, float* rot, float* vel, float* ang, const float* act, float* cv, float* ca) {
  const int nb = C.nb;
  LaneState L[4];
  LegK k[4];
  LegK2 k2[4];
  for (int l = 0; l < 4; ++l) {
    Rig r;
    r.T = body_from(pos, rot, vel, ang, 0);
    r.A = body_from(pos, rot, vel, ang, 1 + 2 * l);
    r.B = body_from(pos, rot, vel, ang, 2 + 2 * l);
    k[l] = leg_consts(C, l);
    k2[l] = leg_consts2(C, k[l], act[2 * l], act[2 * l + 1]);
    L[l].p = pack_rig(r);
    L[l].Bv = L[l].Bw = mk(0.f, 0.f, 0.f);
    L[l].mT = L[l].mA = 0u;
  }
  std::vector<float> row_cv(3 * nb, 0.f), row_ca(3 * nb, 0.f);   // the staged observation row's contact blocks
  emu::Quad& q = emu::q;
  for (int s = 0; s <= C.substeps; ++s) {
    if (s > 0) {
      LaneState saved[4];
      std::memcpy(saved, L, sizeof(L));
      const std::vector<float> cv0 = row_cv, ca0 = row_ca;
      std::memset(q.prev, 0, sizeof(q.prev));
      int pass = 0;
      for (;; ++pass) {
        if (pass > 8) return 2;   // the shuffle fix point did not converge
        std::memcpy(L, saved, sizeof(L));
        row_cv = cv0; row_ca = ca0;
        std::memset(q.cur, 0, sizeof(q.cur));
        int calls = 0;
        for (int l = 0; l < 4; ++l) {
          q.lane = l; q.call = 0;
          ContactAcc acc;
          acc.Bv = L[l].Bv; acc.Bw = L[l].Bw;
          acc.cv = row_cv.data(); acc.ca = row_ca.data();
          substep2<W>(L[l].p, k[l], k2[l], C, l, L[l].mT, L[l].mA, acc);
          L[l].Bv = acc.Bv; L[l].Bw = acc.Bw;
          if (q.call > emu::kMaxCalls) return 3;
          calls = q.call;
        }
        const bool same = std::memcmp(q.cur, q.prev, sizeof(float) * 4 * calls) == 0;
        std::memcpy(q.prev, q.cur, sizeof(q.cur));
        if (same && pass > 0) break;
      }
    }
    if (s < C.substeps)
      for (int l = 0; l < 4; ++l) advance2<W>(L[l].p, C, L[l].mT, L[l].mA);
  }
  for (int l = 0; l < 4; ++l) {
    const Rig r = unpack_rig(L[l].p);
    if (l == 0) body_to(r.T, pos, rot, vel, ang, 0);
    body_to(r.A, pos, rot, vel, ang, 1 + 2 * l);
    body_to(r.B, pos, rot, vel, ang, 2 + 2 * l);
    // Info.contact of the lower leg lives in the lane's registers; torso / Aux accumulated into the row
    row_cv[3 * (2 + 2 * l)] = L[l].Bv.x; row_cv[3 * (2 + 2 * l) + 1] = L[l].Bv.y; row_cv[3 * (2 + 2 * l) + 2] = L[l].Bv.z;
    row_ca[3 * (2 + 2 * l)] = L[l].Bw.x; row_ca[3 * (2 + 2 * l) + 1] = L[l].Bw.y; row_ca[3 * (2 + 2 * l) + 2] = L[l].Bw.z;
  }
  std::memcpy(cv, row_cv.data(), sizeof(float) * 3 * nb);
  std::memcpy(ca, row_ca.data(), sizeof(float) * 3 * nb);
  return 0;
}
}  // namespace

extern "C" {

// DevConst + wall tables from the product's own host code (api.cu: build_dev_const). Returns 0 / non-zero
// (pobrax_last_error() of THIS library has the message).
int emu_create(const PobraxParams* p, void** handle) {
  EmuHandle* h = new EmuHandle();
  if (int rc = build_dev_const(p, &h->C, &h->sdf, &h->grid, &h->walls, &h->tip)) { delete h; return rc; }
  h->C.walls = h->walls.data();
  h->C.wall_tex = (unsigned long long)(uintptr_t)h->sdf.data();
  h->C.tip_tex = (unsigned long long)(uintptr_t)h->tip.data();
  *handle = h;
  return 0;
}

void emu_destroy(void* handle) { delete static_cast<EmuHandle*>(handle); }

int emu_num_bodies(void* handle) { return static_cast<EmuHandle*>(handle)->C.nb; }

// The lower leg's capsule-end table (ant_physics.cuh tip_mask_at) at n points, and the handle's wall boxes, for the
// exact-cull property test (tests/test_host_emu.py).
int emu_tip_masks(void* handle, long n, const float* xy, unsigned* out) {
  const DevConst& C = static_cast<EmuHandle*>(handle)->C;
  if (C.n_walls <= 0) return 1;
  for (long i = 0; i < n; ++i) out[i] = tip_mask_at(C, xy[2 * i], xy[2 * i + 1]);
  return 0;
}
int emu_walls(void* handle, float* lo_hi /* [n_walls][6] */) {
  const EmuHandle* h = static_cast<EmuHandle*>(handle);
  for (int w = 0; w < h->C.n_walls; ++w) {
    const float4 l = h->walls[2 * w], u = h->walls[2 * w + 1];
    const float v[6] = {l.x, l.y, l.z, u.x, u.y, u.z};
    for (int c = 0; c < 6; ++c) lo_hi[6 * w + c] = v[c];
  }
  return h->C.n_walls;
}

// brax.System.step on n envs, in place: QP arrays [n][nb][3|4] (brax body order; only the 9 ant bodies move),
// act [n][8], cv / ca [n][nb][3] = Info.contact.vel / .ang summed over the substeps (unclipped).
int emu_step(void* handle, long n, float* pos, float* rot, float* vel, float* ang, const float* act, float* cv,
             float* ca) {
  const EmuHandle* h = static_cast<EmuHandle*>(handle);
  const DevConst& C = h->C;
  const int nb = C.nb;
  const bool walls = C.env_kind != POBRAX_ANT;   // the template argument the step kernels use (KIND != POBRAX_ANT)
  for (long e = 0; e < n; ++e) {
    float *p = pos + e * nb * 3, *q = rot + e * nb * 4, *v = vel + e * nb * 3, *w = ang + e * nb * 3;
    const int rc = walls ? step_env<true>(C, p, q, v, w, act + e * 8, cv + e * nb * 3, ca + e * nb * 3)
                         : step_env<false>(C, p, q, v, w, act + e * 8, cv + e * nb * 3, ca + e * nb * 3);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
