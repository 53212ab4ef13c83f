// C ABI of libpobrax.so (include/pobrax.h): parameter tables, validation, handles, launches.
// No torch types; all buffers are caller-owned device pointers.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pobrax.h"
#include "dev_const.h"
#include "threefry.cuh"

namespace pobrax {
cudaError_t launch_step(const DevConst& C, const PobraxState& S, const float* action, cudaStream_t st);
cudaError_t launch_reset(const DevConst& C, const PobraxState& S, const uint32_t* keys, const float2* grid,
                         int only_done, uint32_t* chain, cudaStream_t st);
cudaError_t launch_unpack(const DevConst& C, const float* qp, const float* aux, float* pos, float* rot, float* vel,
                          float* ang, cudaStream_t st);
cudaError_t launch_pack(const DevConst& C, const float* pos, const float* rot, const float* vel, const float* ang,
                        float* qp, float* aux, cudaStream_t st);
cudaError_t launch_split_keys(const uint32_t key[2], int n, int first, int count, uint32_t* out, cudaStream_t st);
cudaError_t launch_fma_probe(float* out, int blocks, int iters, cudaStream_t st);
cudaError_t launch_eval_update(const float* reward, const float* done, float* ret, float* dret, long long* len, float* disc,
                               double* sums, float discount, int n, cudaStream_t st);
cudaError_t launch_split_pairs(const uint32_t* keys, int n, uint32_t* a, uint32_t* b, cudaStream_t st);
cudaError_t setup_device(DevConst& C, size_t smem_limit, const char** what);

struct Handle {
  DevConst C;
  int device;
  cudaArray_t sdf_array = nullptr;       // wall candidate mask tables: layered 2D array behind a texture object
  cudaTextureObject_t sdf_tex = 0;
  cudaArray_t tip_array = nullptr;       // lower-leg capsule-end candidate masks (finer cells), 2D array + texture object
  cudaTextureObject_t tip_tex = 0;
  float4* walls = nullptr;
  float2* grid = nullptr;
  uint8_t* tag_choice = nullptr;         // Tag: the step's opponent moves (dev_const.h)
};
}  // namespace pobrax

using pobrax::DevConst;
using pobrax::Handle;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
static int fail_cuda(const char* what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 2;
}

// ------------------------------------------------------------------------------------------ defaults
static void add_wall(PobraxParams* p, float x0, float y0, float x1, float y1, float width, float half_height) {
  // utils.py:6-28 add_box_wall_to_body for an axis-aligned segment: box centred on the midpoint,
  // half extents (len/2, width) along / across the segment, +-half_height around the Arena body z.
  const int w = p->num_walls++;
  const float mx = 0.5f * (x0 + x1), my = 0.5f * (y0 + y1);
  const float len = std::sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
  const bool horizontal = std::fabs(y1 - y0) < std::fabs(x1 - x0);
  const float hx = horizontal ? 0.5f * len : width, hy = horizontal ? width : 0.5f * len;
  p->wall_lo[w][0] = mx - hx; p->wall_hi[w][0] = mx + hx;
  p->wall_lo[w][1] = my - hy; p->wall_hi[w][1] = my + hy;
  p->wall_lo[w][2] = p->arena_z - half_height; p->wall_hi[w][2] = p->arena_z + half_height;
}

extern "C" int pobrax_draw_arena(PobraxParams* p, float x, float y, float half_height) {
  if (!p) return fail("pobrax_draw_arena: null params");
  p->num_walls = 0;
  p->arena_z = half_height;
  const float r = half_height / 2;
  const float pts[4][2] = {{x + r, y + r}, {x + r, -y - r}, {-x - r, -y - r}, {-x - r, y + r}};
  for (int i = 0; i < 4; ++i) add_wall(p, pts[i][0], pts[i][1], pts[(i + 1) % 4][0], pts[(i + 1) % 4][1], r, half_height);
  return 0;
}

extern "C" int pobrax_draw_t_maze(PobraxParams* p, float t_x, float t_y, float w, float half_height) {
  if (!p) return fail("pobrax_draw_t_maze: null params");
  p->num_walls = 0;
  p->arena_z = half_height;
  const float r = half_height;
  const float pts[8][2] = {{-t_x - r, t_y + r}, {t_x + r, t_y + r}, {t_x + r, t_y - w - r}, {w + r, t_y - w - r},
                           {w + r, -r}, {-w - r, -r}, {-w - r, t_y - w - r}, {-t_x - r, t_y - w - r}};
  for (int i = 0; i < 8; ++i) add_wall(p, pts[i][0], pts[i][1], pts[(i + 1) % 8][0], pts[(i + 1) % 8][1], r, half_height);
  return 0;
}

extern "C" int pobrax_abi_version(void) { return POBRAX_ABI_VERSION; }
extern "C" const char* pobrax_last_error(void) { return g_err.c_str(); }
extern "C" int pobrax_struct_sizes(int32_t* params, int32_t* state, int32_t* layout) {
  if (params) *params = (int32_t)sizeof(PobraxParams);
  if (state) *state = (int32_t)sizeof(PobraxState);
  if (layout) *layout = (int32_t)sizeof(PobraxLayout);
  return 0;
}

extern "C" int pobrax_default_params(int env_kind, PobraxParams* p) {
  if (!p) return fail("pobrax_default_params: null out");
  if (env_kind < POBRAX_ANT || env_kind > POBRAX_ANT_TAG) return fail("pobrax_default_params: unknown env_kind");
  std::memset(p, 0, sizeof(*p));
  p->env_kind = env_kind;
  p->num_envs = 1;
  p->episode_length = 1000;
  p->auto_reset = POBRAX_AUTORESET_CACHED;
  p->action_repeat = 1;
  p->track_metrics = 0;
  // brax.envs.ant._SYSTEM_CONFIG (legacy spring era; tests/golden/ant_tag_config.json is the reference's own copy)
  p->dt = 0.05f; p->substeps = 10; p->gravity_z = -9.8f;
  p->velocity_damping = 0.0f; p->angular_damping = -0.05f; p->baumgarte_erp = 0.1f;
  p->friction = 1.0f; p->elasticity = 0.0f;
  p->torso_mass = 10.0f; p->leg_mass = 1.0f; p->torso_radius = 0.25f; p->leg_radius = 0.08f;
  p->aux_length = 0.44284272f; p->foot_length = 0.7256854f;
  const float sx[4] = {1, -1, -1, 1}, sy[4] = {1, 1, -1, -1};
  const float ceul[4][3] = {{90, -45, 0}, {90, 45, 0}, {-90, 45, 0}, {-90, -45, 0}};
  const float aeul[4] = {135, 45, 135, 45};
  const float alim[4][2] = {{30, 70}, {-70, -30}, {-70, -30}, {30, 70}};
  for (int l = 0; l < 4; ++l) {
    for (int c = 0; c < 3; ++c) p->collider_euler[l][c] = ceul[l][c];
    p->hip_off_p[l][0] = 0.2f * sx[l]; p->hip_off_p[l][1] = 0.2f * sy[l];
    p->hip_off_c[l][0] = -0.1f * sx[l]; p->hip_off_c[l][1] = -0.1f * sy[l];
    p->ank_off_p[l][0] = 0.1f * sx[l]; p->ank_off_p[l][1] = 0.1f * sy[l];
    p->ank_off_c[l][0] = -0.2f * sx[l]; p->ank_off_c[l][1] = -0.2f * sy[l];
    p->hip_euler[l][1] = -90.0f;
    p->ank_euler[l][2] = aeul[l];
    p->hip_limit[l][0] = -30.0f; p->hip_limit[l][1] = 30.0f;
    p->ank_limit[l][0] = alim[l][0]; p->ank_limit[l][1] = alim[l][1];
  }
  p->joint_stiffness = 18000.0f; p->joint_spring_damping = 80.0f; p->joint_angular_damping = 20.0f;
  p->joint_limit_strength = 18000.0f;
  p->actuator_strength = 350.0f;
  p->arena_z = 0.5f;
  // task defaults (ant_heavenhell.py:51-56, ant_tag.py:38-45, ant_gather.py:59-69)
  p->n_apples = 8; p->n_bombs = 8; p->n_bins = 10;
  p->catch_range = 1.0f; p->sensor_range = 6.0f; p->sensor_span = 3.14159265358979323846f;
  p->robot_object_spacing = 2.0f; p->gather_cage_xy[0] = 6.0f; p->gather_cage_xy[1] = 6.0f;
  p->tag_radius = 1.5f; p->target_step = 0.5f; p->min_spawn_distance = 5.0f;
  p->cage_xy[0] = 4.5f; p->cage_xy[1] = 4.5f;
  p->heaven_hell_xy[0][0] = -5.25f; p->heaven_hell_xy[0][1] = 7.0f;
  p->heaven_hell_xy[1][0] = 5.25f; p->heaven_hell_xy[1][1] = 7.0f;
  p->priest_xy[0] = 0.0f; p->priest_xy[1] = 7.0f;
  switch (env_kind) {
    case POBRAX_ANT:
      p->dying_cost = 0.0f;
      break;
    case POBRAX_ANT_HEAVENHELL: {
      p->dying_cost = -2.0f; p->visible_radius = 2.0f;
      p->init_lo[0] = -0.5f; p->init_lo[1] = 0.5f; p->init_hi[0] = 0.5f; p->init_hi[1] = 1.5f;
      const float hw = 2.0f;  // ant_heavenhell.py:32-33: t_x = max x + w/2, t_y = max y + w/2
      pobrax_draw_t_maze(p, 5.25f + hw / 2, 7.0f + hw / 2, hw, 0.5f);
      break;
    }
    case POBRAX_ANT_GATHER:
      p->dying_cost = -10.0f;
      pobrax_draw_arena(p, 6.0f + 1.0f, 6.0f + 1.0f, 0.5f);  // ant_gather.py:25: cage + offset 1
      break;
    case POBRAX_ANT_TAG:
      p->dying_cost = -1.0f; p->visible_radius = 3.0f;
      p->init_lo[0] = -4.5f; p->init_lo[1] = -4.5f; p->init_hi[0] = 4.5f; p->init_hi[1] = 4.5f;
      pobrax_draw_arena(p, 4.5f + 1.0f, 4.5f + 1.0f, 0.5f);  // ant_tag.py:22
      break;
  }
  return 0;
}

// ---------------------------------------------------------------------------------- params -> DevConst
static void euler_to_quat(const float deg[3], double q[4]) {
  double c[3], s[3];
  for (int i = 0; i < 3; ++i) { c[i] = std::cos(deg[i] * M_PI / 360.0); s[i] = std::sin(deg[i] * M_PI / 360.0); }
  q[0] = c[0] * c[1] * c[2] - s[0] * s[1] * s[2];
  q[1] = s[0] * c[1] * c[2] + c[0] * s[1] * s[2];
  q[2] = c[0] * s[1] * c[2] - s[0] * c[1] * s[2];
  q[3] = c[0] * c[1] * s[2] + s[0] * s[1] * c[2];
}
static void qrotate(const double v[3], const double q[4], double o[3]) {
  const double s = q[0], u[3] = {q[1], q[2], q[3]};
  const double uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  const double cx[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
  for (int i = 0; i < 3; ++i) o[i] = 2 * uv * u[i] + (s * s - uu) * v[i] + 2 * s * cx[i];
}
static bool near(double a, double b, double tol = 1e-5) { return std::fabs(a - b) <= tol; }

static int layout_of(const PobraxParams* p, PobraxLayout* L) {
  const int n_obj = p->n_apples + p->n_bombs;
  switch (p->env_kind) {
    case POBRAX_ANT: L->num_bodies = 10; L->aux_dim = 0; L->metrics_dim = 4; L->obs_dim = 27 + 6 * 10; break;
    case POBRAX_ANT_HEAVENHELL: L->num_bodies = 14; L->aux_dim = 3; L->metrics_dim = 1; L->obs_dim = 29 + 6 * 14 + 1; break;
    case POBRAX_ANT_TAG: L->num_bodies = 12; L->aux_dim = 5; L->metrics_dim = 1; L->obs_dim = 29 + 6 * 12 + 2; break;
    case POBRAX_ANT_GATHER:
      if (p->n_apples < 0 || p->n_bombs < 0 || n_obj < 1 || n_obj > pobrax::kMaxObjects)
        return fail("gather: n_apples + n_bombs must be in [1, 16]");
      if (p->n_bins < 1 || 2 * p->n_bins > pobrax::kMaxBins) return fail("gather: n_bins must be in [1, 16]");
      if (n_obj > 0 && p->n_apples + p->n_bins > 2 * p->n_bins && p->n_bombs > 0)
        return fail("gather: bomb bins (offset by n_apples) would overflow the 2*n_bins readings");
      L->num_bodies = 11 + n_obj; L->aux_dim = 3 * n_obj; L->metrics_dim = 2;
      L->obs_dim = 29 + 6 * L->num_bodies + 2 * p->n_bins;
      break;
    default: return fail("unknown env_kind");
  }
  if (p->obs_col_lo != 0 || p->obs_col_hi != 0) {
    if (p->obs_col_lo < 0 || p->obs_col_hi > L->obs_dim || p->obs_col_lo >= p->obs_col_hi)
      return fail("obs_col_lo/obs_col_hi must select a non-empty column range inside the observation");
    L->obs_dim = p->obs_col_hi - p->obs_col_lo;
  }
  L->action_dim = 8;
  L->qp_planes = POBRAX_QP_PLANES;
  return 0;
}

extern "C" int pobrax_layout(const PobraxParams* p, PobraxLayout* out) {
  if (!p || !out) return fail("pobrax_layout: null argument");
  return layout_of(p, out);
}

static std::vector<float2> gather_grid(const PobraxParams* p) {
  // ant_gather.py:88-90: meshgrid(arange(-cage_x, cage_x + 1), arange(-cage_y, cage_y + 1)) in 'xy' order (y-major, x
  // fastest), |g| > spacing. arange keeps a fractional origin (cage 4.5 -> -4.5, -3.5, ..., 4.5), like the reference.
  std::vector<float2> g;
  const double cx = p->gather_cage_xy[0], cy = p->gather_cage_xy[1];
  const int nx = (int)std::ceil((cx + 1.0) - (-cx) - 1e-9), ny = (int)std::ceil((cy + 1.0) - (-cy) - 1e-9);
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
      const float fx = (float)(-cx + ix), fy = (float)(-cy + iy);
      if (std::sqrt(fx * fx + fy * fy) > p->robot_object_spacing) g.push_back(make_float2(fx, fy));
    }
  return g;
}

static int build_dev_const(const PobraxParams* p, DevConst* Cp, std::vector<uint8_t>* sdf, std::vector<float2>* grid,
                           std::vector<float4>* walls, std::vector<uint8_t>* tip) {
  DevConst& C = *Cp;
  std::memset(&C, 0, sizeof(C));
  PobraxLayout L;
  if (int rc = layout_of(p, &L)) return rc;
  if (p->num_envs <= 0) return fail("num_envs must be positive");
  if (p->action_repeat < 1) return fail("action_repeat must be >= 1");
  if (p->substeps < 1 || !(p->dt > 0.0f)) return fail("dt and substeps must be positive");
  if (p->num_walls < 0 || p->num_walls > POBRAX_MAX_WALLS) return fail("num_walls out of range");
  if (p->auto_reset != POBRAX_AUTORESET_OFF && p->auto_reset != POBRAX_AUTORESET_CACHED) return fail("unknown auto_reset mode");
  C.n_envs = p->num_envs; C.env_kind = p->env_kind; C.nb = L.num_bodies;
  {  // full observation width (staging) vs the emitted column range
    PobraxParams full = *p;
    full.obs_col_lo = full.obs_col_hi = 0;
    PobraxLayout Lf;
    if (int rc = layout_of(&full, &Lf)) return rc;
    C.obs_dim = Lf.obs_dim;
    C.obs_lo = (p->obs_col_lo == 0 && p->obs_col_hi == 0) ? 0 : p->obs_col_lo;
    C.obs_out = L.obs_dim;
  }
  C.aux_dim = L.aux_dim; C.metrics_dim = L.metrics_dim;
  C.episode_length = p->episode_length; C.auto_reset = p->auto_reset; C.track_metrics = p->track_metrics;
  C.has_rng = p->env_kind != POBRAX_ANT;
  // ActionRepeatWrapper (wrappers.py:16-24): dt *= k, substeps *= k
  C.substeps = p->substeps * p->action_repeat;
  C.dt = p->dt * (float)p->action_repeat;
  C.h = (float)((double)p->dt / (double)p->substeps);
  C.gravity_z = p->gravity_z;
  C.vel_damp = std::exp(p->velocity_damping * C.h);
  C.ang_damp = std::exp(p->angular_damping * C.h);
  C.baumgarte = (float)((double)p->baumgarte_erp * p->substeps / (double)p->dt);
  C.friction = p->friction; C.elasticity = p->elasticity;
  C.m_torso = p->torso_mass; C.m_leg = p->leg_mass;
  C.inv_m_torso = 1.0f / p->torso_mass; C.inv_m_leg = 1.0f / p->leg_mass;
  C.r_torso = p->torso_radius; C.r_leg = p->leg_radius;
  C.k_joint = p->joint_stiffness; C.sd_joint = p->joint_spring_damping; C.ad_joint = p->joint_angular_damping;
  C.ls_joint = p->joint_limit_strength; C.act_strength = p->actuator_strength;
  C.h_k = C.h * C.k_joint; C.h_sd = C.h * C.sd_joint; C.h_ad = C.h * C.ad_joint; C.h_ls = C.h * C.ls_joint;
  C.h_act = C.h * C.act_strength; C.h_g = C.h * C.gravity_z;
  C.seg_aux = p->aux_length / 2 - p->leg_radius; C.seg_foot = p->foot_length / 2 - p->leg_radius;
  // ---- factor the leg geometry: offsets / capsule ends = scalar * u[l]
  double s_hc = 0, s_ap = 0, s_ac = 0, s_ft = 0, s_ax = 0;
  for (int l = 0; l < 4; ++l) {
    const double ux = p->hip_off_p[l][0], uy = p->hip_off_p[l][1], uu = ux * ux + uy * uy;
    if (uu < 1e-12 || p->hip_off_p[l][2] != 0 || p->hip_off_c[l][2] != 0 || p->ank_off_p[l][2] != 0 || p->ank_off_c[l][2] != 0)
      return fail("unsupported Ant geometry: joint offsets must lie in the body xy-plane");
    C.leg_u[l][0] = (float)ux; C.leg_u[l][1] = (float)uy;
    auto scale = [&](const float v[3], double* s_out) -> bool {
      const double s = (v[0] * ux + v[1] * uy) / uu;
      if (!near(v[0], s * ux) || !near(v[1], s * uy)) return false;
      if (l == 0) *s_out = s;
      return near(*s_out, s);
    };
    if (!scale(p->hip_off_c[l], &s_hc) || !scale(p->ank_off_p[l], &s_ap) || !scale(p->ank_off_c[l], &s_ac))
      return fail("unsupported Ant geometry: joint offsets must be uniform multiples of the leg direction");
    double q[4], ez[3] = {0, 0, 1}, ex[3] = {1, 0, 0}, axis[3], a0[3], a2[3];
    euler_to_quat(p->collider_euler[l], q);
    qrotate(ez, q, axis);
    if (!near(axis[2], 0.0)) return fail("unsupported Ant geometry: leg capsules must lie in the body xy-plane");
    const float foot[3] = {(float)(-axis[0] * C.seg_foot), (float)(-axis[1] * C.seg_foot), 0.f};
    const float aux[3] = {(float)(-axis[0] * C.seg_aux), (float)(-axis[1] * C.seg_aux), 0.f};
    if (!scale(foot, &s_ft) || !scale(aux, &s_ax))
      return fail("unsupported Ant geometry: leg capsules must be aligned with the leg direction");
    euler_to_quat(p->hip_euler[l], q);
    qrotate(ex, q, a0); qrotate(ez, q, a2);
    if (!near(a0[0], 0) || !near(a0[1], 0) || !near(a0[2], 1) || !near(a2[0], -1) || !near(a2[1], 0) || !near(a2[2], 0))
      return fail("unsupported Ant geometry: hip joints must have axis +z and reference -x (euler (0,-90,0))");
    euler_to_quat(p->ank_euler[l], q);
    qrotate(ex, q, a0); qrotate(ez, q, a2);
    if (!near(a0[2], 0) || !near(a2[0], 0) || !near(a2[1], 0) || !near(a2[2], 1))
      return fail("unsupported Ant geometry: ankle joints must rotate about an axis in the xy-plane (euler (0,0,phi))");
    C.ank_ax[l][0] = (float)a0[0]; C.ank_ax[l][1] = (float)a0[1];
    const float d2r = 3.14159265358979323846f;
    if (l > 0 && (p->hip_limit[l][0] != p->hip_limit[0][0] || p->hip_limit[l][1] != p->hip_limit[0][1]))
      return fail("unsupported Ant geometry: hip limits must be equal for all legs");
    C.ank_lo[l] = p->ank_limit[l][0] * d2r / 180.0f; C.ank_hi[l] = p->ank_limit[l][1] * d2r / 180.0f;
    C.ank_default[l] = (C.ank_lo[l] + C.ank_hi[l]) / 2.0f;
    if (l == 0) {
      C.hip_lo = p->hip_limit[0][0] * d2r / 180.0f; C.hip_hi = p->hip_limit[0][1] * d2r / 180.0f;
      C.hip_default = (C.hip_lo + C.hip_hi) / 2.0f;
    }
  }
  C.s_hip_p = 1.0f; C.s_hip_c = (float)s_hc; C.s_ank_p = (float)s_ap; C.s_ank_c = (float)s_ac;
  C.s_foot = (float)s_ft; C.s_aux = (float)s_ax;
  // ---- walls + conservative distance field
  C.n_walls = (p->env_kind == POBRAX_ANT) ? 0 : p->num_walls;
  C.arena_z = p->arena_z;
  float wlo[pobrax::kMaxWalls][3], whi[pobrax::kMaxWalls][3];
  walls->clear();
  for (int w = 0; w < C.n_walls; ++w) {
    for (int c = 0; c < 3; ++c) {
      if (!(p->wall_lo[w][c] <= p->wall_hi[w][c])) return fail("wall box with lo > hi");
      wlo[w][c] = p->wall_lo[w][c]; whi[w][c] = p->wall_hi[w][c];
    }
    walls->push_back(make_float4(wlo[w][0], wlo[w][1], wlo[w][2], 0.f));
    C.wall_box[w][0] = make_float4(wlo[w][0], wlo[w][1], wlo[w][2], 0.f);
    C.wall_box[w][1] = make_float4(whi[w][0], whi[w][1], whi[w][2], 0.f);
    walls->push_back(make_float4(whi[w][0], whi[w][1], whi[w][2], 0.f));
  }
  sdf->clear();
  if (C.n_walls > 0) {
    const double cell = 0.125, margin = 2.0;
    double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
    for (int w = 0; w < C.n_walls; ++w) {
      x0 = std::fmin(x0, wlo[w][0]); y0 = std::fmin(y0, wlo[w][1]);
      x1 = std::fmax(x1, whi[w][0]); y1 = std::fmax(y1, whi[w][1]);
    }
    x0 -= margin; y0 -= margin; x1 += margin; y1 += margin;
    const int nx = (int)std::ceil((x1 - x0) / cell), ny = (int)std::ceil((y1 - y0) / cell);
    if (nx <= 2 || ny <= 2 || (long long)nx * ny > (1 << 22)) return fail("wall extent unsupported (mask table too large)");
    // Bit w of a cell (per body type): some point of the cell is within that body's reach (capsule half
    // segment + radius, plus 2 mm of slack for the cell lookup: float rounding of the coordinate, and a texture
    // unit that may quantise the coordinate to 1/256 of a 125 mm cell before flooring) of wall w in the xy-plane. Border cells (and
    // everything outside the table, which clamps onto them) list every wall.
    const uint8_t all = (uint8_t)((1u << C.n_walls) - 1u);
    const size_t plane = (size_t)nx * ny;
    sdf->assign(2 * plane, all);
    const double reach[2] = {C.r_torso + 2e-3, C.seg_aux + C.r_leg + 2e-3};   // (the lower leg has its own table below)
    for (int k = 0; k < 2; ++k)
      for (int iy = 1; iy < ny - 1; ++iy)
        for (int ix = 1; ix < nx - 1; ++ix) {
          const double cx0 = x0 + ix * cell - 1e-4, cx1 = x0 + (ix + 1) * cell + 1e-4;
          const double cy0 = y0 + iy * cell - 1e-4, cy1 = y0 + (iy + 1) * cell + 1e-4;
          uint8_t m = 0;
          for (int w = 0; w < C.n_walls; ++w) {
            const double dx = std::fmax(std::fmax(wlo[w][0] - cx1, 0.0), cx0 - whi[w][0]);
            const double dy = std::fmax(std::fmax(wlo[w][1] - cy1, 0.0), cy0 - whi[w][1]);
            if (std::sqrt(dx * dx + dy * dy) <= reach[k]) m |= (uint8_t)(1u << w);
          }
          (*sdf)[k * plane + (size_t)iy * nx + ix] = m;
        }
    C.sdf_x0 = (float)x0; C.sdf_y0 = (float)y0; C.sdf_inv_cell = (float)(1.0 / cell);
    C.sdf_bx = (float)(-x0 / cell); C.sdf_by = (float)(-y0 / cell);
    C.sdf_nx = nx; C.sdf_ny = ny;
  }
  tip->clear();
  if (C.n_walls > 0) {
    // Capsule-END table of the lower leg (ant_physics.cuh tip_mask_at): bit w of a cell <=> some point of the cell is
    // within r_leg of wall w's footprint, or within seg_foot + r_leg of one of the footprint's four vertices (the
    // end lookups of a capsule then cover its interior points too, see tip_mask_at). 1/32 m cells; 1 mm of slack for
    // float rounding and the texture unit's coordinate quantisation; border cells list every wall.
    const double cell = 1.0 / 32.0, margin = 0.75, slack = 1e-3;
    double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
    for (int w = 0; w < C.n_walls; ++w) {
      x0 = std::fmin(x0, wlo[w][0]); y0 = std::fmin(y0, wlo[w][1]);
      x1 = std::fmax(x1, whi[w][0]); y1 = std::fmax(y1, whi[w][1]);
    }
    x0 -= margin; y0 -= margin; x1 += margin; y1 += margin;
    const int nx = (int)std::ceil((x1 - x0) / cell), ny = (int)std::ceil((y1 - y0) / cell);
    if ((long long)nx * ny > (1 << 24)) return fail("wall extent unsupported (capsule-end table too large)");
    const uint8_t all = (uint8_t)((1u << C.n_walls) - 1u);
    tip->assign((size_t)nx * ny, all);
    const double r_face = C.r_leg + slack, r_vert = C.seg_foot + C.r_leg + slack;
    for (int iy = 1; iy < ny - 1; ++iy)
      for (int ix = 1; ix < nx - 1; ++ix) {
        const double cx0 = x0 + ix * cell - 1e-4, cx1 = x0 + (ix + 1) * cell + 1e-4;
        const double cy0 = y0 + iy * cell - 1e-4, cy1 = y0 + (iy + 1) * cell + 1e-4;
        uint8_t m = 0;
        for (int w = 0; w < C.n_walls; ++w) {
          const double dx = std::fmax(std::fmax(wlo[w][0] - cx1, 0.0), cx0 - whi[w][0]);
          const double dy = std::fmax(std::fmax(wlo[w][1] - cy1, 0.0), cy0 - whi[w][1]);
          bool near = std::sqrt(dx * dx + dy * dy) <= r_face;
          for (int v = 0; v < 4 && !near; ++v) {
            const double vx = (v & 1) ? whi[w][0] : wlo[w][0], vy = (v & 2) ? whi[w][1] : wlo[w][1];
            const double ex = std::fmax(std::fmax(cx0 - vx, 0.0), vx - cx1), ey = std::fmax(std::fmax(cy0 - vy, 0.0), vy - cy1);
            near = std::sqrt(ex * ex + ey * ey) <= r_vert;
          }
          if (near) m |= (uint8_t)(1u << w);
        }
        (*tip)[(size_t)iy * nx + ix] = m;
      }
    C.tip_inv_cell = (float)(1.0 / cell);
    C.tip_bx = (float)(-x0 / cell); C.tip_by = (float)(-y0 / cell);
    C.tip_nx = nx; C.tip_ny = ny;
  }
  // ---- task
  C.dying_cost = p->dying_cost; C.visible_radius = p->visible_radius;
  for (int i = 0; i < 2; ++i) {
    C.hh_xy[i][0] = p->heaven_hell_xy[i][0]; C.hh_xy[i][1] = p->heaven_hell_xy[i][1];
    C.priest_xy[i] = p->priest_xy[i];
    C.init_lo[i] = p->init_lo[i]; C.init_hi[i] = p->init_hi[i];
    C.cage_xy[i] = p->cage_xy[i];
  }
  C.hh_z = 1.0f; C.priest_z = 1.0f;  // ant_heavenhell.py:58 (jp.ones((3,1))) and :21
  C.tag_radius = p->tag_radius; C.target_step = p->target_step; C.min_spawn = p->min_spawn_distance;
  C.n_apples = p->n_apples; C.n_bombs = p->n_bombs; C.n_bins = p->n_bins;
  C.catch_range = p->catch_range; C.sensor_range = p->sensor_range;
  C.half_span = p->sensor_span / 2.0f;
  C.bin_res = (2.0f * C.half_span) / (float)p->n_bins;
  C.spacing = p->robot_object_spacing;
  grid->clear();
  if (p->env_kind == POBRAX_ANT_GATHER) {
    if (!(p->gather_cage_xy[0] >= 0.0f) || !(p->gather_cage_xy[1] >= 0.0f) || p->gather_cage_xy[0] > 64.0f ||
        p->gather_cage_xy[1] > 64.0f)
      return fail("gather: cage_xy must be in [0, 64]");
    *grid = gather_grid(p);
    if ((int)grid->size() < p->n_apples + p->n_bombs) return fail("gather: fewer grid cells than objects");
    if (grid->size() > 1024) return fail("gather: cage too large (more than 1024 candidate cells)");
    C.n_grid = (int)grid->size();
    // ant_gather.py:91: waiting_area = last grid position + 2*sensor_range
    C.waiting[0] = grid->back().x + p->sensor_range * 2; C.waiting[1] = grid->back().y + p->sensor_range * 2;
    C.waiting[2] = 0.0f + p->sensor_range * 2;
  }
  return 0;
}

// -------------------------------------------------------------------------------------------- handles
extern "C" int pobrax_create(const PobraxParams* p, int device, void** handle) {
  if (!p || !handle) return fail("pobrax_create: null argument");
  *handle = nullptr;
  DevConst C;
  std::vector<uint8_t> sdf;
  std::vector<float2> grid;
  std::vector<float4> walls;
  std::vector<uint8_t> tip;
  if (int rc = build_dev_const(p, &C, &sdf, &grid, &walls, &tip)) return rc;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail_cuda("pobrax_create: no CUDA device (this library has no CPU path)", e);
  if (device < 0 || device >= count) return fail("pobrax_create: device index out of range");
  int prev = 0;
  cudaGetDevice(&prev);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail_cuda("cudaSetDevice", e);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10) { cudaSetDevice(prev); return fail("pobrax_create: kernels are built for sm_100a (B200) only"); }
  Handle* h = new Handle();
  h->device = device;
  if (!grid.empty()) {
    if ((e = cudaMalloc(&h->grid, grid.size() * sizeof(float2))) != cudaSuccess) { delete h; cudaSetDevice(prev); return fail_cuda("cudaMalloc(grid)", e); }
    cudaMemcpy(h->grid, grid.data(), grid.size() * sizeof(float2), cudaMemcpyHostToDevice);
  }
  if (!walls.empty()) {
    if ((e = cudaMalloc(&h->walls, walls.size() * sizeof(float4))) != cudaSuccess) { cudaFree(h->grid); delete h; cudaSetDevice(prev); return fail_cuda("cudaMalloc(walls)", e); }
    cudaMemcpy(h->walls, walls.data(), walls.size() * sizeof(float4), cudaMemcpyHostToDevice);
  }
  if (!sdf.empty()) {  // layered 2D texture over the two body-centre tables (torso, Aux) (point sampling, clamp, unnormalised coordinates)
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaExtent ext = make_cudaExtent((size_t)C.sdf_nx, (size_t)C.sdf_ny, 2);
    if ((e = cudaMalloc3DArray(&h->sdf_array, &fmt, ext, cudaArrayLayered)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaMalloc3DArray(wall masks)", e);
    }
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(sdf.data(), (size_t)C.sdf_nx, (size_t)C.sdf_nx, (size_t)C.sdf_ny);
    cp.dstArray = h->sdf_array;
    cp.extent = ext;
    cp.kind = cudaMemcpyHostToDevice;
    if ((e = cudaMemcpy3D(&cp)) != cudaSuccess) { cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaMemcpy3D(wall masks)", e); }
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = h->sdf_array;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    if ((e = cudaCreateTextureObject(&h->sdf_tex, &res, &td, nullptr)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaCreateTextureObject(wall masks)", e);
    }
    C.wall_tex = (unsigned long long)h->sdf_tex;
  }
  if (!tip.empty()) {  // capsule-end table: plain 2D texture (point sampling, clamp, unnormalised coordinates)
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    if ((e = cudaMallocArray(&h->tip_array, &fmt, (size_t)C.tip_nx, (size_t)C.tip_ny)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaMallocArray(capsule-end masks)", e);
    }
    if ((e = cudaMemcpy2DToArray(h->tip_array, 0, 0, tip.data(), (size_t)C.tip_nx, (size_t)C.tip_nx, (size_t)C.tip_ny,
                                 cudaMemcpyHostToDevice)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaMemcpy2DToArray(capsule-end masks)", e);
    }
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = h->tip_array;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    if ((e = cudaCreateTextureObject(&h->tip_tex, &res, &td, nullptr)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaCreateTextureObject(capsule-end masks)", e);
    }
    C.tip_tex = (unsigned long long)h->tip_tex;
  }
  C.walls = h->walls;
  C.tag_choice = nullptr;
  if (C.env_kind == POBRAX_ANT_TAG) {
    if ((e = cudaMalloc(&h->tag_choice, (size_t)C.n_envs)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h); return fail_cuda("cudaMalloc(tag move scratch)", e);
    }
    C.tag_choice = h->tag_choice;
  }
  {  // per-device kernel attributes + occupancy for THIS handle's device (a process may hold handles on several GPUs)
    const char* what = "";
    if ((e = pobrax::setup_device(C, (size_t)prop.sharedMemPerBlockOptin, &what)) != cudaSuccess) {
      cudaSetDevice(prev); pobrax_destroy(h);
      return fail_cuda((std::string("pobrax_create: ") + what).c_str(), e);
    }
  }
  h->C = C;
  cudaSetDevice(prev);
  *handle = h;
  return 0;
}

extern "C" int pobrax_destroy(void* handle) {
  if (!handle) return 0;
  Handle* h = static_cast<Handle*>(handle);
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  if (h->sdf_tex) cudaDestroyTextureObject(h->sdf_tex);
  if (h->sdf_array) cudaFreeArray(h->sdf_array);
  if (h->tip_tex) cudaDestroyTextureObject(h->tip_tex);
  if (h->tip_array) cudaFreeArray(h->tip_array);
  if (h->walls) cudaFree(h->walls);
  if (h->grid) cudaFree(h->grid);
  if (h->tag_choice) cudaFree(h->tag_choice);
  cudaSetDevice(prev);
  delete h;
  return 0;
}

static int check_state(const Handle* h, const PobraxState* st, bool need_first) {
  const DevConst& C = h->C;
  if (!st) return fail("null state");
  if (!st->qp || !st->obs || !st->reward || !st->done || !st->steps || !st->truncation || !st->metrics)
    return fail("state: qp/obs/reward/done/steps/truncation/metrics must be non-null");
  if (C.aux_dim > 0 && !st->aux) return fail("state: aux is required for this env");
  if (C.has_rng && !st->rng) return fail("state: rng is required for this env");
  if (C.track_metrics && (!st->ep_return || !st->acc)) return fail("state: ep_return/acc required when track_metrics=1");
  if (need_first && C.auto_reset == POBRAX_AUTORESET_CACHED && (!st->first_qp || !st->first_obs || (C.aux_dim > 0 && !st->first_aux)))
    return fail("state: first_qp/first_aux/first_obs required when auto_reset=CACHED");
  return 0;
}

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int pobrax_reset(void* handle, const uint32_t* keys, PobraxState* st, void* stream) {
  if (!handle || !keys) return fail("pobrax_reset: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (int rc = check_state(h, st, true)) return rc;
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_reset(h->C, *st, keys, h->grid, 0, nullptr, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_reset launch", e);
}

extern "C" int pobrax_reset_where_done(void* handle, const uint32_t* keys, PobraxState* st, void* stream) {
  if (!handle || !keys) return fail("pobrax_reset_where_done: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (int rc = check_state(h, st, false)) return rc;
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_reset(h->C, *st, keys, h->grid, 1, nullptr, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_reset_where_done launch", e);
}

extern "C" int pobrax_reset_where_done_chain(void* handle, uint32_t* chain, PobraxState* st, void* stream) {
  if (!handle || !chain) return fail("pobrax_reset_where_done_chain: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (int rc = check_state(h, st, false)) return rc;
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_reset(h->C, *st, nullptr, h->grid, 1, chain, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_reset_where_done_chain launch", e);
}

extern "C" int pobrax_step(void* handle, PobraxState* st, const float* action, void* stream) {
  if (!handle || !action) return fail("pobrax_step: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (int rc = check_state(h, st, true)) return rc;
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_step(h->C, *st, action, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_step launch", e);
}

extern "C" int pobrax_unpack_qp(void* handle, const float* qp, const float* aux, float* pos, float* rot, float* vel,
                                float* ang, void* stream) {
  if (!handle || !qp || !pos || !rot || !vel || !ang) return fail("pobrax_unpack_qp: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (h->C.aux_dim > 0 && !aux) return fail("pobrax_unpack_qp: aux required");
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_unpack(h->C, qp, aux, pos, rot, vel, ang, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_unpack_qp launch", e);
}

extern "C" int pobrax_pack_qp(void* handle, const float* pos, const float* rot, const float* vel, const float* ang,
                              float* qp, float* aux, void* stream) {
  if (!handle || !qp || !pos || !rot || !vel || !ang) return fail("pobrax_pack_qp: null argument");
  Handle* h = static_cast<Handle*>(handle);
  if (h->C.aux_dim > 0 && !aux) return fail("pobrax_pack_qp: aux required");
  DeviceGuard g(h->device);
  cudaError_t e = pobrax::launch_pack(h->C, pos, rot, vel, ang, qp, aux, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_pack_qp launch", e);
}

extern "C" int pobrax_split_keys(const uint32_t key[2], int n, int first, int count, uint32_t* out, void* stream) {
  if (!key || !out) return fail("pobrax_split_keys: null argument");
  if (n <= 0 || first < 0 || count < 0 || first + count > n) return fail("pobrax_split_keys: bad range");
  cudaError_t e = pobrax::launch_split_keys(key, n, first, count, out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_split_keys launch", e);
}

extern "C" int pobrax_eval_update(const float* reward, const float* done, float* returns, float* disc_returns,
                                  long long* lengths, float* cur_discount, double* sums, float discount, int n,
                                  void* stream) {
  if (!reward || !done || !returns || !disc_returns || !lengths || !cur_discount || !sums)
    return fail("pobrax_eval_update: null argument");
  if (n <= 0) return fail("pobrax_eval_update: n must be positive");
  cudaError_t e = pobrax::launch_eval_update(reward, done, returns, disc_returns, lengths, cur_discount, sums, discount, n,
                                             static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_eval_update launch", e);
}

extern "C" int pobrax_fp32_probe(float* out, int blocks, int iters, void* stream, double* flops) {
  if (!out || blocks <= 0 || iters <= 0) return fail("pobrax_fp32_probe: bad argument");
  cudaError_t e = pobrax::launch_fma_probe(out, blocks, iters, static_cast<cudaStream_t>(stream));
  if (flops) *flops = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_fp32_probe launch", e);
}

extern "C" int pobrax_split_pairs(const uint32_t* keys, int n, uint32_t* out_a, uint32_t* out_b, void* stream) {
  if (!keys || !out_a || !out_b || n < 0) return fail("pobrax_split_pairs: bad argument");
  cudaError_t e = pobrax::launch_split_pairs(keys, n, out_a, out_b, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? 0 : fail_cuda("pobrax_split_pairs launch", e);
}
