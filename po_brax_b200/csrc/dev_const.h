// Device-side constant block derived on the host from PobraxParams (see ant_system.cpp).
// Passed to every kernel as a __grid_constant__ parameter (constant bank, LDC-indexable per lane).
#pragma once
#include <stdint.h>

namespace pobrax {

constexpr int kMaxWalls = 8;
constexpr int kQpPlanes = 32;   // float4 planes per env
constexpr int kTorsoPlanes = 4; // planes 0..3: torso (13 floats + 3 pad)
constexpr int kLegPlanes = 7;   // planes 4+7l .. 10+7l: Aux l (13) + lower l (13) + 2 pad
constexpr int kNumAcc = 8;

// aux rows (float[aux_dim][N])
//   Ant        : none
//   HeavenHell : 0 ground_x, 1 ground_y, 2 target_x (heaven side; hell is the other one)
//   Tag        : 0 ground_x, 1 ground_y, 2 tgt_x, 3 tgt_y, 4 tgt_z
//   Gather     : 3*k + {0,1,2} = object k xyz (k = 0..15: apples then bombs)
// metrics rows (float[metrics_dim][N])
//   Ant        : 0 reward_ctrl_cost, 1 reward_contact_cost, 2 reward_forward, 3 reward_survive
//   HeavenHell : 0 hits      Tag: 0 hits      Gather: 0 apples, 1 bombs

struct DevConst {
  int32_t n_envs, env_kind, nb, obs_dim, aux_dim, metrics_dim;
  int32_t episode_length, auto_reset, substeps, track_metrics, n_walls, pad0;
  float h, dt, gdt /* gravity_z*h */, vel_damp, ang_damp, baumgarte, friction, elasticity;
  float m_torso, m_leg, inv_m_torso, inv_m_leg, r_torso, r_leg;
  float k_joint, sd_joint, ad_joint, ls_joint, act_strength;
  float default_angle[8];
  // per leg l: joint 2l = hip (Torso->Aux), joint 2l+1 = ankle (Aux->lower); offsets have z = 0
  float hip_op[4][2], hip_oc[4][2], ank_op[4][2], ank_oc[4][2];
  float ank_ax[4][2];                 // ankle axis (cos phi, sin phi, 0); ankle ref = e_z
  float foot_e[4][2];                 // lower-leg capsule end (-1) in body frame; the other end is -foot_e
  float aux_e[4][2];                  // Aux capsule ends are +-aux_e
  float hip_lo[4], hip_hi[4], ank_lo[4], ank_hi[4];
  float wall_lo[kMaxWalls][3], wall_hi[kMaxWalls][3];
  // task
  float dying_cost, visible_radius;
  float hh_xy[2][2], priest_xy[2];
  float init_lo[2], init_hi[2];
  float tag_radius, target_step, min_spawn, cage_xy[2];
  int32_t n_apples, n_bombs, n_bins, n_grid;
  float catch_range, sensor_range, half_span, bin_res, spacing, waiting[3];
  float gather_cage[2];
};

}  // namespace pobrax
