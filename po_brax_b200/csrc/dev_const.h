// Device-side constant block derived on the host from PobraxParams (api.cu: build_dev_const).
// Passed to every kernel by value as a __grid_constant__ parameter (constant bank; uniform loads).
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace pobrax {

constexpr int kMaxWalls = 8;
constexpr int kQpPlanes = 32;   // float4 planes per env in the packed state
constexpr int kNumAcc = 8;
constexpr int kMaxObjects = 16; // Gather: n_apples + n_bombs
constexpr int kMaxBins = 32;    // Gather: 2 * n_bins readings

// ---- packed state: float4 qp[kQpPlanes][N] ------------------------------------------------------
// planes 0..3   torso:  (px py pz qw) (qx qy qz vx) (vy vz wx wy) (wz - - -)
// planes 4+7l.. leg l:  A = Aux l (body 1+2l), B = lower leg l (body 2+2l)
//   +0 (A.px A.py A.pz A.qw) +1 (A.qx A.qy A.qz A.vx) +2 (A.vy A.vz A.wx A.wy) +3 (A.wz B.px B.py B.pz)
//   +4 (B.qw B.qx B.qy B.qz) +5 (B.vx B.vy B.vz B.wx) +6 (B.wy B.wz - -)
// ---- aux rows: float aux[aux_dim][N] (frozen bodies that env code moves) ------------------------
//   Ant        : none
//   HeavenHell : 0 ground_x, 1 ground_y, 2 heaven side (0: Target at heaven_hell[0], 1: at heaven_hell[1])
//   Tag        : 0 ground_x, 1 ground_y, 2 tgt_x, 3 tgt_y, 4 tgt_z
//   Gather     : 3k + {0,1,2} = object k xyz (k = 0..n_apples+n_bombs-1: apples then bombs)
// ---- metrics rows: float metrics[metrics_dim][N] ------------------------------------------------
//   Ant        : 0 reward_ctrl_cost, 1 reward_contact_cost, 2 reward_forward, 3 reward_survive
//   HeavenHell : 0 hits      Tag: 0 hits      Gather: 0 apples, 1 bombs
// ---- acc: double[kNumAcc] (track_metrics) --------------------------------------------------------
//   0 finished episodes, 1 sum of episode returns, 2 sum of episode lengths, 3 truncations,
//   4 hits (HeavenHell: any terminal reward; Tag: tags) / apples caught (Gather),
//   5 heavens reached (HeavenHell) / bombs caught (Gather), 6 hells reached (HeavenHell),
//   7 env-steps on which the ant was "dead" (torso z outside [0.2, 1.0])

struct DevConst {
  int32_t n_envs, env_kind, nb, obs_dim, aux_dim, metrics_dim;   // obs_dim: full row (staging stride)
  int32_t obs_lo, obs_out;                   // emitted columns [obs_lo, obs_lo + obs_out); obs_out == obs_dim: all
  int32_t episode_length, auto_reset, substeps, track_metrics, n_walls, has_rng;
  int32_t prefetch_ctas;                     // step kernel: L2 prefetch distance in CTAs (0: this CTA's own, harmless)
  int32_t small_batch_envs;                  // batches up to this size run step_kernel_small (kernels.cu StepCfg)
  // Tag, batches above small_batch_envs: the opponent's move (ant_tag.py:131-132: split + randint) is drawn for the
  // whole batch by tag_rng_kernel, one THREAD per env (5 threefry blocks per env), ahead of the step kernel, which
  // reads it here -- inside the step kernel the 4 lanes of an env cost 12 block executions per env.
  uint8_t* tag_choice;                       // handle-owned device scratch [n_envs], null for the other env families
  float h, dt, gravity_z, vel_damp, ang_damp, baumgarte, friction, elasticity;
  float m_torso, m_leg, inv_m_torso, inv_m_leg, r_torso, r_leg;
  float k_joint, sd_joint, ad_joint, ls_joint, act_strength;
  float h_k, h_sd, h_ad, h_ls, h_act, h_g;   // the same scaled by h (impulses per substep); h_g = h*gravity_z
  // Leg geometry in factored form (validated at create): every joint offset / capsule end of leg l is a
  // uniform scalar times the leg's direction u[l] = (ux, uy, 0) in the body frame.
  float leg_u[4][2];
  float s_hip_p, s_hip_c, s_ank_p, s_ank_c;  // hip parent/child offset, ankle parent/child offset scales
  float s_foot, s_aux;                       // lower-leg capsule end (-1) = s_foot*u; Aux ends = +-s_aux*u
  float seg_aux, seg_foot;                   // capsule half segment lengths (length/2 - r)
  float ank_ax[4][2];                        // ankle axis (cos phi, sin phi, 0); ankle ref = e_z
  float hip_lo, hip_hi;                      // hip: axis e_z, ref -e_x, limits uniform over legs
  float ank_lo[4], ank_hi[4];
  float hip_default, ank_default[4];         // default_angle(): limit midpoints
  // walls: axis-aligned boxes in world coordinates + a per-cell candidate mask for exact culling
  const float4* walls;                       // device float4[n_walls][2]: (lo.xyz, -), (hi.xyz, -)
  float4 wall_box[kMaxWalls][2];             // the same boxes in the constant bank (inline fast path)
  unsigned long long wall_tex;               // cudaTextureObject_t: [3 body types][sdf_ny][sdf_nx] candidate-wall bit
                                             // mask of each xy cell as a layered 2D texture (point sampling, clamped,
                                             // unnormalised coordinates)
  float sdf_x0, sdf_y0, sdf_inv_cell, sdf_bx, sdf_by;   // sdf_bx = -x0 * inv_cell (cell = fma(p, inv_cell, b))
  int32_t sdf_nx, sdf_ny;
  unsigned long long tip_tex;                // cudaTextureObject_t: [tip_ny][tip_nx] candidate-wall bit mask of a lower-leg
                                             // capsule END at that xy cell (finer cells; ant_physics.cuh tip_mask_at)
  float tip_inv_cell, tip_bx, tip_by;
  int32_t tip_nx, tip_ny;
  // task
  float dying_cost, visible_radius;
  float hh_xy[2][2], priest_xy[2], hh_z, priest_z;
  float init_lo[2], init_hi[2];
  float tag_radius, target_step, min_spawn, cage_xy[2];
  int32_t n_apples, n_bombs, n_bins, n_grid;
  float catch_range, sensor_range, half_span, bin_res, spacing, waiting[3];
  float arena_z;                             // Arena body z (half height)
};

}  // namespace pobrax
