/* pobrax.h -- C ABI of libpobrax.so: the B200 (sm_100a) fused stepper for po-brax's Ant POMDP envs.
 *
 * Plain C types only (no torch / C++ types): device buffers are passed as raw device pointers that
 * the caller owns (the Python host passes torch tensors' data_ptr()), streams as void* (cudaStream_t).
 * Every entry point returns 0 on success, non-zero on failure; pobrax_last_error() has the message.
 * Calls enqueue on the given stream and never synchronise.
 *
 * What each entry point replaces in the reference (/root/reference/po_brax):
 *   pobrax_default_params  envs/__init__.py:29-33 (_envs registry) + the env constructors' defaults
 *                          envs/ant_heavenhell.py:51-73, envs/ant_gather.py:59-91, envs/ant_tag.py:38-61,
 *                          envs/utils.py:60-119 (arena / T-maze walls) and brax.envs.ant._SYSTEM_CONFIG
 *   pobrax_create/destroy  envs/__init__.py:50-72 create(): env ctor -> ActionRepeat -> Episode -> Vmap -> AutoReset
 *   pobrax_reset           env.reset(rng) under VmapWrapper: envs/ant_heavenhell.py:75-103,
 *                          envs/ant_gather.py:93-123, envs/ant_tag.py:63-105, brax.envs.ant.Ant.reset
 *   pobrax_step            env.step(state, action): envs/ant_heavenhell.py:106-158, envs/ant_gather.py:125-213,
 *                          envs/ant_tag.py:107-181, brax.envs.ant.Ant.step, i.e. brax.System.step (10 substeps)
 *                          + task logic + brax EpisodeWrapper + brax AutoResetWrapper (envs/wrappers.py:27)
 *   pobrax_reset_where_done  gym-level autoreset, envs/wrappers.py:245-262 (fresh keys, select by done)
 *   pobrax_reset_where_done_chain  the same, gym key chain (wrappers.py:160-163) on the device: no host sync
 *   pobrax_unpack_qp/pack_qp  State.qp pytree (brax.QP pos/rot/vel/ang [N,nb,*]) <-> packed SoA state
 *   pobrax_split_keys      jax.random.split(key, n) as used by VmapGymWrapper._reset, envs/wrappers.py:160-163
 */
#ifndef POBRAX_H_
#define POBRAX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POBRAX_ABI_VERSION 1

enum PobraxEnvKind { POBRAX_ANT = 0, POBRAX_ANT_HEAVENHELL = 1, POBRAX_ANT_GATHER = 2, POBRAX_ANT_TAG = 3 };

/* autoreset modes (PobraxParams.auto_reset) */
enum PobraxAutoReset {
  POBRAX_AUTORESET_OFF = 0,    /* create(auto_reset=False): the gym layer resets (pobrax_reset_where_done) */
  POBRAX_AUTORESET_CACHED = 1  /* brax AutoResetWrapper: qp/obs <- first_qp/first_obs where done */
};

#define POBRAX_MAX_WALLS 8
#define POBRAX_QP_PLANES 32  /* float4 planes per env in the packed state (512 B) */
#define POBRAX_NUM_ACC 8     /* episode-metric accumulators (double) */

/* Constants of one env family. pobrax_default_params() fills it; the host may edit before create. */
typedef struct PobraxParams {
  int32_t env_kind;
  int32_t num_envs;
  int32_t episode_length;   /* brax EpisodeWrapper; <= 0 disables */
  int32_t auto_reset;       /* PobraxAutoReset */
  int32_t action_repeat;    /* wrappers.py:16-24: dt *= k, substeps *= k (applied at create) */
  int32_t track_metrics;    /* 1: maintain ep_return + acc[] (device-side episode statistics) */
  int32_t obs_col_lo, obs_col_hi; /* emit only observation columns [lo, hi) (0, 0 = all): the contiguous index sets of
                                     standard_observability_masks.py ('ant': 0:13, 13:27, 27:87) fused into the store */
  /* ---- brax system (Ant) ---- */
  float dt;                 /* 0.05 */
  int32_t substeps;         /* 10 */
  float gravity_z;          /* -9.8 */
  float velocity_damping;   /* 0 */
  float angular_damping;    /* -0.05 (global) */
  float baumgarte_erp;      /* 0.1 */
  float friction;           /* 1 */
  float elasticity;         /* 0 */
  float torso_mass, leg_mass;               /* 10, 1 */
  float torso_radius, leg_radius;           /* 0.25, 0.08 */
  float aux_length, foot_length;            /* capsule lengths 0.44284272, 0.7256854 */
  float collider_euler[4][3];               /* capsule rotation (deg) of leg l's Aux and lower body */
  float hip_off_p[4][3], hip_off_c[4][3];   /* joint 2l   (Torso -> Aux l): parent/child offsets */
  float ank_off_p[4][3], ank_off_c[4][3];   /* joint 2l+1 (Aux l -> lower l) */
  float hip_euler[4][3], ank_euler[4][3];   /* joint frame rotations (deg) */
  float hip_limit[4][2], ank_limit[4][2];   /* angle limits (deg) */
  float joint_stiffness, joint_spring_damping, joint_angular_damping, joint_limit_strength; /* 18000 80 20 18000 */
  float actuator_strength;                  /* 350 */
  /* ---- arena walls: axis-aligned boxes in world coordinates (Arena body at z = half height) ---- */
  int32_t num_walls;
  float wall_lo[POBRAX_MAX_WALLS][3], wall_hi[POBRAX_MAX_WALLS][3];
  float arena_z;            /* z of the frozen Arena body (= wall half height, 0.5) */
  /* ---- task ---- */
  float dying_cost;
  float visible_radius;     /* HeavenHell 2.0, Tag 3.0 */
  float heaven_hell_xy[2][2], priest_xy[2];   /* HeavenHell */
  float init_lo[2], init_hi[2];             /* ant spawn box (HeavenHell: [-.5,.5]x[.5,1.5]; Tag: +-cage) */
  float tag_radius, target_step, min_spawn_distance, cage_xy[2];   /* Tag */
  int32_t n_apples, n_bombs, n_bins;        /* Gather: 8, 8, 10 */
  float catch_range, sensor_range, sensor_span, robot_object_spacing, gather_cage_xy[2];
} PobraxParams;

/* Device buffers of one batched State. All pointers are device memory owned by the caller; arrays are
 * contiguous. [N] = num_envs. Sizes/dtypes are reported by pobrax_layout(). Pointers that a given
 * env/mode does not use may be NULL. */
typedef struct PobraxState {
  float* qp;          /* float4[POBRAX_QP_PLANES][N]: 9 ant bodies, packed SoA */
  float* aux;         /* float[aux_dim][N]: per-env frozen-body data (ground xy, target, objects...) */
  float* obs;         /* float[N][obs_dim] */
  float* reward;      /* float[N] */
  float* done;        /* float[N] 0/1 */
  float* steps;       /* float[N]   info['steps'] */
  float* truncation;  /* float[N]   info['truncation'] */
  uint32_t* rng;      /* uint32[N][2] info['rng'] (NULL for plain Ant) */
  float* metrics;     /* float[metrics_dim][N] per-step State.metrics */
  float* first_qp;    /* like qp   (info['first_qp'],  AUTORESET_CACHED) */
  float* first_aux;   /* like aux */
  float* first_obs;   /* like obs  (info['first_obs']) */
  float* ep_return;   /* float[N] running undiscounted return (track_metrics) */
  double* acc;        /* double[POBRAX_NUM_ACC]: episodes, sum return, sum length, truncations, hits|apples,
                         heavens|bombs, hells, dead steps (see csrc/dev_const.h) */
} PobraxState;

typedef struct PobraxLayout {
  int32_t num_bodies;   /* nb of the brax system: Ant 10, Tag 12, HeavenHell 14, Gather 27 */
  int32_t obs_dim;      /* 87 / 103 / 114 / 211, or obs_col_hi - obs_col_lo when a column range is selected */
  int32_t aux_dim;
  int32_t metrics_dim;
  int32_t action_dim;   /* 8 */
  int32_t qp_planes;    /* POBRAX_QP_PLANES */
} PobraxLayout;

int pobrax_abi_version(void);
/* sizeof(PobraxParams), sizeof(PobraxState), sizeof(PobraxLayout): lets an FFI binding verify its struct mirrors. */
int pobrax_struct_sizes(int32_t* params, int32_t* state, int32_t* layout);
const char* pobrax_last_error(void);

int pobrax_default_params(int env_kind, PobraxParams* out);
/* utils.py:60-83 draw_arena(cage_x, cage_y, half_height) and utils.py:87-119 draw_t_maze(t_x, t_y, hallway_width,
 * half_height), box walls (use_boxes=True): overwrite num_walls / wall_lo / wall_hi of *p. */
int pobrax_draw_arena(PobraxParams* p, float cage_x, float cage_y, float half_height);
int pobrax_draw_t_maze(PobraxParams* p, float t_x, float t_y, float hallway_width, float half_height);
int pobrax_layout(const PobraxParams* p, PobraxLayout* out);

/* Validates *p, builds the per-handle constants (wall boxes, candidate-wall tables behind texture objects, Gather
 * grid) on `device` and sets up the kernels for that device (dynamic shared-memory limits, occupancy): a process may
 * hold handles on several GPUs. Fails with a message (pobrax_last_error) when e.g. the observation staging + Gather
 * grid exceed the device's shared memory per block. Tuning knobs read here from the environment (never needed for
 * correctness): POBRAX_PREFETCH_CTAS (L2 prefetch distance of the step kernel), POBRAX_SMALL_BATCH_ENVS (batches up to
 * this size run the small-batch instantiation of the step kernel -- same results bit for bit; default 2 warps per SM
 * sub-partition = 9 472 envs on a B200; 0 = never). */
int pobrax_create(const PobraxParams* p, int device, void** handle);
int pobrax_destroy(void* handle);

/* keys: device uint32[N][2]. Writes every buffer of st (first_* if present). */
int pobrax_reset(void* handle, const uint32_t* keys, PobraxState* st, void* stream);
/* action: device float[N][8]. In-place update of st. One launch (the fused step kernel); for Ant-Tag batches above the
 * small-batch threshold two: tag_rng_kernel draws the opponent's move of ant_tag.py:131-132 for the whole batch into
 * an N-byte scratch buffer the handle owns (allocated in pobrax_create) and advances st->rng, then the step kernel. */
int pobrax_step(void* handle, PobraxState* st, const float* action, void* stream);
/* gym autoreset: where st->done != 0, replace qp/aux/obs by reset(keys[i]) and zero steps. */
int pobrax_reset_where_done(void* handle, const uint32_t* keys, PobraxState* st, void* stream);
/* The same with the gym key chain of VmapGymWrapper._reset (envs/wrappers.py:160-163, 247-248) kept on the device,
 * so AutoresetVmapGymWrapper.step needs no `done.any()` host round trip: chain = device uint32[4] = {gym key k0, k1,
 * scratch flag (0 between calls), spare}. If some env is done: keys = split(gym key, N + 1), done envs reset from
 * keys[i + 1] and gym key <- keys[0]; if none is done nothing changes (the reference draws no keys either). */
int pobrax_reset_where_done_chain(void* handle, uint32_t* chain, PobraxState* st, void* stream);

/* brax.QP views: pos[N][nb][3], rot[N][nb][4], vel[N][nb][3], ang[N][nb][3] (device). */
int pobrax_unpack_qp(void* handle, const float* qp, const float* aux, float* pos, float* rot, float* vel,
                     float* ang, void* stream);
int pobrax_pack_qp(void* handle, const float* pos, const float* rot, const float* vel, const float* ang,
                   float* qp, float* aux, void* stream);

/* EvalGymWrapper.step (envs/wrappers.py:202-219) in one launch, all pointers device memory of n elements (sums: 4
 * doubles = finished episodes, sum of their returns, discounted returns and lengths -- what get_stats() averages):
 * returns += r; lengths += 1; disc_returns += r * cur_discount; cur_discount *= discount; where done: the episode goes
 * into sums and its four slots restart (0, 0, 0, 1). */
int pobrax_eval_update(const float* reward, const float* done, float* returns, float* disc_returns, long long* lengths,
                       float* cur_discount, double* sums, float discount, int n, void* stream);

/* jax.random.split(key, n): key = host uint32[2]; out = device uint32[n][2] (rows first..first+count). */
int pobrax_split_keys(const uint32_t key[2], int n, int first, int count, uint32_t* out, void* stream);

/* vmapped jax.random.split(key, 2) on device keys uint32[n][2]: out_a = split[0], out_b = split[1]
 * (the `rng, rng1 = jp.random_split(state.info['rng'], 2)` of envs/wrappers.py:102). */
int pobrax_split_pairs(const uint32_t* keys, int n, uint32_t* out_a, uint32_t* out_b, void* stream);

/* Measurement aid (bench.py): one launch of `blocks` x 256 threads x `iters` x 64 dependent-chain FMAs;
 * *flops = FLOPs of the launch (FMA = 2). out: device float[blocks * 256] scratch. */
int pobrax_fp32_probe(float* out, int blocks, int iters, void* stream, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* POBRAX_H_ */
