"""ctypes binding of libpobrax.so (include/pobrax.h). There is no CPU or PyTorch fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('POBRAX_LIB') or os.path.join(HERE, 'libpobrax.so')   # POBRAX_LIB: tuning builds

ABI_VERSION = 1
ANT, ANT_HEAVENHELL, ANT_GATHER, ANT_TAG = 0, 1, 2, 3
AUTORESET_OFF, AUTORESET_CACHED = 0, 1
MAX_WALLS, QP_PLANES, NUM_ACC = 8, 32, 8

_f, _i = C.c_float, C.c_int32


class PobraxParams(C.Structure):
    _fields_ = [
        ('env_kind', _i), ('num_envs', _i), ('episode_length', _i), ('auto_reset', _i), ('action_repeat', _i),
        ('track_metrics', _i), ('obs_col_lo', _i), ('obs_col_hi', _i),
        ('dt', _f), ('substeps', _i), ('gravity_z', _f), ('velocity_damping', _f), ('angular_damping', _f),
        ('baumgarte_erp', _f), ('friction', _f), ('elasticity', _f),
        ('torso_mass', _f), ('leg_mass', _f), ('torso_radius', _f), ('leg_radius', _f),
        ('aux_length', _f), ('foot_length', _f),
        ('collider_euler', _f * 3 * 4),
        ('hip_off_p', _f * 3 * 4), ('hip_off_c', _f * 3 * 4), ('ank_off_p', _f * 3 * 4), ('ank_off_c', _f * 3 * 4),
        ('hip_euler', _f * 3 * 4), ('ank_euler', _f * 3 * 4),
        ('hip_limit', _f * 2 * 4), ('ank_limit', _f * 2 * 4),
        ('joint_stiffness', _f), ('joint_spring_damping', _f), ('joint_angular_damping', _f),
        ('joint_limit_strength', _f), ('actuator_strength', _f),
        ('num_walls', _i), ('wall_lo', _f * 3 * MAX_WALLS), ('wall_hi', _f * 3 * MAX_WALLS), ('arena_z', _f),
        ('dying_cost', _f), ('visible_radius', _f),
        ('heaven_hell_xy', _f * 2 * 2), ('priest_xy', _f * 2),
        ('init_lo', _f * 2), ('init_hi', _f * 2),
        ('tag_radius', _f), ('target_step', _f), ('min_spawn_distance', _f), ('cage_xy', _f * 2),
        ('n_apples', _i), ('n_bombs', _i), ('n_bins', _i),
        ('catch_range', _f), ('sensor_range', _f), ('sensor_span', _f), ('robot_object_spacing', _f),
        ('gather_cage_xy', _f * 2),
    ]


class PobraxState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'qp', 'aux', 'obs', 'reward', 'done', 'steps', 'truncation', 'rng', 'metrics', 'first_qp', 'first_aux',
        'first_obs', 'ep_return', 'acc')]


class PobraxLayout(C.Structure):
    _fields_ = [(n, _i) for n in ('num_bodies', 'obs_dim', 'aux_dim', 'metrics_dim', 'action_dim', 'qp_planes')]


EXPORTS = {
    'pobrax_abi_version': (C.c_int, []),
    'pobrax_last_error': (C.c_char_p, []),
    'pobrax_struct_sizes': (C.c_int, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    'pobrax_default_params': (C.c_int, [C.c_int, C.POINTER(PobraxParams)]),
    'pobrax_draw_arena': (C.c_int, [C.POINTER(PobraxParams), _f, _f, _f]),
    'pobrax_draw_t_maze': (C.c_int, [C.POINTER(PobraxParams), _f, _f, _f, _f]),
    'pobrax_layout': (C.c_int, [C.POINTER(PobraxParams), C.POINTER(PobraxLayout)]),
    'pobrax_create': (C.c_int, [C.POINTER(PobraxParams), C.c_int, C.POINTER(C.c_void_p)]),
    'pobrax_destroy': (C.c_int, [C.c_void_p]),
    'pobrax_reset': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PobraxState), C.c_void_p]),
    'pobrax_step': (C.c_int, [C.c_void_p, C.POINTER(PobraxState), C.c_void_p, C.c_void_p]),
    'pobrax_reset_where_done': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PobraxState), C.c_void_p]),
    'pobrax_reset_where_done_chain': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PobraxState), C.c_void_p]),
    'pobrax_unpack_qp': (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    'pobrax_pack_qp': (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    'pobrax_split_pairs': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pobrax_eval_update': (C.c_int, [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_void_p]),
    'pobrax_fp32_probe': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]),
    'pobrax_split_keys': (C.c_int, [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Loads libpobrax.so (built by `python -m po_brax_b200.build`). Raises if absent: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} not found: build it with `python -m po_brax_b200.build` '
                           '(po_brax_b200 has no CPU / PyTorch fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.pobrax_abi_version() != ABI_VERSION:
        raise RuntimeError('libpobrax.so ABI version mismatch; rebuild with `python -m po_brax_b200.build --force`')
    a, b, c = _i(), _i(), _i()
    lib.pobrax_struct_sizes(C.byref(a), C.byref(b), C.byref(c))
    if (a.value, b.value, c.value) != (C.sizeof(PobraxParams), C.sizeof(PobraxState), C.sizeof(PobraxLayout)):
        raise RuntimeError('ctypes struct mirrors do not match libpobrax.so (include/pobrax.h changed?)')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pobrax_last_error()
        raise RuntimeError(f'{what} failed ({rc}): {msg.decode() if msg else "unknown error"}')
