"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/pobrax.h declares, its
parameter tables match the reference's config builders, and without a GPU it fails loudly (no CPU path).
No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import brax_v1 as bx
from oracle import config as ocfg
from po_brax_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from po_brax_b200 import build
    build.build()  # no-op when the in-tree .so is fresh
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'pobrax.h')).read()
    declared = set(re.findall(r'\b(pobrax_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/pobrax.h but not exported'
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert lib.pobrax_abi_version() == _lib.ABI_VERSION


def test_struct_mirrors_match(lib):
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.pobrax_struct_sizes(C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (C.sizeof(_lib.PobraxParams), C.sizeof(_lib.PobraxState),
                                           C.sizeof(_lib.PobraxLayout))


@pytest.mark.parametrize('kind,nb,obs,aux,met', [(_lib.ANT, 10, 87, 0, 4), (_lib.ANT_HEAVENHELL, 14, 114, 3, 1),
                                                 (_lib.ANT_GATHER, 27, 211, 48, 2), (_lib.ANT_TAG, 12, 103, 5, 1)])
def test_layouts(lib, kind, nb, obs, aux, met):
    p, L = _lib.PobraxParams(), _lib.PobraxLayout()
    assert lib.pobrax_default_params(kind, C.byref(p)) == 0
    assert lib.pobrax_layout(C.byref(p), C.byref(L)) == 0
    assert (L.num_bodies, L.obs_dim, L.aux_dim, L.metrics_dim, L.action_dim, L.qp_planes) == (nb, obs, aux, met, 8, 32)
    assert (p.dt, p.substeps, p.episode_length) == (np.float32(0.05), 10, 1000)


@pytest.mark.parametrize('kind,cfg_fn', [(_lib.ANT_HEAVENHELL, ocfg.heavenhell_config), (_lib.ANT_TAG, ocfg.tag_config),
                                         (_lib.ANT_GATHER, ocfg.gather_config)])
def test_walls_match_reference_config_builders(lib, kind, cfg_fn):
    """pobrax_draw_t_maze / pobrax_draw_arena vs the restated utils.py:60-119 builders (world-frame boxes)."""
    p = _lib.PobraxParams()
    assert lib.pobrax_default_params(kind, C.byref(p)) == 0
    sys_ = bx.System(cfg_fn())
    assert p.num_walls == len(sys_.boxes)
    got = np.array([list(p.wall_lo[w]) + list(p.wall_hi[w]) for w in range(p.num_walls)], np.float32)
    want = sys_.boxes.copy()
    want[:, 2] += 0.5  # Arena body sits at z = half height (default_qp lifts it onto the ground)
    want[:, 5] += 0.5
    assert np.abs(got - want).max() < 1e-6


def test_ant_constants_match_reference_config(lib):
    """The library's Ant table vs the reference's own Config (tests/golden/ant_tag_config.json)."""
    p = _lib.PobraxParams()
    lib.pobrax_default_params(_lib.ANT, C.byref(p))
    s = bx.System(ocfg.ant_config())
    assert np.allclose([p.torso_mass, p.leg_mass], [s.mass[0], s.mass[1]])
    for l in range(4):
        hip, ank = 2 * l, 2 * l + 1
        assert np.allclose(list(p.hip_off_p[l]), s.j_off_p[hip]) and np.allclose(list(p.hip_off_c[l]), s.j_off_c[hip])
        assert np.allclose(list(p.ank_off_p[l]), s.j_off_p[ank]) and np.allclose(list(p.ank_off_c[l]), s.j_off_c[ank])
        assert np.allclose(np.deg2rad(list(p.hip_limit[l])), s.j_limit[hip], atol=1e-6)
        assert np.allclose(np.deg2rad(list(p.ank_limit[l])), s.j_limit[ank], atol=1e-6)
    assert (p.joint_stiffness, p.joint_spring_damping, p.joint_angular_damping) == (s.j_stiff[0], s.j_sdamp[0], s.j_adamp[0])
    assert p.actuator_strength == s.a_strength[0]
    assert np.isclose(p.foot_length / 2 - p.leg_radius, np.linalg.norm(s.cp_end[1]), atol=1e-6)


def test_bad_arguments_are_reported(lib):
    p = _lib.PobraxParams()
    assert lib.pobrax_default_params(7, C.byref(p)) != 0
    assert b'env_kind' in lib.pobrax_last_error()
    lib.pobrax_default_params(_lib.ANT_GATHER, C.byref(p))
    p.n_apples = 20
    L = _lib.PobraxLayout()
    assert lib.pobrax_layout(C.byref(p), C.byref(L)) != 0


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    p = _lib.PobraxParams()
    lib.pobrax_default_params(_lib.ANT, C.byref(p))
    h = C.c_void_p()
    assert lib.pobrax_create(C.byref(p), 0, C.byref(h)) != 0 and not h.value
    assert b'no CPU path' in lib.pobrax_last_error() or b'CUDA' in lib.pobrax_last_error()
    from po_brax_b200 import envs
    with pytest.raises(RuntimeError):
        envs.create('ant', batch_size=4)


def test_observation_column_range(lib):
    """obs_col_lo/hi (the fused observability masks) change only the reported observation width."""
    p, L = _lib.PobraxParams(), _lib.PobraxLayout()
    lib.pobrax_default_params(_lib.ANT, C.byref(p))
    for lo, hi in ((0, 13), (13, 27), (27, 87)):
        p.obs_col_lo, p.obs_col_hi = lo, hi
        assert lib.pobrax_layout(C.byref(p), C.byref(L)) == 0 and L.obs_dim == hi - lo
    p.obs_col_lo, p.obs_col_hi = 50, 100
    assert lib.pobrax_layout(C.byref(p), C.byref(L)) != 0
