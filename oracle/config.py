"""System configs for the oracle: the reference's own Ant Config (golden fixture) + the
task-specific extensions, restated from the reference's config builders.

TEST INFRASTRUCTURE ONLY (see oracle/threefry.py header).

Base Ant: tests/golden/ant_tag_config.json is the brax Config the reference serialised into
/root/reference/notebooks/ant_tag.ipynb:449 (= brax.envs.ant._SYSTEM_CONFIG of the pinned era +
that notebook revision's Target/Arena, which we strip). The product package carries its own
hand-written constant table (po_brax_b200/ant_config.py); parity tests cross-check the two.

Extensions follow:
  * /root/reference/po_brax/envs/utils.py:6-28    add_box_wall_to_body
  * /root/reference/po_brax/envs/utils.py:60-83   draw_arena
  * /root/reference/po_brax/envs/utils.py:87-119  draw_t_maze
  * /root/reference/po_brax/envs/ant_heavenhell.py:13-39, ant_tag.py:13-25, ant_gather.py:17-39
"""
import copy
import json
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden',
                       'ant_tag_config.json')


def _vec(d, default=0.0):
    d = d or {}
    return [float(d.get(k, default)) for k in 'xyz']


def base_ant_config() -> dict:
    """brax.envs.ant._SYSTEM_CONFIG as a dict: 9 ant bodies + Ground, 8 joints, 8 actuators,
    collide_include = {Torso, Body 4/7/10/13} x Ground."""
    cfg = json.load(open(_GOLDEN))
    keep = [b for b in cfg['bodies'] if b['name'] not in ('Target', 'Arena')]
    assert [b['name'] for b in keep][-1] == 'Ground' and len(keep) == 10
    cfg['bodies'] = keep
    cfg['collideInclude'] = [c for c in cfg['collideInclude'] if c['second'] == 'Ground']
    cfg['defaults'] = []
    return cfg


def _frozen_body(name, colliders):
    return {'name': name, 'mass': 1.0, 'inertia': {'x': 1.0, 'y': 1.0, 'z': 1.0},
            'colliders': colliders, 'frozen': {'all': True}}


def _sphere(radius):
    return {'sphere': {'radius': radius}, 'material': {'friction': 1.0, 'elasticity': 0.0}}


def add_box_wall(body, from_xy, to_xy, half_height=0.5, wall_width=0.25):
    """utils.py:6-28: box centred on the segment midpoint, z-rotation = acos(x_hat . v / |v|) deg,
    halfsize (|v|/2, wall_width, half_height)."""
    from_xy = np.asarray(from_xy, np.float32)
    to_xy = np.asarray(to_xy, np.float32)
    vector = to_xy - from_xy
    length = np.linalg.norm(vector)
    mid = (from_xy + to_xy) / 2
    z_rot = np.arccos(np.dot(np.array([1., 0.], np.float32), vector) / length) * 180 / np.pi
    body['colliders'].append({
        'position': {'x': float(mid[0]), 'y': float(mid[1]), 'z': 0.0},
        'rotation': {'x': 0.0, 'y': 0.0, 'z': float(z_rot)},
        'box': {'halfsize': {'x': float(length / 2), 'y': float(wall_width), 'z': float(half_height)}},
        'material': {'friction': 1.0, 'elasticity': 0.0}})


def draw_arena(cfg, cage_x, cage_y, half_height=0.5, name='Arena'):
    """utils.py:60-83 with use_boxes=True (the default every env takes)."""
    x, y = float(cage_x), float(cage_y)
    arena = _frozen_body(name, [])
    cfg['bodies'].append(arena)
    cfg['defaults'].append({'qps': [{'name': name, 'pos': {'x': 0.0, 'y': 0.0, 'z': half_height}}]})
    r = half_height / 2
    pts = np.array([[x + r, y + r], [x + r, -y - r], [-x - r, -y - r], [-x - r, y + r]], np.float32)
    for i in range(4):
        add_box_wall(arena, pts[i], pts[(i + 1) % 4], half_height, r)


def draw_t_maze(cfg, t_x, t_y, hallway_width=2., half_height=0.5, name='Arena'):
    """utils.py:87-119 with use_boxes=True."""
    r = half_height
    w = hallway_width
    arena = _frozen_body(name, [])
    cfg['bodies'].append(arena)
    cfg['defaults'].append({'qps': [{'name': name, 'pos': {'x': 0.0, 'y': 0.0, 'z': half_height}}]})
    pts = np.array([
        [-t_x - r, t_y + r], [t_x + r, t_y + r], [t_x + r, t_y - w - r], [w + r, t_y - w - r],
        [w + r, -r], [-w - r, -r], [-w - r, t_y - w - r], [-t_x - r, t_y - w - r]], np.float32)
    for i in range(len(pts)):
        add_box_wall(arena, pts[i], pts[(i + 1) % len(pts)], half_height, r)


def _collide_with_arena(cfg, ant_body_names):
    for b in ant_body_names:
        cfg['collideInclude'].append({'first': b, 'second': 'Arena'})


def ant_config():
    return base_ant_config()


def heavenhell_config(hhp=((-5.25, 7.), (5.25, 7.), (0., 7.)), hallway_width=2.):
    """ant_heavenhell.py:13-39."""
    cfg = base_ant_config()
    ant_names = [b['name'] for b in cfg['bodies'] if b['name'] != 'Ground']
    cfg['bodies'].append(_frozen_body('Priest', [_sphere(0.5)]))
    cfg['defaults'].append({'qps': [{'name': 'Priest',
                                     'pos': {'x': float(hhp[-1][0]), 'y': float(hhp[-1][1]), 'z': 1.0}}]})
    cfg['bodies'].append(_frozen_body('Target', [_sphere(0.5)]))
    cfg['bodies'].append(_frozen_body('Hell', [_sphere(0.5)]))
    xs = [p[0] for p in hhp]
    ys = [p[1] for p in hhp]
    draw_t_maze(cfg, t_x=max(xs) + hallway_width / 2, t_y=max(ys) + hallway_width / 2,
                hallway_width=hallway_width)
    _collide_with_arena(cfg, ant_names)
    return cfg


def tag_config(cage_max_xy=(4.5, 4.5), offset=1.):
    """ant_tag.py:13-25."""
    cfg = base_ant_config()
    ant_names = [b['name'] for b in cfg['bodies'] if b['name'] != 'Ground']
    cfg['bodies'].append(_frozen_body('Target', [_sphere(0.5)]))
    draw_arena(cfg, cage_max_xy[0] + offset, cage_max_xy[1] + offset, 0.5)
    _collide_with_arena(cfg, ant_names)
    return cfg


def gather_config(cage_max_xy=(6., 6.), offset=1., n_apples=8, n_bombs=8):
    """ant_gather.py:17-39."""
    cfg = base_ant_config()
    ant_names = [b['name'] for b in cfg['bodies'] if b['name'] != 'Ground']
    draw_arena(cfg, cage_max_xy[0] + offset, cage_max_xy[1] + offset, 0.5)
    _collide_with_arena(cfg, ant_names)
    for i in range(n_apples):
        cfg['bodies'].append(_frozen_body(f'Target_{i + 1}', [_sphere(0.25)]))
    for i in range(n_bombs):
        cfg['bodies'].append(_frozen_body(f'Bomb_{i + 1}', [_sphere(0.25)]))
    return cfg


def scale_action_repeat(cfg, action_repeat):
    """wrappers.py:16-24 ActionRepeatWrapper: dt *= k, substeps *= k."""
    cfg = copy.deepcopy(cfg)
    cfg['dt'] = cfg['dt'] * action_repeat
    cfg['substeps'] = int(cfg['substeps'] * action_repeat)
    return cfg
