"""How far is the documented box-wall substitute (oracle/brax_v1.py `_wall_contacts`, what the CUDA path
implements) from brax's capsule-vs-triangulated-box collider as recalled (`_wall_contacts_mesh`)?

TEST INFRASTRUCTURE / ANALYSIS ONLY (see oracle/threefry.py header). Neither collider is pinned by a reference
artefact (the reference has no wall-contact data); this script only measures one against the other.

Teacher-forced: a rollout is driven by the substitute (C twin); at every step BOTH colliders take one env step from
the same state, and the per-env difference of the resulting velocities is classified:
  * no wall contact in either                         -> identical by construction
  * contact, |d vel| <= 3e-4                          -> the two colliders agree (one triangle active, same point)
  * differ, a capsule point inside a box (dist = 0)   -> the substitute gives no impulse; the mesh pushes along
                                                         (segment point - triangle point), i.e. further INTO the wall
  * differ, >= 2 triangles active on a body           -> the mesh averages contacts of neighbouring triangles
  * differ, other
    python -m oracle.wall_deviation [env] [n_envs] [steps] [action_period]
"""
import os
import sys

import numpy as np

from . import brax_v1 as bx
from . import cstep
from . import envs as oenvs
from . import threefry as tf


TOL = 3e-4   # the velocity parity bar of tests/_parity.py (the mesh path's own 1e-6 epsilons move results by ~4e-5)


def run(kind='ant_heavenhell', n=512, T=120, period=4, seed=0):
    drive = oenvs.ENVS[kind]()
    cstep.attach(drive.sys, threads=os.cpu_count() or 1)
    box = oenvs.ENVS[kind]().sys
    mesh = oenvs.ENVS[kind](walls='mesh').sys
    s = drive.reset(tf.split(tf.prng_key(seed), n + 1)[1:])
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (period, n, 8)).astype(np.float32)
    tot = dict(env_steps=0, wall_contact=0, agree=0, inside=0, multi=0, other=0)
    dv_all = []
    for t in range(T):
        a = acts[t % period]
        qb, ib = box.step(s.qp, a)
        mesh.mesh_active_max = None
        qm, im = mesh.step(s.qp, a)
        d = np.maximum(np.abs(qb.vel - qm.vel), np.abs(qb.ang - qm.ang))[:, :9].reshape(n, -1).max(1)
        ant = slice(0, 9)
        # wall impulses of the step: Info.contact of the Aux bodies is wall-only; for the other bodies compare with a
        # wall-free evaluation is overkill -- use "either collider's result differs from the no-wall step"
        wall_b = np.abs(ib.contact_vel[:, ant] - im.contact_vel[:, ant]).reshape(n, -1).max(1) > 0
        touched = _touching(box, s.qp) | wall_b
        inside = _inside(box, s.qp) | _inside(box, qb)
        mam = mesh.mesh_active_max
        multi = (mam.max(1) >= 2) if mam is not None else np.zeros(n, bool)
        differ = d > TOL
        tot['env_steps'] += n
        tot['wall_contact'] += int(touched.sum())
        tot['agree'] += int((touched & ~differ).sum())
        tot['inside'] += int((differ & inside).sum())
        tot['multi'] += int((differ & ~inside & multi).sum())
        tot['other'] += int((differ & ~inside & ~multi).sum())
        dv_all.append(d[differ])
        s = drive.step(s, a)
    dv = np.concatenate(dv_all) if dv_all else np.zeros(0)
    q = np.quantile(dv, [0.5, 0.9, 0.99]) if len(dv) else [0, 0, 0]
    return tot, q


def _seg_box(sysm, qp):
    nbx = len(sysm.boxes)
    b = np.repeat(sysm.cap_body, nbx)
    ca, cb = np.repeat(sysm.cap_a, nbx, axis=0), np.repeat(sysm.cap_b, nbx, axis=0)
    rad = np.repeat(sysm.cap_rad, nbx)
    box = np.tile(sysm.boxes, (len(sysm.cap_body), 1))
    pos, rot = qp.pos[:, b], qp.rot[:, b]
    apos = qp.pos[:, sysm.arena][:, None, :]
    a_w = pos + bx.rotate(np.broadcast_to(ca, pos.shape), rot)
    b_w = pos + bx.rotate(np.broadcast_to(cb, pos.shape), rot)
    sp, bp = sysm._closest_segment_box(a_w, b_w, apos + box[:, :3], apos + box[:, 3:])
    return np.sqrt(((sp - bp) ** 2).sum(-1)), rad


def _touching(sysm, qp):
    d, rad = _seg_box(sysm, qp)
    return ((d < rad) & (d > 0)).any(1)


def _inside(sysm, qp):
    d, _ = _seg_box(sysm, qp)
    return (d == 0).any(1)


if __name__ == '__main__':
    kind = sys.argv[1] if len(sys.argv) > 1 else 'ant_heavenhell'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 120
    period = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    tot, q = run(kind, n, T, period)
    es = tot['env_steps']
    print(f'{kind}: {n} envs x {T} steps, action period {period}')
    print(f"  env-steps with a capsule touching a wall at the start of the step: {tot['wall_contact'] / es:.4f}")
    differ = tot['inside'] + tot['multi'] + tot['other']
    print(f"  env-steps on which the two colliders differ by > {TOL:g}: {differ / es:.4f} "
          f"(inside a box {tot['inside'] / es:.4f}, >= 2 triangles active {tot['multi'] / es:.4f}, other {tot['other'] / es:.4f})")
    print(f'  |d vel| where they differ: median {q[0]:.3g}, 90 % {q[1]:.3g}, 99 % {q[2]:.3g}')
