"""The wall collider, specified pose by pose (CPU).

brax resolves ant-capsule vs Arena-box pairs through a triangulated-box mesh collider; oracle and CUDA use a documented
substitute -- ONE contact per (capsule, axis-aligned box) at the closest box point. Nothing in
the reference pins either. These tests (a) specify the substitute against hand-evaluated formulas in three poses --
a capsule end on a flat face, on a vertical box edge, and in the concave corner of the T junction (two boxes) -- and
(b) hold it against brax's mesh collider as recalled (oracle/brax_v1.py `_wall_contacts_mesh`): identical on a face away
from the triangle diagonals, on an edge and in the corner; different, in a stated direction, on a diagonal and for a
capsule point inside a box."""
import numpy as np
import pytest

from oracle import brax_v1 as bx
from oracle import envs as oenvs

F = np.float32
R = 0.08                      # leg capsule radius
LOWER = {0: 2, 1: 4, 2: 6, 3: 8}   # leg -> lower-leg body


@pytest.fixture(scope='module')
def systems():
    return oenvs.ENVS['ant_heavenhell']().sys, oenvs.ENVS['ant_heavenhell'](walls='mesh').sys


def _pose(S, leg, tip, vel):
    """Default-pose ant translated so that the tip (t = 0 end) of `leg`'s lower leg sits at `tip`, moving rigidly with
    `vel`. Returns (qp, body index, tip offset from the body centre)."""
    qp = S.default_qp(S.default_angle()[None].astype(F), np.zeros((1, 8), F))
    body = LOWER[leg]
    k = list(S.cap_body).index(body)
    a = qp.pos[0, body] + bx.rotate(S.cap_a[k][None], qp.rot[:, body])[0]
    shift = np.asarray(tip, F) - a
    qp.pos[:, :9] += shift
    qp.vel[:, :9] = np.asarray(vel, F)
    return qp, body, a + shift - qp.pos[0, body]


def _one_contact(pos_body, v_body, w_body, point, n, pen, inv_m=1.0):
    """SURVEY App. A.4 for one contact on a unit-inertia body, evaluated in float64 from scratch."""
    rel = point - pos_body
    v = v_body + np.cross(w_body, rel)
    nv = n @ v
    ang = n @ np.cross(np.cross(rel, n), rel)
    J = (20.0 * pen - nv) / (inv_m + ang)
    if not (pen > 0 and nv < 0 and J > 0):
        return np.zeros(3), np.zeros(3)
    dv, dw = inv_m * J * n, np.cross(rel, J * n)
    vd = v - nv * n
    nd = np.linalg.norm(vd)
    if nd > 0.01:
        Jd = min(nd / (inv_m + ang), 1.0 * J)
        jd = -Jd * vd / (1e-6 + nd)
        dv, dw = dv + inv_m * jd, dw + np.cross(rel, jd)
    return dv, dw


def _info(S, qp, body):
    i = S.info(qp)
    return i.contact_vel[0, body].astype(np.float64), i.contact_ang[0, body].astype(np.float64)


def test_capsule_end_on_a_flat_face(systems):
    """Tip 0.06 from the stem's right wall (box x in [2, 3]), away from the face's triangle diagonal: one contact, normal
    -x scaled by d / (1e-6 + d); the mesh collider activates exactly one triangle and agrees."""
    box, mesh = systems
    tip = np.array([2.0 - 0.06, 2.0, 0.30])
    qp, body, e = _pose(box, 0, tip, (1.0, 0.2, 0.0))
    d = 0.06
    n = np.array([-1.0, 0.0, 0.0]) * d / (1e-6 + d)
    want = _one_contact(qp.pos[0, body].astype(np.float64), qp.vel[0, body].astype(np.float64), np.zeros(3),
                        np.array([2.0, 2.0, 0.30]), n, R - d)
    got = _info(box, qp, body)
    assert np.allclose(got[0], want[0], atol=2e-6) and np.allclose(got[1], want[1], atol=2e-6), (got, want)
    assert abs(want[0][0]) > 0.5                                    # a real push back along -x
    gm = _info(mesh, qp, body)
    assert np.allclose(gm[0], got[0], atol=2e-6) and np.allclose(gm[1], got[1], atol=2e-6)
    assert mesh.last_mesh_active[0, body] == 1


def test_capsule_end_on_a_vertical_edge(systems):
    """Tip diagonally off the convex vertical edge (2, 5.5) of the same box: closest box point ON the edge, normal
    along the diagonal. The mesh collider finds the same point on the triangles of BOTH faces meeting there; their
    contacts are identical, so summing and dividing by the active count changes nothing."""
    box, mesh = systems
    tip = np.array([2.0 - 0.04, 5.5 + 0.04, 0.30])
    qp, body, e = _pose(box, 0, tip, (1.0, -0.6, 0.0))
    dvec = np.array([-0.04, 0.04, 0.0])
    d = np.linalg.norm(dvec)
    want = _one_contact(qp.pos[0, body].astype(np.float64), qp.vel[0, body].astype(np.float64), np.zeros(3),
                        np.array([2.0, 5.5, 0.30]), dvec / (1e-6 + d), R - d)
    got = _info(box, qp, body)
    assert np.allclose(got[0], want[0], atol=3e-6) and np.allclose(got[1], want[1], atol=3e-6), (got, want)
    gm = _info(mesh, qp, body)
    assert mesh.last_mesh_active[0, body] >= 2
    assert np.allclose(gm[0], got[0], atol=5e-5) and np.allclose(gm[1], got[1], atol=5e-5), (gm, got)   # the mesh path carries 1e-6 epsilons


def test_capsule_end_in_the_concave_corner_of_the_t_junction(systems):
    """Tip in the notch at (2.5, 5.5): within reach of box 2's face x = 2.5 AND of box 3's end face y = 5.5 -- two
    (capsule, box) pairs, two contacts evaluated on the same state, summed and divided by two (both active)."""
    box, mesh = systems
    tip = np.array([2.5 - 0.05, 5.5 + 0.06, 0.30])
    qp, body, e = _pose(box, 3, tip, (1.0, -1.0, 0.0))
    pb, vb = qp.pos[0, body].astype(np.float64), qp.vel[0, body].astype(np.float64)
    c1 = _one_contact(pb, vb, np.zeros(3), np.array([2.5, 5.56, 0.30]), np.array([-1.0, 0, 0]) * 0.05 / (1e-6 + 0.05), R - 0.05)
    c2 = _one_contact(pb, vb, np.zeros(3), np.array([2.45, 5.5, 0.30]), np.array([0, 1.0, 0]) * 0.06 / (1e-6 + 0.06), R - 0.06)
    assert np.abs(c1[0]).max() > 0 and np.abs(c2[0]).max() > 0
    want = (c1[0] + c2[0]) / 2, (c1[1] + c2[1]) / 2
    got = _info(box, qp, body)
    assert np.allclose(got[0], want[0], atol=3e-6) and np.allclose(got[1], want[1], atol=3e-6), (got, want)
    gm = _info(mesh, qp, body)
    assert mesh.last_mesh_active[0, body] == 2
    assert np.allclose(gm[0], got[0], atol=5e-5) and np.allclose(gm[1], got[1], atol=5e-5), (gm, got)   # the mesh path carries 1e-6 epsilons


def test_on_a_triangle_diagonal_the_mesh_averages_two_normals(systems):
    """Where the substitute and brax's mesh collider part ways, case 1: the tip's foot point lies ON the diagonal that
    splits the face into two triangles. Both triangles then report the same closest point with the same normal, both
    are active, and (sum / 2) equals the single contact -- but a little OFF the diagonal the far triangle's closest point
    slides onto the diagonal, its normal tilts, and the average is a smaller, tilted impulse: |dv_mesh| < |dv_box|."""
    box, mesh = systems
    x = list(np.isclose(mesh.mesh_tris[3][:, :, 0], 2.0).all(axis=1)).index(True)     # a triangle of box 3 in the plane x = 2
    tri = mesh.mesh_tris[3][x] + np.array([0, 0, 0.5])
    tris = [t + np.array([0, 0, 0.5]) for t in mesh.mesh_tris[3] if np.isclose(t[:, 0], 2.0).all()]
    shared = [v for v in tris[0] if any(np.allclose(v, u) for u in tris[1])]
    assert len(tris) == 2 and len(shared) == 2
    mid = (shared[0] + shared[1]) / 2
    along = (shared[1] - shared[0]) / np.linalg.norm(shared[1] - shared[0])
    perp = np.cross(along, [1.0, 0, 0])                                             # in the face, across the diagonal
    d = 0.06
    for off, equal in ((0.0, True), (0.03, False)):
        tip = mid + off * perp + np.array([-d, 0, 0])
        qp, body, e = _pose(box, 0, tip, (1.0, 0.0, 0.0))
        qp.pos[:, :9, 2] += 0.0
        gb, gm = _info(box, qp, body), _info(mesh, qp, body)
        assert mesh.last_mesh_active[0, body] == 2
        if equal:
            assert np.allclose(gm[0], gb[0], atol=1e-5), (gm, gb)
        else:
            assert np.linalg.norm(gm[0]) < np.linalg.norm(gb[0]) - 1e-3, (gm, gb)
            assert abs(gm[0][1]) + abs(gm[0][2]) > 1e-3 and abs(gb[0][1]) + abs(gb[0][2]) < 1e-6   # tilted vs pure -x


def test_a_capsule_point_inside_a_box(systems):
    """Case 2: a leg poking THROUGH a wall face (HeavenHell spawns 14 % of its ants with a leg through the y = 0 wall).
    The closest point of the capsule's segment to the box then has distance 0 -- for the substitute (a segment point
    inside the box) and for the pierced face's triangle alike: normal exactly zero, no impulse from that face. But
    the mesh collider also tests the box's BOTTOM face (z = 0 under the wall): a foot that is inside the footprint and
    lower than its radius touches that face from above and is pushed UP by it, in addition to the ground plane's own
    impulse. The substitute has no such contact."""
    box, mesh = systems
    high = np.array([2.0 + 0.05, 2.0, 0.30])             # 5 cm inside the box x in [2, 3], well above the floor
    qp, body, e = _pose(box, 0, high, (-1.0, 0.0, 0.0))
    gb, gm = _info(box, qp, body), _info(mesh, qp, body)
    assert np.all(gb[0] == 0) and np.all(gb[1] == 0) and np.all(gm[0] == 0)          # both: nothing
    low = np.array([2.0 + 0.05, 2.0, 0.06])              # the same, 2 cm into the ground
    qp, body, e = _pose(box, 0, low, (0.0, 0.0, -0.5))
    gb, gm = _info(box, qp, body), _info(mesh, qp, body)
    assert gb[0][2] > 0.1                                # the ground plane pushes up in both
    assert gm[0][2] > gb[0][2] + 0.1                     # the mesh's bottom face adds its own upward impulse
