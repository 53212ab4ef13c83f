"""CPU restatement (NumPy, batched over envs) of the Brax v1 "legacy spring" rigid-body pipeline
as the reference uses it for the Ant.

TEST INFRASTRUCTURE ONLY (see oracle/threefry.py header): checker for the CUDA path and the
`cpu_baseline` leg of bench.py; never imported by po_brax_b200/.

The arithmetic lives in the un-vendored third-party package `brax` (/root/reference/setup.py:14,
`brax>=0.0.12`; the Config embedded at /root/reference/notebooks/ant_tag.ipynb:449 has
baumgarteErp/springDamping and no dynamicsMode => pre-PBD build). It is absent from
/root/reference, so this file restates the published brax v0.0.12 algorithm
(brax/physics/{system,integrators,joints,actuators,colliders,geometry,bodies,math}.py) and anchors on
the reference's call sites:
  sys.step          /root/reference/po_brax/envs/ant_heavenhell.py:108, ant_gather.py:127, ant_tag.py:109
  sys.default_qp    ant_heavenhell.py:95, ant_gather.py:116, ant_tag.py:72
  sys.default_angle ant_heavenhell.py:89, ant_gather.py:112, ant_tag.py:66
  sys.info          ant_heavenhell.py:77, ant_gather.py:95, ant_tag.py:81
  joints[0].angle_vel  ant_heavenhell.py:128, ant_gather.py:186, ant_tag.py:156
PINNED by tests/test_oracle_golden.py against the 21-frame rollout embedded in that notebook
(ground-contact physics, default_qp). NOT pinned (no reference artefact exists): velocities,
torso-ground contact, and the capsule-vs-box wall collider, which is a documented, physically
equivalent substitute (segment vs axis-aligned box) for brax's triangulated-box mesh collider --
see DESIGN.md "Wall collider".
"""
import numpy as np


# ----------------------------------------------------------------------------- math (brax/physics/math.py)
def cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def rotate(v, q):
    """math.rotate: 2(u.v)u + (s^2 - u.u)v + 2s(u x v); q = (w, x, y, z)."""
    u = q[..., 1:]
    s = q[..., 0:1]
    two = v.dtype.type(2)
    return two * (dot(u, v)[..., None] * u) + (s * s - dot(u, u)[..., None]) * v + two * s * cross(u, v)


def quat_mul(u, v):
    return np.stack([
        u[..., 0] * v[..., 0] - u[..., 1] * v[..., 1] - u[..., 2] * v[..., 2] - u[..., 3] * v[..., 3],
        u[..., 0] * v[..., 1] + u[..., 1] * v[..., 0] + u[..., 2] * v[..., 3] - u[..., 3] * v[..., 2],
        u[..., 0] * v[..., 2] - u[..., 1] * v[..., 3] + u[..., 2] * v[..., 0] + u[..., 3] * v[..., 1],
        u[..., 0] * v[..., 3] + u[..., 1] * v[..., 2] - u[..., 2] * v[..., 1] + u[..., 3] * v[..., 0],
    ], axis=-1)


def quat_inv(q):
    return q * np.array([1, -1, -1, -1], q.dtype)


def euler_to_quat(deg, dtype):
    """math.euler_to_quat (degrees)."""
    v = np.asarray(deg, dtype)
    c = np.cos(v * dtype(np.pi) / dtype(360))
    s = np.sin(v * dtype(np.pi) / dtype(360))
    c1, c2, c3 = c
    s1, s2, s3 = s
    return np.array([c1 * c2 * c3 - s1 * s2 * s3, s1 * c2 * c3 + c1 * s2 * s3,
                     c1 * s2 * c3 - s1 * c2 * s3, c1 * c2 * s3 + s1 * s2 * c3], dtype)


def quat_rot_axis(axis, angle):
    """math.quat_rot_axis: (cos a/2, axis sin a/2). axis[..., 3], angle[...]."""
    half = angle / angle.dtype.type(2)
    return np.concatenate([np.cos(half)[..., None], axis * np.sin(half)[..., None]], axis=-1)


def safe_norm(x):
    """math.safe_norm: 0 where all components are 0 (the where-guard only matters for gradients)."""
    return np.sqrt((x[..., 0] * x[..., 0] + x[..., 1] * x[..., 1]) + x[..., 2] * x[..., 2])


def _vec(d, default=0.0):
    d = d or {}
    return [float(d.get(k, default)) for k in 'xyz']


class QP:
    """brax.QP: pos[N,nb,3], rot[N,nb,4] (w,x,y,z), vel[N,nb,3], ang[N,nb,3]."""
    __slots__ = ('pos', 'rot', 'vel', 'ang')

    def __init__(self, pos, rot, vel, ang):
        self.pos, self.rot, self.vel, self.ang = pos, rot, vel, ang

    def copy(self):
        return QP(self.pos.copy(), self.rot.copy(), self.vel.copy(), self.ang.copy())

    def take(self, idx):
        return QP(self.pos[idx], self.rot[idx], self.vel[idx], self.ang[idx])


class Info:
    """brax.Info restricted to what the envs read: contact.vel / contact.ang [N,nb,3]."""
    __slots__ = ('contact_vel', 'contact_ang')

    def __init__(self, contact_vel, contact_ang):
        self.contact_vel, self.contact_ang = contact_vel, contact_ang


class System:
    """brax.System for Ant-family configs (1-DoF revolute joints, torque actuators, capsule/sphere
    bodies vs one plane and vs axis-aligned boxes of a frozen Arena body)."""

    def __init__(self, cfg: dict, dtype=np.float32, walls=True):
        """walls: True = the documented box substitute (what the CUDA path implements), False = no Arena colliders,
        'mesh' = brax's capsule-vs-triangulated-box collider as recalled (ANALYSIS ONLY, see _wall_contacts_mesh)."""
        self.cfg = cfg
        self.dtype = dt = dtype
        self.names = [b['name'] for b in cfg['bodies']]
        self.index = {n: i for i, n in enumerate(self.names)}
        self.num_bodies = nb = len(self.names)
        self.mass = np.array([b['mass'] for b in cfg['bodies']], dt)
        self.inv_inertia = np.array([[1.0 / x for x in _vec(b.get('inertia'), 1.0)] for b in cfg['bodies']], dt)
        frozen = np.array([1.0 if (b.get('frozen') or {}).get('all') else 0.0 for b in cfg['bodies']], dt)
        self.active = (1 - frozen).astype(dt)  # pos/rot/vel/ang masks (Ant-family: all-or-nothing)
        self.dt = dt(cfg['dt'])
        self.substeps = int(cfg['substeps'])
        self.h = dt(cfg['dt'] / cfg['substeps'])
        self.gravity = np.array(_vec(cfg.get('gravity')), dt)
        self.vel_damp = dt(np.exp(dt(cfg.get('velocityDamping', 0.0)) * self.h))
        self.ang_damp = dt(np.exp(dt(cfg.get('angularDamping', 0.0)) * self.h))
        self.baumgarte = dt(cfg['baumgarteErp'] * cfg['substeps'] / cfg['dt'])
        self.friction = dt(cfg.get('friction', 1.0))  # material friction a*b = 1 for every pair here
        self.elasticity = dt(cfg.get('elasticity', 0.0))

        # ---- joints (brax/physics/joints.py: Revolute; all 1-DoF, reference_rotation = 0)
        J = cfg['joints']
        self.num_joints = len(J)
        self.j_parent = np.array([self.index[j['parent']] for j in J])
        self.j_child = np.array([self.index[j['child']] for j in J])
        self.j_off_p = np.array([_vec(j.get('parentOffset')) for j in J], dt)
        self.j_off_c = np.array([_vec(j.get('childOffset')) for j in J], dt)
        self.j_stiff = np.array([j['stiffness'] for j in J], dt)
        self.j_sdamp = np.array([j.get('springDamping', 0.0) for j in J], dt)
        self.j_adamp = np.array([j.get('angularDamping', 0.0) for j in J], dt)
        self.j_lstr = np.array([j.get('limitStrength', j['stiffness']) for j in J], dt)
        lim = np.array([[j['angleLimit'][0]['min'], j['angleLimit'][0]['max']] for j in J], dt)
        self.j_limit = (lim * dt(np.pi) / dt(180)).astype(dt)
        eye = np.eye(3, dtype=dt)
        self.j_axis = np.stack([
            np.stack([rotate(eye[i], euler_to_quat(_vec(j.get('rotation')), dt)) for i in range(3)])
            for j in J]).astype(dt)  # [J, 3(axis idx), 3]
        # ---- actuators (torque), actuator k <-> joint by name
        jname = {j['name']: i for i, j in enumerate(J)}
        self.a_joint = np.array([jname[a['joint']] for a in cfg['actuators']])
        self.a_strength = np.array([a['strength'] for a in cfg['actuators']], dt)

        # ---- colliders
        include = {(c['first'], c['second']) for c in cfg.get('collideInclude', [])}
        include |= {(b, a) for a, b in include}
        plane_bodies = [b['name'] for b in cfg['bodies'] if any('plane' in c for c in b.get('colliders', []))]
        self.ground = self.index[plane_bodies[0]] if plane_bodies else None
        cp_body, cp_end, cp_rad = [], [], []       # capsule-end / plane candidates (geometry.CapsuleEnd)
        cap_body, cap_a, cap_b, cap_rad = [], [], [], []   # full capsules (for box walls)
        self.body_caps = {}                        # body idx -> list of (end points, radius) for default_qp
        for bi, b in enumerate(cfg['bodies']):
            for c in b.get('colliders', []):
                if 'sphere' in c:  # spheres = capsules with one end and zero segment
                    r, seg, ends = c['sphere']['radius'], 0.0, [1]
                elif 'capsule' in c:
                    r = c['capsule']['radius']
                    seg = c['capsule']['length'] / 2 - r
                    e = c['capsule'].get('end', 0)
                    ends = [-1, 1] if e == 0 else [e]
                else:
                    continue
                axis = rotate(eye[2], euler_to_quat(_vec(c.get('rotation')), dt))
                cpos = np.array(_vec(c.get('position')), dt)
                pts = [cpos + dt(e) * axis * dt(seg) for e in ends]
                self.body_caps.setdefault(bi, []).append((pts, dt(r)))
                if self.ground is not None and (b['name'], self.names[self.ground]) in include:
                    for p in pts:
                        cp_body.append(bi); cp_end.append(p); cp_rad.append(r)
                if (b['name'], 'Arena') in include:
                    cap_body.append(bi)
                    cap_a.append(cpos - axis * dt(seg)); cap_b.append(cpos + axis * dt(seg)); cap_rad.append(r)
        self.cp_body = np.array(cp_body, int)
        self.cp_end = np.array(cp_end, dt).reshape(-1, 3)
        self.cp_rad = np.array(cp_rad, dt)
        self.cap_body = np.array(cap_body, int)
        self.cap_a = np.array(cap_a, dt).reshape(-1, 3)
        self.cap_b = np.array(cap_b, dt).reshape(-1, 3)
        self.cap_rad = np.array(cap_rad, dt)
        # Arena boxes -> axis-aligned (lo, hi) in the Arena body frame; z-rotations are 0/90/180 deg
        self.arena = self.index.get('Arena')
        boxes, self.box_corners_z = [], {}
        for bi, b in enumerate(cfg['bodies']):
            for c in b.get('colliders', []):
                if 'box' not in c:
                    continue
                hs = np.array(_vec(c['box']['halfsize']), np.float64)
                rz = _vec(c.get('rotation'))[2]
                k = int(round(rz / 90.0))
                assert abs(rz - 90.0 * k) < 1e-3 and _vec(c.get('rotation'))[:2] == [0.0, 0.0], \
                    'wall boxes must be axis aligned'
                if k % 2:
                    hs = hs[[1, 0, 2]]
                cpos = np.array(_vec(c.get('position')), np.float64)
                self.box_corners_z.setdefault(bi, []).append(cpos[2] - hs[2])
                if bi == self.arena:
                    boxes.append(np.concatenate([cpos - hs, cpos + hs]))
        self.boxes = np.array(boxes, dt).reshape(-1, 6) if walls else np.zeros((0, 6), dt)
        self.wall_mode = 'mesh' if walls == 'mesh' else 'box'
        if self.wall_mode == 'mesh':
            self._build_mesh(cfg)
        self.defaults = cfg.get('defaults', [])
        self.track_margin = False   # test aid, see _note_margin
        self.margin = None

    # ------------------------------------------------------------------ default_angle / default_qp
    def default_angle(self):
        """System.default_angle: midpoint of each joint limit (radians)."""
        return ((self.j_limit[:, 0] + self.j_limit[:, 1]) / self.dtype(2)).astype(self.dtype)

    def default_qp(self, joint_angle, joint_velocity) -> QP:
        """System.default_qp(joint_angle[N,J], joint_velocity[N,J])."""
        dt = self.dtype
        N, nb = joint_angle.shape[0], self.num_bodies
        pos = np.zeros((N, nb, 3), dt)
        rot = np.zeros((N, nb, 4), dt); rot[..., 0] = 1
        vel = np.zeros((N, nb, 3), dt)
        ang = np.zeros((N, nb, 3), dt)
        overridden = set()
        if self.defaults:  # only defaults[0] is applied
            for q in self.defaults[0].get('qps', []):
                bi = self.index[q['name']]
                overridden.add(bi)
                pos[:, bi] = np.array(_vec(q.get('pos')), dt)
        # joints in depth order: parents before children
        depth = {}
        for j in range(self.num_joints):
            d, p = 0, self.j_parent[j]
            while p in set(self.j_child):
                p = self.j_parent[list(self.j_child).index(p)]; d += 1
            depth[j] = d
        for j in sorted(range(self.num_joints), key=lambda j: (depth[j], j)):
            p, c = self.j_parent[j], self.j_child[j]
            axis = np.broadcast_to(self.j_axis[j, 0], (N, 3))
            local_rot = quat_rot_axis(axis, joint_angle[:, j])
            local_ang = axis * joint_velocity[:, j:j + 1]
            rot[:, c] = quat_mul(rot[:, p], local_rot)
            off_c = rotate(np.broadcast_to(self.j_off_c[j], (N, 3)), local_rot)
            pos[:, c] = pos[:, p] + rotate(self.j_off_p[j] - off_c, rot[:, p])
            ang[:, c] = rotate(local_ang, rot[:, p])
        # z-lift every kinematic tree whose root has no default override so its lowest point is at z=0
        root = {}
        for bi in range(nb):
            r = bi
            while r in set(self.j_child):
                r = self.j_parent[list(self.j_child).index(r)]
            root[bi] = r
        for r in sorted(set(root.values())):
            if r in overridden:
                continue
            members = [b for b in range(nb) if root[b] == r]
            min_z = None
            for b in members:
                for pts, rad in self.body_caps.get(b, []):
                    for p in pts:
                        z = pos[:, b, 2] + rotate(np.broadcast_to(p, (N, 3)), rot[:, b])[:, 2] - rad
                        min_z = z if min_z is None else np.minimum(min_z, z)
                for z0 in self.box_corners_z.get(b, []):
                    z = pos[:, b, 2] + dt(z0)
                    min_z = z if min_z is None else np.minimum(min_z, z)
            if min_z is None:  # planes: nothing to lift
                continue
            for b in members:
                pos[:, b, 2] -= min_z
        return QP(pos, rot, vel, ang)

    # ------------------------------------------------------------------ joints
    def _axis_angle(self, qp: QP):
        """Revolute.axis_angle: (axis_p [N,J,3], psi [N,J])."""
        rot_p, rot_c = qp.rot[:, self.j_parent], qp.rot[:, self.j_child]
        axis_p = rotate(np.broadcast_to(self.j_axis[:, 0], rot_p.shape[:2] + (3,)), rot_p)
        ref_p = rotate(np.broadcast_to(self.j_axis[:, 2], rot_p.shape[:2] + (3,)), rot_p)
        ref_c = rotate(np.broadcast_to(self.j_axis[:, 2], rot_p.shape[:2] + (3,)), rot_c)
        psi = np.arctan2(dot(cross(ref_p, ref_c), axis_p), dot(ref_p, ref_c))
        return axis_p, psi

    def angle_vel(self, qp: QP):
        """joints[0].angle_vel(qp): (angles[N,J], vels[N,J]); vel = (ang_p - ang_c) . axis_p."""
        axis_p, psi = self._axis_angle(qp)
        v = dot(qp.ang[:, self.j_parent] - qp.ang[:, self.j_child], axis_p)
        return psi.astype(self.dtype), v.astype(self.dtype)

    def _joints_and_actuators(self, qp: QP, act):
        """Sum of Revolute.apply + Torque.apply: (dvel[N,nb,3], dang[N,nb,3])."""
        N, nb = qp.pos.shape[0], self.num_bodies
        P, C = self.j_parent, self.j_child
        pos_p, rot_p, vel_p, ang_p = qp.pos[:, P], qp.rot[:, P], qp.vel[:, P], qp.ang[:, P]
        pos_c, rot_c, vel_c, ang_c = qp.pos[:, C], qp.rot[:, C], qp.vel[:, C], qp.ang[:, C]
        shp = pos_p.shape
        rp = rotate(np.broadcast_to(self.j_off_p, shp), rot_p)
        rc = rotate(np.broadcast_to(self.j_off_c, shp), rot_c)
        wp, wc = pos_p + rp, pos_c + rc
        wvp, wvc = vel_p + cross(ang_p, rp), vel_c + cross(ang_c, rc)
        F = (wp - wc) * self.j_stiff[:, None] + self.j_sdamp[:, None] * (wvp - wvc)
        inv_ip, inv_ic = self.inv_inertia[P], self.inv_inertia[C]
        dvel_p = -F / self.mass[P][:, None]
        dang_p = inv_ip * cross(wp - pos_p, -F)
        dvel_c = F / self.mass[C][:, None]
        dang_c = inv_ic * cross(wc - pos_c, F)
        axis_p, psi = self._axis_angle(qp)
        axis_c = rotate(np.broadcast_to(self.j_axis[:, 0], shp), rot_c)
        torque = self.j_stiff[:, None] * cross(axis_p, axis_c)
        lo, hi = self.j_limit[:, 0], self.j_limit[:, 1]
        zero = self.dtype(0)
        dang = np.where(psi < lo, lo - psi, zero)
        dang = np.where(psi > hi, hi - psi, dang)
        torque = torque - self.j_lstr[:, None] * axis_p * dang[..., None]
        torque = torque - self.j_adamp[:, None] * (ang_p - ang_c)
        dang_p = dang_p + inv_ip * torque
        dang_c = dang_c + inv_ic * (-torque)
        # actuators: actuator k drives joint a_joint[k]; zeroed outside the limits; -axis on the parent
        aj = self.a_joint
        t = act * self.a_strength  # brax: where(limits) then *= strength; same value
        out = (psi[:, aj] < lo[aj]) | (psi[:, aj] > hi[aj])
        t = np.where(out, zero, t).astype(self.dtype)
        if self.track_margin:  # the actuator switches off discontinuously at the joint limits
            m = np.minimum(np.abs(psi - lo), np.abs(psi - hi)).astype(np.float64).min(axis=-1) * 100.0
            self.margin = m if self.margin is None else np.minimum(self.margin, m)
        a_tau = -(axis_p[:, aj] * t[..., None])
        a_dang_p = self.inv_inertia[P[aj]] * a_tau
        a_dang_c = self.inv_inertia[C[aj]] * (-a_tau)
        dvel = np.zeros((N, nb, 3), self.dtype)
        dangs = np.zeros((N, nb, 3), self.dtype)
        # segment_sum in index order: joint parents, then joint children (dp_j), then dp_a added
        for j in range(self.num_joints):
            dvel[:, P[j]] += dvel_p[:, j]; dangs[:, P[j]] += dang_p[:, j]
        for j in range(self.num_joints):
            dvel[:, C[j]] += dvel_c[:, j]; dangs[:, C[j]] += dang_c[:, j]
        adang = np.zeros((N, nb, 3), self.dtype)
        for k in range(len(aj)):
            adang[:, P[aj[k]]] += a_dang_p[:, k]
        for k in range(len(aj)):
            adang[:, C[aj[k]]] += a_dang_c[:, k]
        return dvel, dangs + adang

    # ------------------------------------------------------------------ contacts
    def _impulse(self, qp_pos, body, cpos, cvel, normal, pen, degenerate=None, surface=None):
        """OneWayCollider._contact. All [N,K,...]; body = body index per K.
        degenerate [N,K] (test aid only): contacts whose normal is exactly the zero vector -- nv = J-part = 0 exactly
        in any evaluation order, so they are not rounding-ambiguous and stay out of the margin bookkeeping.
        surface [N,K] (test aid only): an extra margin per contact (see _wall_contacts)."""
        zero, one = self.dtype(0), self.dtype(1)
        inv_m = (one / self.mass[body])
        inv_i = self.inv_inertia[body]
        rel = cpos - qp_pos
        bv = self.baumgarte * pen
        nv = dot(normal, cvel)
        temp1 = inv_i * cross(rel, normal)
        ang = dot(normal, cross(temp1, rel))
        denom = inv_m + ang
        J = (-one * (one + self.elasticity) * nv + bv) / denom
        Jn = J[..., None] * normal
        dpn_vel = Jn / self.mass[body][:, None]
        dpn_ang = inv_i * cross(rel, Jn)
        vel_d = cvel - nv[..., None] * normal
        nd = safe_norm(vel_d)
        Jd = np.minimum(nd / denom, self.friction * J)
        dir_d = vel_d / (self.dtype(1e-6) + nd)[..., None]
        Jdv = -Jd[..., None] * dir_d
        dpd_vel = Jdv / self.mass[body][:, None]
        dpd_ang = inv_i * cross(rel, Jdv)
        apply_n = np.where((pen > zero) & (nv < zero) & (J > zero), one, zero)
        apply_d = apply_n * np.where(nd > self.dtype(0.01), one, zero)
        if self.track_margin:
            self._note_margin(pen, nv, J, nd, degenerate, surface)
        dvel = dpn_vel * apply_n[..., None] + dpd_vel * apply_d[..., None]
        dang = dpn_ang * apply_n[..., None] + dpd_ang * apply_d[..., None]
        return dvel.astype(self.dtype), dang.astype(self.dtype)

    def _note_margin(self, pen, nv, J, nd, degenerate=None, surface=None):
        """Test aid: distance of each contact from its nearest discontinuous branch (pen > 0, nv < 0, J > 0,
        |v_d| > 0.01; the actuator cut-off at the joint limits is noted in _joints_and_actuators). An env whose margin is ~1 ulp may legitimately take the other branch in a different
        float32 evaluation order (the reference's XLA program included); parity tests skip those envs."""
        big = np.float64(1e9)
        pen, nv, J, nd = (np.asarray(x, np.float64) for x in (pen, nv, J, nd))
        m = np.abs(pen) * 100.0                                    # touching / not touching (length -> velocity scale)
        live = pen > 0
        m = np.minimum(m, np.where(live, np.abs(nv), big))         # approaching / separating
        m = np.minimum(m, np.where(live & (nv < 0), np.abs(J), big))
        m = np.minimum(m, np.where(live & (nv < 0) & (J > 0), np.abs(nd - 0.01), big))
        if degenerate is not None:
            m = np.where(degenerate, big, m)
        if surface is not None:
            m = np.minimum(m, surface)
        m = m.min(axis=-1)
        self.margin = m if self.margin is None else np.minimum(self.margin, m)

    def _group_reduce(self, N, body, dvel, dang):
        """Collider.apply tail: per body, sum contacts and divide by (1e-8 + #contacts with any(dvel != 0))."""
        nb = self.num_bodies
        cnt = np.zeros((N, nb), self.dtype)
        sv = np.zeros((N, nb, 3), self.dtype)
        sa = np.zeros((N, nb, 3), self.dtype)
        hit = np.any(dvel != 0, axis=-1).astype(self.dtype)
        for k in range(len(body)):
            cnt[:, body[k]] += hit[:, k]
            sv[:, body[k]] += dvel[:, k]
            sa[:, body[k]] += dang[:, k]
        d = (self.dtype(1e-8) + cnt)[..., None]
        return sv / d, sa / d

    def _ground_contacts(self, qp: QP):
        """capsule_plane (geometry.CapsuleEnd vs Plane)."""
        N = qp.pos.shape[0]
        if self.ground is None or len(self.cp_body) == 0:
            z = np.zeros((N, self.num_bodies, 3), self.dtype)
            return z, z.copy()
        b = self.cp_body
        pos, rot, vel, ang = qp.pos[:, b], qp.rot[:, b], qp.vel[:, b], qp.ang[:, b]
        g = self.ground
        n = rotate(np.broadcast_to(np.array([0, 0, 1], self.dtype), (N, 1, 3)), qp.rot[:, g:g + 1])
        n = np.broadcast_to(n, pos.shape)
        end_w = pos + rotate(np.broadcast_to(self.cp_end, pos.shape), rot)
        cpos = end_w - n * self.cp_rad[:, None]
        cvel = vel + cross(ang, cpos - pos)
        pen = dot(qp.pos[:, g:g + 1] - cpos, n)
        dvel, dang = self._impulse(pos, b, cpos, cvel, n, pen)
        return self._group_reduce(N, b, dvel, dang)

    def _closest_segment_box(self, a, b, lo, hi):
        """Closest points between segment [a,b] and axis-aligned box [lo,hi] (all [...,3]).
        dist^2(t) along the segment is convex and C1; its derivative g(t) is monotone piecewise
        linear: 16 bisection steps then one false-position step (exact unless a breakpoint falls in
        the final 2^-16 bracket). Returns (seg_pt, box_pt)."""
        dt = self.dtype
        d = b - a

        def g(t):
            p = a + t[..., None] * d
            return dot(p - np.clip(p, lo, hi), d)

        t0 = np.zeros(a.shape[:-1], dt)
        t1 = np.ones(a.shape[:-1], dt)
        g0, g1 = g(t0), g(t1)
        tl, tr, gl, gr = t0.copy(), t1.copy(), g0.copy(), g1.copy()
        for _ in range(16):
            tm = dt(0.5) * (tl + tr)
            gm = g(tm)
            left = gm > 0  # root is to the left of tm
            tr = np.where(left, tm, tr); gr = np.where(left, gm, gr)
            tl = np.where(left, tl, tm); gl = np.where(left, gl, gm)
        den = gr - gl
        ts = np.where(den > 0, tl - gl * (tr - tl) / np.where(den > 0, den, dt(1)), tl)
        t = np.where(g0 >= 0, t0, np.where(g1 <= 0, t1, ts)).astype(dt)
        p = a + t[..., None] * d
        return p, np.clip(p, lo, hi)

    def _wall_contacts(self, qp: QP):
        """Capsule vs Arena boxes. Substitute for brax's capsule_mesh/TriangulatedBox:
        one contact per (capsule, box) at the closest box point; normal = (seg_pt - box_pt)/(1e-6+d);
        penetration = r - d; a segment point inside the box (d = 0) gives no impulse."""
        N = qp.pos.shape[0]
        nbx = len(self.boxes)
        if nbx == 0 or len(self.cap_body) == 0:
            z = np.zeros((N, self.num_bodies, 3), self.dtype)
            return z, z.copy()
        b = np.repeat(self.cap_body, nbx)               # K = caps * boxes, box index fastest
        ca = np.repeat(self.cap_a, nbx, axis=0)
        cb = np.repeat(self.cap_b, nbx, axis=0)
        rad = np.repeat(self.cap_rad, nbx)
        box = np.tile(self.boxes, (len(self.cap_body), 1))
        pos, rot, vel, ang = qp.pos[:, b], qp.rot[:, b], qp.vel[:, b], qp.ang[:, b]
        apos = qp.pos[:, self.arena][:, None, :]
        a_w = pos + rotate(np.broadcast_to(ca, pos.shape), rot)
        b_w = pos + rotate(np.broadcast_to(cb, pos.shape), rot)
        lo, hi = apos + box[:, :3], apos + box[:, 3:]
        seg_p, box_p = self._closest_segment_box(a_w, b_w, lo, hi)
        dvec = seg_p - box_p
        dist = safe_norm(dvec)
        n = dvec / (self.dtype(1e-6) + dist)[..., None]
        pen = rad - dist
        cvel = vel + cross(ang, box_p - pos)
        # a segment point inside the box: dvec = 0 exactly => n = 0, nv = 0, no impulse, in any evaluation order
        surface = None
        if self.track_margin:
            # ... but AT the surface the normal dvec / (1e-6 + dist) swings from 0 to unit length within ~1e-5 m: a
            # closest point that close to a box face (outside: dist < 5e-5; inside: depth < 5e-6) is rounding-
            # ambiguous although no predicate switches. Margin in the units of _note_margin.
            depth = np.minimum(seg_p - lo, hi - seg_p).min(axis=-1).astype(np.float64)
            surface = np.where(dist > 0, dist.astype(np.float64) * 10.0, np.where(depth >= 0, depth * 100.0, 1e9))
        dvel, dang = self._impulse(pos, b, box_p, cvel, n, pen, degenerate=(dist == 0), surface=surface)
        return self._group_reduce(N, b, dvel, dang)

    # ------------------------------------------------------------------ brax's mesh wall collider (analysis only)
    # brax 0.0.12 resolves ('capsule', 'box') pairs through capsule_mesh against a TriangulatedBox. RECALLED from the
    # published source (brax/physics/{geometry,colliders,geometry_utils? math}.py), not pinned by any reference artefact:
    #   corners  = itertools.product((-1, 1), repeat=3) * halfsize       (corner i: x = bit 2, y = bit 1, z = bit 0)
    #   faces    = [0,4,1, 4,5,1 | 0,2,4, 2,6,4 | 6,5,4, 6,7,5 | 2,3,6, 3,7,6 | 1,5,3, 5,7,3 | 0,1,2, 1,3,2]
    #              (-y, -z, +x, +y, +z, -x; two triangles per side), rotated by the collider's euler rotation
    #   per (capsule, triangle): closest_segment_triangle_points -> contact at the TRIANGLE point, normal
    #   (seg_pt - tri_pt) / (1e-6 + dist), penetration radius - dist; all triangle contacts of all boxes for one
    #   body form one group: summed and divided by the number of active ones.
    # Used by oracle/wall_deviation.py and tests/test_wall_collider.py to quantify how the box substitute differs.
    _BOX_FACES = np.array([0, 4, 1, 4, 5, 1, 0, 2, 4, 2, 6, 4, 6, 5, 4, 6, 7, 5, 2, 3, 6, 3, 7, 6, 1, 5, 3, 5, 7, 3,
                           0, 1, 2, 1, 3, 2]).reshape(12, 3)
    _BOX_NORMALS = np.repeat(np.array([[0, -1., 0], [0, 0, -1.], [1., 0, 0], [0, 1., 0], [0, 0, 1.], [-1., 0, 0]]), 2, axis=0)

    def _build_mesh(self, cfg):
        dt = self.dtype
        corners = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], np.float64)
        tris, normals = [], []
        for b in cfg['bodies']:
            if b['name'] != 'Arena':
                continue
            for c in b.get('colliders', []):
                if 'box' not in c:
                    continue
                hs = np.array(_vec(c['box']['halfsize']), np.float64)
                q = euler_to_quat(_vec(c.get('rotation')), np.float64)
                v = rotate(corners * hs, np.broadcast_to(q, (8, 4))) + np.array(_vec(c.get('position')), np.float64)
                tris.append(v[self._BOX_FACES])                                           # [12, 3, 3]
                normals.append(rotate(self._BOX_NORMALS, np.broadcast_to(q, (12, 4))))
        self.mesh_tris = np.array(tris, dt)          # [nbox, 12, 3 vertices, 3] in the Arena body frame
        self.mesh_normals = np.array(normals, dt)    # [nbox, 12, 3]

    @staticmethod
    def _closest_segment_point(a, b, pt):
        ab = b - a
        t = dot(pt - a, ab) / (dot(ab, ab) + 1e-6)
        return a + np.clip(t, 0, 1)[..., None] * ab

    def _closest_segment_segment(self, a0, a1, b0, b1):
        """brax geometry.closest_segment_to_segment_points."""
        def normalize(v):
            n = safe_norm(v)
            return v / (1e-6 + n)[..., None], n
        dir_a, len_a = normalize(a1 - a0)
        dir_b, len_b = normalize(b1 - b0)
        ha, hb = len_a * 0.5, len_b * 0.5
        a_mid, b_mid = a0 + dir_a * ha[..., None], b0 + dir_b * hb[..., None]
        trans = a_mid - b_mid
        dd, dat, dbt = dot(dir_a, dir_b), dot(dir_a, trans), dot(dir_b, trans)
        denom = 1 - dd * dd
        ta = (-dat + dd * dbt) / (denom + 1e-6)
        tb = dbt + ta * dd
        ta, tb = np.clip(ta, -ha, ha), np.clip(tb, -hb, hb)
        best_a, best_b = a_mid + dir_a * ta[..., None], b_mid + dir_b * tb[..., None]
        new_a = self._closest_segment_point(a0, a1, best_b)
        new_b = self._closest_segment_point(b0, b1, best_a)
        d1, d2 = dot(new_a - best_b, new_a - best_b), dot(new_b - best_a, new_b - best_a)
        pick = (d1 < d2)[..., None]
        return np.where(pick, new_a, best_a), np.where(pick, best_b, new_b)

    def _closest_segment_triangle(self, a, b, p0, p1, p2, n):
        """brax geometry.closest_segment_triangle_points: min over the three edge pairs and the plane candidate (the
        segment point nearest the plane, if its projection falls inside the triangle); ties averaged."""
        cands = [self._closest_segment_segment(a, b, p0, p1), self._closest_segment_segment(a, b, p1, p2),
                 self._closest_segment_segment(a, b, p0, p2)]
        d = dot(p0, n)
        den = dot(n, b - a)
        t = (d - dot(n, a)) / (den + 1e-6 * (den == 0))
        seg4 = a + np.clip(t, 0, 1)[..., None] * (b - a)
        tri4 = seg4 - (dot(seg4, n) - d)[..., None] * n
        # inside test: same side of all three edges
        def side(u, v):
            return dot(cross(v - u, tri4 - u), n)
        s0, s1, s2 = side(p0, p1), side(p1, p2), side(p2, p0)
        inside = ((s0 >= 0) & (s1 >= 0) & (s2 >= 0)) | ((s0 <= 0) & (s1 <= 0) & (s2 <= 0))
        dist = [dot(sp - tp, sp - tp) for sp, tp in cands]
        dist.append(np.where(inside, dot(seg4 - tri4, seg4 - tri4), np.inf))
        cands.append((seg4, tri4))
        dist = np.stack(dist, axis=-1)
        mask = (dist == dist.min(axis=-1, keepdims=True)).astype(a.dtype)
        seg = sum(c[0] * mask[..., i:i + 1] for i, c in enumerate(cands)) / mask.sum(-1, keepdims=True)
        tri = sum(c[1] * mask[..., i:i + 1] for i, c in enumerate(cands)) / mask.sum(-1, keepdims=True)
        return seg, tri

    def _wall_contacts_mesh(self, qp: QP):
        """capsule_mesh over every (ant capsule, Arena box) pair whose bounding boxes come within the capsule radius
        (exact cull: further apart, all 12 triangle contacts have penetration < 0 and contribute exactly zero)."""
        N, nb = qp.pos.shape[0], self.num_bodies
        zv, za = np.zeros((N, nb, 3), self.dtype), np.zeros((N, nb, 3), self.dtype)
        if len(self.boxes) == 0 or len(self.cap_body) == 0:
            return zv, za
        apos = qp.pos[:, self.arena]                                           # [N,3]
        cb = self.cap_body
        pos, rot = qp.pos[:, cb], qp.rot[:, cb]
        a_w = pos + rotate(np.broadcast_to(self.cap_a, pos.shape), rot)        # [N,K,3]
        b_w = pos + rotate(np.broadcast_to(self.cap_b, pos.shape), rot)
        smin, smax = np.minimum(a_w, b_w), np.maximum(a_w, b_w)
        lo = apos[:, None, None, :] + self.boxes[None, None, :, :3]            # [N,1,X,3]
        hi = apos[:, None, None, :] + self.boxes[None, None, :, 3:]
        gap = np.maximum(lo - smax[:, :, None, :], smin[:, :, None, :] - hi).max(-1)   # [N,K,X]
        e, k, x = np.nonzero(gap < (self.cap_rad[None, :, None] + 1e-3))
        if len(e) == 0:
            return zv, za
        body = cb[k]
        a, b = a_w[e, k][:, None, :], b_w[e, k][:, None, :]                    # [M,1,3]
        tri = apos[e][:, None, None, :] + self.mesh_tris[x]                    # [M,12,3,3]
        n = self.mesh_normals[x]                                               # [M,12,3] (Arena rot = identity)
        sp, tp = self._closest_segment_triangle(np.broadcast_to(a, tri[:, :, 0].shape), np.broadcast_to(b, tri[:, :, 0].shape),
                                                tri[:, :, 0], tri[:, :, 1], tri[:, :, 2], n)
        dvec = sp - tp
        dist = safe_norm(dvec)
        normal = dvec / (self.dtype(1e-6) + dist)[..., None]
        pen = self.cap_rad[k][:, None] - dist
        bp, bv, bw = qp.pos[e, body][:, None, :], qp.vel[e, body][:, None, :], qp.ang[e, body][:, None, :]
        cvel = bv + cross(np.broadcast_to(bw, tp.shape), tp - bp)
        keep, self.track_margin = self.track_margin, False
        dvel, dang = self._impulse(bp, body[:, None], tp, cvel, normal, pen)
        self.track_margin = keep
        hit = np.any(dvel != 0, axis=-1).astype(self.dtype)                    # [M,12]
        cnt = np.zeros((N, nb), self.dtype)
        np.add.at(cnt, (e, body), hit.sum(-1))
        np.add.at(zv, (e, body), dvel.sum(1))
        np.add.at(za, (e, body), dang.sum(1))
        d = (self.dtype(1e-8) + cnt)[..., None]
        self.last_mesh_active = cnt
        prev = getattr(self, 'mesh_active_max', None)   # analysis aid: callers reset it to None before a step
        self.mesh_active_max = cnt if prev is None else np.maximum(prev, cnt)
        return (zv / d).astype(self.dtype), (za / d).astype(self.dtype)

    def _contacts(self, qp: QP):
        gv, ga = self._ground_contacts(qp)
        wv, wa = self._wall_contacts_mesh(qp) if self.wall_mode == 'mesh' else self._wall_contacts(qp)
        return gv + wv, ga + wa

    # ------------------------------------------------------------------ step / info
    def substep(self, qp: QP, act):
        h, m = self.h, self.active[:, None]
        # kinetic
        pos = qp.pos + qp.vel * h * m
        raq = np.concatenate([np.zeros_like(qp.ang[..., :1]), qp.ang * m], axis=-1) * self.dtype(0.5) * h
        rot = qp.rot + quat_mul(raq, qp.rot)
        rot = rot / np.sqrt(np.sum(rot * rot, axis=-1, keepdims=True))
        qp = QP(pos, rot, qp.vel, qp.ang)
        # joints + actuators -> potential
        dvel, dang = self._joints_and_actuators(qp, act)
        vel = self.vel_damp * qp.vel
        vel = (vel + (dvel + self.gravity) * h) * m
        ang = self.ang_damp * qp.ang
        ang = (ang + dang * h) * m
        qp = QP(qp.pos, qp.rot, vel.astype(self.dtype), ang.astype(self.dtype))
        # contacts -> collision velocity update
        cvel, cang = self._contacts(qp)
        qp = QP(qp.pos, qp.rot, (qp.vel + cvel) * m, (qp.ang + cang) * m)
        return qp, cvel, cang

    def step(self, qp: QP, act):
        """System.step(qp, act[N,8]) -> (qp, Info) : `substeps` substeps, contact impulses accumulated."""
        act = np.asarray(act, self.dtype)
        cv = np.zeros_like(qp.pos)
        ca = np.zeros_like(qp.pos)
        for _ in range(self.substeps):
            qp, dv, da = self.substep(qp, act)
            cv = cv + dv
            ca = ca + da
        return qp, Info(cv, ca)

    def info(self, qp: QP) -> Info:
        """System.info(qp): one collider evaluation, no integration."""
        cv, ca = self._contacts(qp)
        return Info(cv, ca)
