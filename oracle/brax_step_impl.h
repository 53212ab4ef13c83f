/* Scalar C restatement of oracle/brax_v1.py `System.step` (one env at a time, OpenMP over envs).
 *
 * TEST INFRASTRUCTURE ONLY: CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) and a second checker
 * for tests/. Never linked, loaded or called by po_brax_b200/.
 *
 * Restates the published brax v0.0.12 "legacy spring" pipeline (brax/physics/{system,integrators,joints,
 * actuators,colliders,geometry}.py; absent from /root/reference, pinned only as `brax>=0.0.12`,
 * /root/reference/setup.py:14) at the reference's call sites `self.sys.step(state.qp, action)`:
 * /root/reference/po_brax/envs/ant_heavenhell.py:108, ant_gather.py:127, ant_tag.py:109.
 * Every expression keeps the evaluation order of oracle/brax_v1.py (compile with -ffp-contract=off), so the
 * float64 build agrees with the NumPy float64 oracle to rounding of libm's atan2 and the float32 build to a few ulp.
 * The wall collider is the documented substitute of DESIGN.md ("Wall collider"); pairs whose bounding boxes are
 * further apart than the capsule radius are skipped -- their impulse is exactly zero in the NumPy oracle too.
 *
 * This header is a template: brax_step.c includes it once per REAL type with SUFFIX set.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#ifndef BRAX_BRANCH_CTL
#define BRAX_BRANCH_CTL
/* Test aid (tests/_parity.py "two-branch check"): every DISCONTINUOUS predicate of the algorithm -- a contact's
 * pen > 0, nv < 0, J > 0, |v_d| > 0.01 and the actuator cut-off at a joint limit -- goes through branch_pred with
 * its distance from the switching point (velocity units; lengths and angles x100). The distance feeds the env's
 * margin (brax_v1._note_margin); with a BranchCtl, evaluations closer than `thr` are counted (in evaluation order)
 * and the k-th one is INVERTED when bit k of `mask` is set, so a test can ask "what if rounding had taken the
 * other branch here?". cause[]: the smallest distance seen per predicate class (0-3 ground pen/nv/J/nd, 4-7 the
 * same for walls, 8 actuator cut-off). mask = 0 changes nothing. */
#define BRAX_NCAUSE 9
typedef struct { double thr; unsigned mask; int count; double cause[BRAX_NCAUSE]; } BranchCtl;
static inline int branch_pred(int truth, double dist, int cause, BranchCtl *bc, double *mg) {
    if (mg && dist < *mg) *mg = dist;
    if (bc) {
        if (dist < bc->cause[cause]) bc->cause[cause] = dist;
        if (dist < bc->thr) {
            int k = bc->count++;
            if (k < 32 && ((bc->mask >> k) & 1u)) truth = !truth;
        }
    }
    return truth;
}
#endif

typedef struct {
    int nb, nj, na, ncp, ncap, nbox, substeps, ground, arena;
    REAL h, vel_damp, ang_damp, baumgarte, friction, elasticity;
    REAL gravity[3];
    const REAL *mass, *inv_inertia, *active;                 /* [nb], [nb][3], [nb] */
    const int *j_parent, *j_child;                           /* [nj] */
    const REAL *j_off_p, *j_off_c;                           /* [nj][3] */
    const REAL *j_stiff, *j_sdamp, *j_adamp, *j_lstr;        /* [nj] */
    const REAL *j_limit;                                     /* [nj][2] radians */
    const REAL *j_axis;                                      /* [nj][3][3] */
    const int *a_joint; const REAL *a_strength;              /* [na] */
    const int *cp_body; const REAL *cp_end, *cp_rad;         /* [ncp], [ncp][3], [ncp] */
    const int *cap_body; const REAL *cap_a, *cap_b, *cap_rad;/* [ncap], [ncap][3] x2, [ncap] */
    const REAL *boxes;                                       /* [nbox][6] lo, hi in the Arena frame */
} FN(SysDesc);

#define MAXB 32   /* bodies */
#define MAXJ 16   /* joints */

static inline REAL FN(dot3)(const REAL *a, const REAL *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

static inline void FN(cross3)(const REAL *a, const REAL *b, REAL *o) {
    REAL x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}

/* brax_v1.rotate: 2(u.v)u + (s^2 - u.u)v + (2s)(u x v) */
static inline void FN(rotate)(const REAL *v, const REAL *q, REAL *o) {
    const REAL *u = q + 1; REAL s = q[0], c[3];
    REAL uv = FN(dot3)(u, v), k = s * s - FN(dot3)(u, u), s2 = (REAL)2 * s;
    FN(cross3)(u, v, c);
    for (int i = 0; i < 3; i++) o[i] = ((REAL)2 * (uv * u[i]) + k * v[i]) + s2 * c[i];
}

static inline void FN(quat_mul)(const REAL *u, const REAL *v, REAL *o) {
    REAL w = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
    REAL x = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
    REAL y = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
    REAL z = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}

/* brax_v1.System._impulse for one contact of body b; adds nothing, returns dvel/dang of this contact. */
static inline void FN(impulse)(const FN(SysDesc) *S, int b, const REAL *bpos, const REAL *cpos, const REAL *cvel,
                               const REAL *n, REAL pen, REAL *dvel, REAL *dang, double *mg, BranchCtl *bc, int wall) {
    const REAL one = 1, zero = 0;
    REAL mass = S->mass[b]; const REAL *ii = S->inv_inertia + 3 * b;
    REAL inv_m = one / mass;
    REAL rel[3] = {cpos[0] - bpos[0], cpos[1] - bpos[1], cpos[2] - bpos[2]};
    REAL bv = S->baumgarte * pen;
    REAL nv = FN(dot3)(n, cvel);
    REAL t1[3], t2[3];
    FN(cross3)(rel, n, t1);
    for (int i = 0; i < 3; i++) t1[i] = ii[i] * t1[i];
    FN(cross3)(t1, rel, t2);
    REAL denom = inv_m + FN(dot3)(n, t2);
    REAL J = ((-one * (one + S->elasticity)) * nv + bv) / denom;
    REAL Jn[3] = {J * n[0], J * n[1], J * n[2]};
    REAL pn_v[3], pn_a[3], pd_v[3], pd_a[3], vd[3], Jdv[3];
    FN(cross3)(rel, Jn, pn_a);
    for (int i = 0; i < 3; i++) { pn_v[i] = Jn[i] / mass; pn_a[i] = ii[i] * pn_a[i]; vd[i] = cvel[i] - nv * n[i]; }
    REAL nd = SQRT((vd[0] * vd[0] + vd[1] * vd[1]) + vd[2] * vd[2]);
    REAL a = nd / denom, f = S->friction * J;
    REAL Jd = a < f ? a : f;   /* np.minimum (no NaNs on this path) */
    REAL dd = (REAL)1e-6 + nd;
    for (int i = 0; i < 3; i++) Jdv[i] = -Jd * (vd[i] / dd);
    FN(cross3)(rel, Jdv, pd_a);
    for (int i = 0; i < 3; i++) { pd_v[i] = Jdv[i] / mass; pd_a[i] = ii[i] * pd_a[i]; }
    /* apply_n = (pen > 0) & (nv < 0) & (J > 0); apply_d = apply_n & (nd > 0.01): each predicate through branch_pred
     * (margin bookkeeping of brax_v1.System._note_margin: a later predicate only counts while the earlier ones hold) */
    const int c0 = wall ? 4 : 0;
    int live = branch_pred(pen > zero, fabs((double)pen) * 100.0, c0, bc, mg);
    if (live) live = branch_pred(nv < zero, fabs((double)nv), c0 + 1, bc, mg);
    if (live) live = branch_pred(J > zero, fabs((double)J), c0 + 2, bc, mg);
    int drag = live ? branch_pred(nd > (REAL)0.01, fabs((double)nd - 0.01), c0 + 3, bc, mg) : 0;
    REAL an = live ? one : zero;
    REAL ad = an * (drag ? one : zero);
    for (int i = 0; i < 3; i++) { dvel[i] = pn_v[i] * an + pd_v[i] * ad; dang[i] = pn_a[i] * an + pd_a[i] * ad; }
}

/* g(t) of brax_v1.System._closest_segment_box */
static inline REAL FN(seg_g)(const REAL *a, const REAL *d, const REAL *lo, const REAL *hi, REAL t) {
    REAL e[3];
    for (int i = 0; i < 3; i++) {
        REAL p = a[i] + t * d[i];
        REAL c = p < lo[i] ? lo[i] : (p > hi[i] ? hi[i] : p);
        e[i] = p - c;
    }
    return FN(dot3)(e, d);
}

static void FN(closest_segment_box)(const REAL *a, const REAL *b, const REAL *lo, const REAL *hi, REAL *sp, REAL *bp) {
    REAL d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    REAL g0 = FN(seg_g)(a, d, lo, hi, 0), g1 = FN(seg_g)(a, d, lo, hi, 1), t;
    if (g0 >= 0) t = 0;
    else if (g1 <= 0) t = 1;
    else {
        REAL tl = 0, tr = 1, gl = g0, gr = g1;
        for (int it = 0; it < 16; it++) {
            REAL tm = (REAL)0.5 * (tl + tr), gm = FN(seg_g)(a, d, lo, hi, tm);
            if (gm > 0) { tr = tm; gr = gm; } else { tl = tm; gl = gm; }
        }
        REAL den = gr - gl;
        t = den > 0 ? tl - gl * (tr - tl) / den : tl;
    }
    for (int i = 0; i < 3; i++) {
        REAL p = a[i] + t * d[i];
        sp[i] = p;
        bp[i] = p < lo[i] ? lo[i] : (p > hi[i] ? hi[i] : p);
    }
}

/* brax_v1.System._contacts on one env: cv/ca [nb][3] = ground group + wall group */
static void FN(contacts)(const FN(SysDesc) *S, const REAL *pos, const REAL *rot, const REAL *vel, const REAL *ang,
                         REAL *cv, REAL *ca, double *mg, BranchCtl *bc) {
    const int nb = S->nb;
    REAL cnt[MAXB], sv[MAXB][3], sa[MAXB][3];
    for (int b = 0; b < nb; b++) for (int i = 0; i < 3; i++) { cv[3 * b + i] = 0; ca[3 * b + i] = 0; }
    /* ---- capsule ends vs plane (_ground_contacts) */
    if (S->ground >= 0 && S->ncp > 0) {
        for (int b = 0; b < nb; b++) { cnt[b] = 0; for (int i = 0; i < 3; i++) sv[b][i] = sa[b][i] = 0; }
        const int g = S->ground;
        const REAL up[3] = {0, 0, 1}; REAL n[3];
        FN(rotate)(up, rot + 4 * g, n);
        for (int k = 0; k < S->ncp; k++) {
            int b = S->cp_body[k];
            const REAL *p = pos + 3 * b;
            REAL e[3], cpos[3], r[3], cvel[3], w[3], gp[3], dv[3], da[3];
            FN(rotate)(S->cp_end + 3 * k, rot + 4 * b, e);
            for (int i = 0; i < 3; i++) { cpos[i] = (p[i] + e[i]) - n[i] * S->cp_rad[k]; r[i] = cpos[i] - p[i]; }
            FN(cross3)(ang + 3 * b, r, w);
            for (int i = 0; i < 3; i++) { cvel[i] = vel[3 * b + i] + w[i]; gp[i] = pos[3 * g + i] - cpos[i]; }
            REAL pen = FN(dot3)(gp, n);
            FN(impulse)(S, b, p, cpos, cvel, n, pen, dv, da, mg, bc, 0);
            cnt[b] += (dv[0] != 0 || dv[1] != 0 || dv[2] != 0) ? (REAL)1 : (REAL)0;
            for (int i = 0; i < 3; i++) { sv[b][i] += dv[i]; sa[b][i] += da[i]; }
        }
        for (int b = 0; b < nb; b++) {
            REAL d = (REAL)1e-8 + cnt[b];
            for (int i = 0; i < 3; i++) { cv[3 * b + i] = sv[b][i] / d; ca[3 * b + i] = sa[b][i] / d; }
        }
    }
    /* ---- capsules vs Arena boxes (_wall_contacts) */
    if (S->nbox > 0 && S->ncap > 0) {
        for (int b = 0; b < nb; b++) { cnt[b] = 0; for (int i = 0; i < 3; i++) sv[b][i] = sa[b][i] = 0; }
        const REAL *ap = pos + 3 * S->arena;
        int any = 0;
        for (int c = 0; c < S->ncap; c++) {
            int b = S->cap_body[c];
            const REAL *p = pos + 3 * b;
            REAL aw[3], bw[3], mn[3], mx[3], rad = S->cap_rad[c];
            FN(rotate)(S->cap_a + 3 * c, rot + 4 * b, aw);
            FN(rotate)(S->cap_b + 3 * c, rot + 4 * b, bw);
            const REAL slack = rad + (REAL)1e-3;   /* cull: exact (dist > rad => zero impulse), with slack for rounding */
            for (int i = 0; i < 3; i++) {
                aw[i] = p[i] + aw[i]; bw[i] = p[i] + bw[i];
                mn[i] = (aw[i] < bw[i] ? aw[i] : bw[i]) - slack; mx[i] = (aw[i] > bw[i] ? aw[i] : bw[i]) + slack;
            }
            for (int x = 0; x < S->nbox; x++) {
                REAL lo[3], hi[3];
                int out = 0;
                for (int i = 0; i < 3; i++) {
                    lo[i] = ap[i] + S->boxes[6 * x + i]; hi[i] = ap[i] + S->boxes[6 * x + 3 + i];
                    out |= (mx[i] < lo[i]) | (mn[i] > hi[i]);
                }
                if (out) continue;
                REAL sp[3], bp[3], dvec[3], n[3], r[3], w[3], cvel[3], dv[3], da[3];
                FN(closest_segment_box)(aw, bw, lo, hi, sp, bp);
                for (int i = 0; i < 3; i++) dvec[i] = sp[i] - bp[i];
                REAL dist = SQRT((dvec[0] * dvec[0] + dvec[1] * dvec[1]) + dvec[2] * dvec[2]);
                REAL dd = (REAL)1e-6 + dist;
                for (int i = 0; i < 3; i++) { n[i] = dvec[i] / dd; r[i] = bp[i] - p[i]; }
                REAL pen = rad - dist;
                FN(cross3)(ang + 3 * b, r, w);
                for (int i = 0; i < 3; i++) cvel[i] = vel[3 * b + i] + w[i];
                /* a segment point inside the box: dvec = 0 exactly => n = 0, nv = 0 and no impulse in any evaluation
                 * order: not rounding-ambiguous, kept out of the margin / branch bookkeeping */
                const int degenerate = dist == (REAL)0;
                if (mg) {   /* ... but AT the surface the normal swings from 0 to unit length within ~1e-5 m: a closest
                             * point that close to a box face is rounding-ambiguous although no predicate switches
                             * (brax_v1._wall_contacts `surface`) */
                    double depth = 1e9, sf;
                    for (int i = 0; i < 3; i++) depth = fmin(depth, fmin((double)(sp[i] - lo[i]), (double)(hi[i] - sp[i])));
                    sf = dist > 0 ? (double)dist * 10.0 : (depth >= 0 ? depth * 100.0 : 1e9);
                    *mg = fmin(*mg, sf);
                    if (bc && sf < bc->cause[4]) bc->cause[4] = sf;
                }
                FN(impulse)(S, b, p, bp, cvel, n, pen, dv, da, degenerate ? 0 : mg, degenerate ? 0 : bc, 1);
                cnt[b] += (dv[0] != 0 || dv[1] != 0 || dv[2] != 0) ? (REAL)1 : (REAL)0;
                for (int i = 0; i < 3; i++) { sv[b][i] += dv[i]; sa[b][i] += da[i]; }
                any = 1;
            }
        }
        if (any) for (int b = 0; b < nb; b++) {
            REAL d = (REAL)1e-8 + cnt[b];
            for (int i = 0; i < 3; i++) { cv[3 * b + i] = cv[3 * b + i] + sv[b][i] / d; ca[3 * b + i] = ca[3 * b + i] + sa[b][i] / d; }
        }
    }
}

/* brax_v1.System.substep on one env, in place; cvel/cang = this substep's contact impulses */
static void FN(substep)(const FN(SysDesc) *S, REAL *pos, REAL *rot, REAL *vel, REAL *ang, const REAL *act,
                        REAL *cvel, REAL *cang, double *mg, BranchCtl *bc) {
    const int nb = S->nb, nj = S->nj;
    const REAL h = S->h;
    /* kinetic */
    for (int b = 0; b < nb; b++) {
        REAL m = S->active[b];
        for (int i = 0; i < 3; i++) pos[3 * b + i] = pos[3 * b + i] + vel[3 * b + i] * h * m;
        REAL raq[4] = {(REAL)0 * (REAL)0.5 * h, ang[3 * b] * m * (REAL)0.5 * h, ang[3 * b + 1] * m * (REAL)0.5 * h,
                       ang[3 * b + 2] * m * (REAL)0.5 * h};
        REAL dq[4], *q = rot + 4 * b;
        FN(quat_mul)(raq, q, dq);
        for (int i = 0; i < 4; i++) q[i] = q[i] + dq[i];
        REAL nn = SQRT(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        for (int i = 0; i < 4; i++) q[i] = q[i] / nn;
    }
    /* joints + actuators (_joints_and_actuators) */
    REAL dvel[MAXB][3], dang[MAXB][3], adang[MAXB][3];
    REAL jdv_p[MAXJ][3], jda_p[MAXJ][3], jdv_c[MAXJ][3], jda_c[MAXJ][3], axis_p[MAXJ][3], psi[MAXJ];
    int act_out[MAXJ];
    for (int b = 0; b < nb; b++) for (int i = 0; i < 3; i++) dvel[b][i] = dang[b][i] = adang[b][i] = 0;
    for (int j = 0; j < nj; j++) {
        const int P = S->j_parent[j], C = S->j_child[j];
        const REAL *pp = pos + 3 * P, *pc = pos + 3 * C, *qp_ = rot + 4 * P, *qc = rot + 4 * C;
        const REAL *vp = vel + 3 * P, *vc = vel + 3 * C, *ap = ang + 3 * P, *ac = ang + 3 * C;
        const REAL *iip = S->inv_inertia + 3 * P, *iic = S->inv_inertia + 3 * C;
        REAL rp[3], rc[3], wp[3], wc[3], xp[3], xc[3], F[3], nF[3], lp[3], lc[3], t1[3], t2[3];
        FN(rotate)(S->j_off_p + 3 * j, qp_, rp);
        FN(rotate)(S->j_off_c + 3 * j, qc, rc);
        FN(cross3)(ap, rp, xp);
        FN(cross3)(ac, rc, xc);
        for (int i = 0; i < 3; i++) {
            wp[i] = pp[i] + rp[i]; wc[i] = pc[i] + rc[i];
            REAL wvp = vp[i] + xp[i], wvc = vc[i] + xc[i];
            F[i] = (wp[i] - wc[i]) * S->j_stiff[j] + S->j_sdamp[j] * (wvp - wvc);
            nF[i] = -F[i];
            lp[i] = wp[i] - pp[i]; lc[i] = wc[i] - pc[i];
        }
        FN(cross3)(lp, nF, t1);
        FN(cross3)(lc, F, t2);
        REAL ax_c[3], ref_p[3], ref_c[3], cr[3], tq[3];
        FN(rotate)(S->j_axis + 9 * j, qp_, axis_p[j]);
        FN(rotate)(S->j_axis + 9 * j + 6, qp_, ref_p);
        FN(rotate)(S->j_axis + 9 * j + 6, qc, ref_c);
        FN(cross3)(ref_p, ref_c, cr);
        psi[j] = ATAN2(FN(dot3)(cr, axis_p[j]), FN(dot3)(ref_p, ref_c));
        FN(rotate)(S->j_axis + 9 * j, qc, ax_c);
        FN(cross3)(axis_p[j], ax_c, tq);
        REAL lo = S->j_limit[2 * j], hi = S->j_limit[2 * j + 1];
        REAL da = psi[j] < lo ? lo - psi[j] : (REAL)0;
        if (psi[j] > hi) da = hi - psi[j];
        /* the actuator switches off discontinuously outside the joint limits (brax_v1._joints_and_actuators); the
         * limit torque itself (da) is continuous through 0 */
        {
            double dl = fabs((double)(REAL)(psi[j] - lo)), dh = fabs((double)(REAL)(psi[j] - hi));
            act_out[j] = branch_pred(psi[j] < lo || psi[j] > hi, (dl < dh ? dl : dh) * 100.0, 8, bc, mg);
        }
        for (int i = 0; i < 3; i++) {
            REAL t = S->j_stiff[j] * tq[i];
            t = t - S->j_lstr[j] * axis_p[j][i] * da;
            t = t - S->j_adamp[j] * (ap[i] - ac[i]);
            jdv_p[j][i] = nF[i] / S->mass[P];
            jda_p[j][i] = iip[i] * t1[i] + iip[i] * t;
            jdv_c[j][i] = F[i] / S->mass[C];
            jda_c[j][i] = iic[i] * t2[i] + iic[i] * (-t);
        }
    }
    for (int j = 0; j < nj; j++) for (int i = 0; i < 3; i++) { dvel[S->j_parent[j]][i] += jdv_p[j][i]; dang[S->j_parent[j]][i] += jda_p[j][i]; }
    for (int j = 0; j < nj; j++) for (int i = 0; i < 3; i++) { dvel[S->j_child[j]][i] += jdv_c[j][i]; dang[S->j_child[j]][i] += jda_c[j][i]; }
    REAL atau[MAXJ][3];
    for (int k = 0; k < S->na; k++) {
        int j = S->a_joint[k];
        REAL t = act[k] * S->a_strength[k];
        if (act_out[j]) t = 0;
        for (int i = 0; i < 3; i++) atau[k][i] = -(axis_p[j][i] * t);
    }
    for (int k = 0; k < S->na; k++) { int P = S->j_parent[S->a_joint[k]];
        for (int i = 0; i < 3; i++) adang[P][i] += S->inv_inertia[3 * P + i] * atau[k][i]; }
    for (int k = 0; k < S->na; k++) { int C = S->j_child[S->a_joint[k]];
        for (int i = 0; i < 3; i++) adang[C][i] += S->inv_inertia[3 * C + i] * (-atau[k][i]); }
    /* potential */
    for (int b = 0; b < nb; b++) {
        REAL m = S->active[b];
        for (int i = 0; i < 3; i++) {
            REAL v = S->vel_damp * vel[3 * b + i];
            vel[3 * b + i] = (v + (dvel[b][i] + S->gravity[i]) * h) * m;
            REAL a = S->ang_damp * ang[3 * b + i];
            ang[3 * b + i] = (a + (dang[b][i] + adang[b][i]) * h) * m;
        }
    }
    /* contacts -> collision */
    FN(contacts)(S, pos, rot, vel, ang, cvel, cang, mg, bc);
    for (int b = 0; b < nb; b++) {
        REAL m = S->active[b];
        for (int i = 0; i < 3; i++) {
            vel[3 * b + i] = (vel[3 * b + i] + cvel[3 * b + i]) * m;
            ang[3 * b + i] = (ang[3 * b + i] + cang[3 * b + i]) * m;
        }
    }
}

/* System.step on N envs in place (QP arrays [N][nb][3|4], act [N][na]); cv/ca [N][nb][3] = summed contact impulses;
 * margin [N] (may be NULL) = the env's distance from its nearest discontinuous branch in this step (_note_margin).
 * Two-branch test aid (all may be NULL): flip_mask [N] = which of the env's marginal predicate evaluations (those
 * closer than flip_thr, in evaluation order) to invert; n_marginal [N] <- how many there were; cause [N][BRAX_NCAUSE]
 * <- smallest distance per predicate class.
 * threads <= 0: OpenMP default. Returns 0, or -1 if the system exceeds the static limits. */
int FN(brax_step)(const FN(SysDesc) *S, long N, REAL *pos, REAL *rot, REAL *vel, REAL *ang, const REAL *act,
                  REAL *cv, REAL *ca, double *margin, int threads, const unsigned *flip_mask, double flip_thr,
                  int *n_marginal, double *cause) {
    if (S->nb > MAXB || S->nj > MAXJ || S->na > MAXJ) return -1;
    const int nb = S->nb;
    const int ctl = flip_mask || n_marginal || cause;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (long e = 0; e < N; e++) {
        REAL dv[3 * MAXB], da[3 * MAXB];
        REAL *p = pos + e * nb * 3, *q = rot + e * nb * 4, *v = vel + e * nb * 3, *w = ang + e * nb * 3;
        REAL *ocv = cv + e * nb * 3, *oca = ca + e * nb * 3;
        for (int i = 0; i < 3 * nb; i++) { ocv[i] = 0; oca[i] = 0; }
        double *mg = margin ? margin + e : 0;
        if (mg) *mg = 1e9;
        BranchCtl bcs, *bc = ctl ? &bcs : 0;
        if (bc) {
            bc->thr = flip_thr; bc->mask = flip_mask ? flip_mask[e] : 0u; bc->count = 0;
            for (int i = 0; i < BRAX_NCAUSE; i++) bc->cause[i] = 1e9;
        }
        for (int s = 0; s < S->substeps; s++) {
            FN(substep)(S, p, q, v, w, act + e * S->na, dv, da, mg, bc);
            for (int i = 0; i < 3 * nb; i++) { ocv[i] = ocv[i] + dv[i]; oca[i] = oca[i] + da[i]; }
        }
        if (n_marginal) n_marginal[e] = bc->count;
        if (cause) for (int i = 0; i < BRAX_NCAUSE; i++) cause[e * BRAX_NCAUSE + i] = bc->cause[i];
    }
    return 0;
}

/* System.info: one collider evaluation (reset observations). */
int FN(brax_info)(const FN(SysDesc) *S, long N, const REAL *pos, const REAL *rot, const REAL *vel, const REAL *ang,
                  REAL *cv, REAL *ca, int threads) {
    if (S->nb > MAXB) return -1;
    const int nb = S->nb;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (long e = 0; e < N; e++)
        FN(contacts)(S, pos + e * nb * 3, rot + e * nb * 4, vel + e * nb * 3, ang + e * nb * 3, cv + e * nb * 3, ca + e * nb * 3, 0, 0);
    return 0;
}

#undef MAXB
#undef MAXJ
#undef FN
#undef CAT
#undef CAT_
