/* kgo_physics.c -- fp64 restatement of mj_step for the j2s7s300 end-effector + object model.
 * TEST INFRASTRUCTURE (oracle).  Follows the call site kinova_gripper_env.py:1535 (sim.step())
 * and the published MuJoCo pipeline, stage by stage:
 *   P1 kinematics/com  P2 CRBA  P3 RNE bias + passive + actuation  P4 broadphase
 *   P5 narrowphase (plane-hull, hull-hull MPR as in libccd)  P6 constraint rows
 *   P7 Newton solver   P8 sensors (kgo_rays.c)   P9 semi-implicit Euler with implicit damping.
 * Deliberately written in the most literal dense form (6x6 spatial matrices, dense 15x15
 * Cholesky, exhaustive hull support, sorted exact line search): the CUDA kernels use different,
 * structure-exploiting formulations so that agreement between the two means something.
 */
#include "kgo_internal.h"

/* ------------------------------------------------------------------ quaternions */
void kgo_quat2mat(double *M, const double *q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
    M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
    M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
void kgo_mulquat(double *r, const double *a, const double *b) {
    double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                   a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                   a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
    memcpy(r, t, sizeof t);
}
static void normquat(double *q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < 1e-15) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    for (int i = 0; i < 4; i++) q[i] /= n;
}
/* MuJoCo default eulerseq "xyz", intrinsic: q = qx qy qz (SURVEY Appendix A "Euler angles") */
void kgo_euler2quat(double *q, const double *e) {
    q[0] = 1; q[1] = q[2] = q[3] = 0;
    for (int i = 0; i < 3; i++) {
        double r[4] = {cos(0.5 * e[i]), 0, 0, 0};
        r[i + 1] = sin(0.5 * e[i]);
        kgo_mulquat(q, q, r);
    }
    normquat(q);
}
void kgo_mulmm3(double *r, const double *A, const double *B) {
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(r, t, sizeof t);
}

/* ------------------------------------------------------------------ state */
void kgo_state_init(const KgoModel *m, KgoState *s, int shape, const double hand_euler[3], const double obj_pos[3]) {
    (void)m;
    memset(s, 0, sizeof *s);
    s->shape = shape;
    kgo_euler2quat(s->hand_quat, hand_euler);
    s->qpos[9] = obj_pos[0]; s->qpos[10] = obj_pos[1]; s->qpos[11] = obj_pos[2];
    s->qpos[12] = 1.0;
}

/* ------------------------------------------------------------------ P1 kinematics */
typedef struct {
    double cdof[KGO_NV][6]; /* [ang; lin] of each dof's motion axis, expressed at the world origin */
    double xquat[KGO_NBODY][4];
} Kin;

static void kinematics(const KgoModel *m, const KgoState *s, KgoData *d, Kin *k) {
    const double *q = s->qpos;
    /* world */
    memset(d->xpos, 0, sizeof d->xpos);
    for (int b = 0; b < KGO_NBODY; b++) { k->xquat[b][0] = 1; k->xquat[b][1] = k->xquat[b][2] = k->xquat[b][3] = 0; }
    /* link_7: fixed orientation (runtime input), three slides along body-frame axes */
    {
        double R[9];
        memcpy(k->xquat[1], s->hand_quat, 4 * sizeof(double));
        kgo_quat2mat(R, k->xquat[1]);
        copy3(d->xpos + 3, m->body_pos[1]);
        for (int j = 0; j < 3; j++) {
            double ax[3];
            mulmv3(ax, R, m->slide_axes[j]);
            addscl3(d->xpos + 3, d->xpos + 3, ax, q[j]);
            k->cdof[j][0] = k->cdof[j][1] = k->cdof[j][2] = 0;
            copy3(k->cdof[j] + 3, ax);
        }
    }
    /* fingers and tips: hinge about local z at the body origin */
    for (int b = 2; b <= 7; b++) {
        int p = m->body_parent[b];
        double Rp[9], off[3], qj[4], R[9], ax[3], loc[3] = {0, 0, 1};
        kgo_quat2mat(Rp, k->xquat[p]);
        mulmv3(off, Rp, m->body_pos[b]);
        add3(d->xpos + 3 * b, d->xpos + 3 * p, off);
        kgo_mulquat(k->xquat[b], k->xquat[p], m->body_quat[b]);
        double ang = q[1 + b]; /* body 2 -> qpos 3, ... body 7 -> qpos 8 */
        qj[0] = cos(0.5 * ang); qj[1] = 0; qj[2] = 0; qj[3] = sin(0.5 * ang);
        kgo_mulquat(k->xquat[b], k->xquat[b], qj);
        normquat(k->xquat[b]);
        kgo_quat2mat(R, k->xquat[b]);
        mulmv3(ax, R, loc);
        int dof = 1 + b;
        copy3(k->cdof[dof], ax);
        cross3(k->cdof[dof] + 3, d->xpos + 3 * b, ax);
    }
    /* object: free joint */
    {
        double R[9];
        copy3(d->xpos + 24, q + 9);
        memcpy(k->xquat[8], q + 12, 4 * sizeof(double));
        normquat(k->xquat[8]);
        kgo_quat2mat(R, k->xquat[8]);
        for (int i = 0; i < 3; i++) {
            memset(k->cdof[9 + i], 0, 6 * sizeof(double));
            k->cdof[9 + i][3 + i] = 1;
            double ax[3] = {R[i], R[3 + i], R[6 + i]};
            copy3(k->cdof[12 + i], ax);
            cross3(k->cdof[12 + i] + 3, d->xpos + 24, ax);
        }
    }
    /* frames of bodies, inertial frames, geoms, sites */
    for (int b = 0; b < KGO_NBODY; b++) {
        double *R = d->xmat + 9 * b, off[3], qi[4];
        kgo_quat2mat(R, k->xquat[b]);
        const double *ipos = m->body_ipos[b], *iquat = m->body_iquat[b];
        if (b == 8) { ipos = m->obj_gpos[s->shape]; iquat = m->obj_gquat[s->shape]; }
        mulmv3(off, R, ipos);
        add3(d->xipos + 3 * b, d->xpos + 3 * b, off);
        kgo_mulquat(qi, k->xquat[b], iquat);
        kgo_quat2mat(d->ximat + 9 * b, qi);
    }
    for (int g = 0; g < KGO_NGEOM; g++) {
        int b = m->geom_body[g];
        const double *gp = m->geom_pos[g], *gq = m->geom_quat[g];
        if (g == 8) { gp = m->obj_gpos[s->shape]; gq = m->obj_gquat[s->shape]; }
        double off[3], qg[4];
        mulmv3(off, d->xmat + 9 * b, gp);
        add3(d->geom_xpos + 3 * g, d->xpos + 3 * b, off);
        kgo_mulquat(qg, k->xquat[b], gq);
        kgo_quat2mat(d->geom_xmat + 9 * g, qg);
    }
    for (int i = 0; i < KGO_NSITE; i++) {
        int b = 1 + m->site_body[i]; /* blob: 0 link_7, 1..3 proximal k, 4..6 tip k -> body ids */
        if (m->site_body[i] >= 1 && m->site_body[i] <= 3) b = 2 * m->site_body[i];
        else if (m->site_body[i] >= 4) b = 2 * (m->site_body[i] - 3) + 1;
        double off[3];
        mulmv3(off, d->xmat + 9 * b, m->site_pos[i]);
        add3(d->site_xpos + 3 * i, d->xpos + 3 * b, off);
        mulmv3(d->site_zaxis + 3 * i, d->xmat + 9 * b, m->site_z[i]);
    }
}

/* ------------------------------------------------------------------ spatial algebra (origin based) */
static void body_inertia6(double I6[36], double mass, const double *c, const double *Rc, const double *diag) {
    double Ic[9] = {0};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) Ic[3 * i + j] += Rc[3 * i + k] * diag[k] * Rc[3 * j + k];
    double cc = dot3(c, c);
    double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
    memset(I6, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        I6[6 * i + j] = Ic[3 * i + j] + mass * ((i == j ? cc : 0) - c[i] * c[j]);
        I6[6 * i + 3 + j] = mass * cx[3 * i + j];
        I6[6 * (3 + i) + j] = -mass * cx[3 * i + j];
        I6[6 * (3 + i) + 3 + j] = (i == j) ? mass : 0;
    }
}
static void mul6(double *r, const double *I6, const double *v) {
    double t[6];
    for (int i = 0; i < 6; i++) { t[i] = 0; for (int j = 0; j < 6; j++) t[i] += I6[6 * i + j] * v[j]; }
    memcpy(r, t, sizeof t);
}
static void cross_motion(double *r, const double *v, const double *s) {
    double a[3], b[3], c[3];
    cross3(a, v, s); cross3(b, v, s + 3); cross3(c, v + 3, s);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double *r, const double *v, const double *f) {
    double a[3], b[3], c[3];
    cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
    r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* ------------------------------------------------------------------ P2 CRBA, P3 RNE */
static void crba_rne(const KgoModel *m, const KgoState *s, KgoData *d, const Kin *k) {
    double I6[KGO_NBODY][36], crb[KGO_NBODY][36];
    memset(I6, 0, sizeof I6);
    for (int b = 1; b < KGO_NBODY; b++) {
        const double *diag = (b == 8) ? m->obj_inertia[s->shape] : m->body_inertia[b];
        body_inertia6(I6[b], m->body_mass[b], d->xipos + 3 * b, d->ximat + 9 * b, diag);
    }
    memcpy(crb, I6, sizeof I6);
    for (int b = KGO_NBODY - 1; b >= 1; b--) {
        int p = m->body_parent[b];
        if (p > 0) for (int i = 0; i < 36; i++) crb[p][i] += crb[b][i];
    }
    memset(d->M, 0, sizeof d->M);
    for (int i = 0; i < KGO_NV; i++) {
        double f[6];
        mul6(f, crb[m->dof_body[i]], k->cdof[i]);
        for (int j = i; j >= 0; j = m->dof_parent[j]) {
            double v = 0;
            for (int c = 0; c < 6; c++) v += k->cdof[j][c] * f[c];
            d->M[KGO_NV * i + j] = d->M[KGO_NV * j + i] = v;
        }
        d->M[KGO_NV * i + i] += m->armature;
    }
    /* RNE with zero acceleration: bias forces */
    double vel[KGO_NBODY][6], acc[KGO_NBODY][6], frc[KGO_NBODY][6];
    memset(vel, 0, sizeof vel); memset(acc, 0, sizeof acc);
    acc[0][5] = m->gravity; /* -gravity, gravity = (0,0,-9.81) */
    for (int b = 1; b < KGO_NBODY; b++) {
        int p = m->body_parent[b];
        memcpy(vel[b], vel[p], sizeof vel[b]);
        memcpy(acc[b], acc[p], sizeof acc[b]);
        if (b == 8) {
            /* free joint: translations have zero cdof_dot; the three rotations all use the
             * velocity accumulated before any of them (MuJoCo's ball-joint rule) */
            for (int j = 9; j < 12; j++) for (int c = 0; c < 6; c++) vel[b][c] += k->cdof[j][c] * s->qvel[j];
            double dd[3][6];
            for (int j = 0; j < 3; j++) cross_motion(dd[j], vel[b], k->cdof[12 + j]);
            for (int j = 0; j < 3; j++) for (int c = 0; c < 6; c++) {
                acc[b][c] += dd[j][c] * s->qvel[12 + j];
                vel[b][c] += k->cdof[12 + j][c] * s->qvel[12 + j];
            }
        } else {
            for (int j = 0; j < KGO_NV; j++) if (m->dof_body[j] == b) {
                double dd[6];
                cross_motion(dd, vel[b], k->cdof[j]);
                for (int c = 0; c < 6; c++) { acc[b][c] += dd[c] * s->qvel[j]; vel[b][c] += k->cdof[j][c] * s->qvel[j]; }
            }
        }
        double Iv[6], Ia[6], vxIv[6];
        mul6(Iv, I6[b], vel[b]); mul6(Ia, I6[b], acc[b]);
        cross_force(vxIv, vel[b], Iv);
        for (int c = 0; c < 6; c++) frc[b][c] = Ia[c] + vxIv[c];
    }
    for (int b = KGO_NBODY - 1; b >= 1; b--) {
        int p = m->body_parent[b];
        if (p > 0) for (int c = 0; c < 6; c++) frc[p][c] += frc[b][c];
    }
    for (int j = 0; j < KGO_NV; j++) {
        double v = 0;
        for (int c = 0; c < 6; c++) v += k->cdof[j][c] * frc[m->dof_body[j]][c];
        d->qfrc_bias[j] = v;
    }
    /* passive: joint damping on every dof (XML:41); actuation (XML:214-222) */
    for (int j = 0; j < KGO_NV; j++) d->qfrc_passive[j] = -m->damping * s->qvel[j];
}

static void actuation(const KgoModel *m, const KgoState *s, const double *ctrl, KgoData *d) {
    static const int act_dof[KGO_NU] = {0, 0, 1, 1, 2, 2, 3, 5, 7};
    memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
    for (int a = 0; a < KGO_NU; a++) {
        double c = ctrl[a], f;
        int j = act_dof[a];
        if (m->act_kv[a] > 0) {
            if (c < m->act_range[a][0]) c = m->act_range[a][0];
            if (c > m->act_range[a][1]) c = m->act_range[a][1];
            f = m->act_kv[a] * (c - s->qvel[j]);
        } else {
            f = c;
        }
        d->qfrc_actuator[j] += m->act_gear[a] * f;
    }
}

/* ------------------------------------------------------------------ dense Cholesky */
static int chol(double *A, int n) { /* in place, lower */
    for (int j = 0; j < n; j++) {
        double s = A[n * j + j];
        for (int k = 0; k < j; k++) s -= A[n * j + k] * A[n * j + k];
        if (s <= 1e-300) return -1;
        s = sqrt(s);
        A[n * j + j] = s;
        for (int i = j + 1; i < n; i++) {
            double t = A[n * i + j];
            for (int k = 0; k < j; k++) t -= A[n * i + k] * A[n * j + k];
            A[n * i + j] = t / s;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, double *x) {
    for (int i = 0; i < n; i++) { double t = x[i]; for (int k = 0; k < i; k++) t -= L[n * i + k] * x[k]; x[i] = t / L[n * i + i]; }
    for (int i = n - 1; i >= 0; i--) { double t = x[i]; for (int k = i + 1; k < n; k++) t -= L[n * k + i] * x[k]; x[i] = t / L[n * i + i]; }
}

/* ------------------------------------------------------------------ P5 support + MPR (libccd's algorithm) */
typedef struct {
    const double *vert; int nvert; const double *pos, *mat; double margin;
} Hull;

static void hull_support(const Hull *h, const double *dir, double *out) {
    double ld[3];
    multmv3(ld, h->mat, dir);
    int best = 0; double bv = -1e300;
    for (int i = 0; i < h->nvert; i++) {
        double v = dot3(h->vert + 3 * i, ld);
        if (v > bv) { bv = v; best = i; }
    }
    double w[3];
    mulmv3(w, h->mat, h->vert + 3 * best);
    add3(out, w, h->pos);
    addscl3(out, out, dir, h->margin);
}

typedef struct { double v[3], v1[3], v2[3]; } Supp;
#define CCD_EPS 2.220446049250313e-16
static int is_zero(double x) { return fabs(x) < CCD_EPS; }
static int ccd_eq(double a, double b) {
    double ab = fabs(a - b);
    if (ab < CCD_EPS) return 1;
    double fa = fabs(a), fb = fabs(b);
    return fb > fa ? ab < CCD_EPS * fb : ab < CCD_EPS * fa;
}
static void mink_support(const Hull *h1, const Hull *h2, const double *dir, Supp *s) {
    double nd[3] = {-dir[0], -dir[1], -dir[2]};
    hull_support(h1, dir, s->v1);
    hull_support(h2, nd, s->v2);
    sub3(s->v, s->v1, s->v2);
}
static void portal_dir(const Supp *p, double *dir) {
    double a[3], b[3];
    sub3(a, p[2].v, p[1].v); sub3(b, p[3].v, p[1].v);
    cross3(dir, a, b);
    normalize3(dir);
}
static int reach_tolerance(const Supp *p, const Supp *v4, const double *dir, double tol) {
    double dv1 = dot3(p[1].v, dir), dv2 = dot3(p[2].v, dir), dv3 = dot3(p[3].v, dir), dv4 = dot3(v4->v, dir);
    double d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
    double mn = d1 < d2 ? d1 : d2; mn = mn < d3 ? mn : d3;
    return ccd_eq(mn, tol) || mn < tol;
}
static void expand_portal(Supp *p, const Supp *v4) {
    double v4v0[3];
    cross3(v4v0, v4->v, p[0].v);
    if (dot3(p[1].v, v4v0) > 0) {
        if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4;
    } else {
        if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4;
    }
}
static double point_seg_dist2(const double *P, const double *a, const double *b, double *w) {
    double d[3], ap[3];
    sub3(d, b, a); sub3(ap, a, P);
    double t = -dot3(ap, d) / dot3(d, d);
    if (t < 0 || is_zero(t)) copy3(w, a);
    else if (t > 1 || ccd_eq(t, 1)) copy3(w, b);
    else addscl3(w, a, d, t);
    double r[3]; sub3(r, w, P);
    return dot3(r, r);
}
static double point_tri_dist2(const double *P, const double *x0, const double *B, const double *C, double *w) {
    double d1[3], d2[3], a[3];
    sub3(d1, B, x0); sub3(d2, C, x0); sub3(a, x0, P);
    double v = dot3(d1, d1), ww = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
    double div = ww * v - r * r, s, t = 0;
    if (is_zero(div)) s = -1; else { s = (q * r - ww * p) / div; t = (-s * r - q) / ww; }
    if ((is_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (is_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) && (ccd_eq(t + s, 1) || t + s < 1)) {
        addscl3(w, x0, d1, s); addscl3(w, w, d2, t);
        double rr[3]; sub3(rr, w, P);
        return dot3(rr, rr);
    }
    double w2[3], dist = point_seg_dist2(P, x0, B, w), d2v;
    d2v = point_seg_dist2(P, x0, C, w2); if (d2v < dist) { dist = d2v; copy3(w, w2); }
    d2v = point_seg_dist2(P, B, C, w2); if (d2v < dist) { dist = d2v; copy3(w, w2); }
    return dist;
}
static void find_pos(const Supp *p, double *pos) {
    double dir[3], b[4], t[3], sum;
    portal_dir(p, dir);
    cross3(t, p[1].v, p[2].v); b[0] = dot3(t, p[3].v);
    cross3(t, p[3].v, p[2].v); b[1] = dot3(t, p[0].v);
    cross3(t, p[0].v, p[1].v); b[2] = dot3(t, p[3].v);
    cross3(t, p[2].v, p[1].v); b[3] = dot3(t, p[0].v);
    sum = b[0] + b[1] + b[2] + b[3];
    if (is_zero(sum) || sum < 0) {
        b[0] = 0;
        cross3(t, p[2].v, p[3].v); b[1] = dot3(t, dir);
        cross3(t, p[3].v, p[1].v); b[2] = dot3(t, dir);
        cross3(t, p[1].v, p[2].v); b[3] = dot3(t, dir);
        sum = b[1] + b[2] + b[3];
    }
    double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
    for (int i = 0; i < 4; i++) { addscl3(p1, p1, p[i].v1, b[i]); addscl3(p2, p2, p[i].v2, b[i]); }
    for (int i = 0; i < 3; i++) pos[i] = 0.5 * (p1[i] + p2[i]) / sum;
}

/* returns 0 on penetration (depth, dir, pos filled), -1 otherwise */
static int mpr_penetration(const Hull *h1, const Hull *h2, double tol, int max_iter, double *depth, double *dir_out, double *pos) {
    Supp p[4], v4;
    double dir[3], va[3], vb[3], dot;
    const double origin[3] = {0, 0, 0};
    /* discover portal */
    copy3(p[0].v1, h1->pos); copy3(p[0].v2, h2->pos); sub3(p[0].v, p[0].v1, p[0].v2);
    if (is_zero(p[0].v[0]) && is_zero(p[0].v[1]) && is_zero(p[0].v[2])) p[0].v[0] += CCD_EPS * 10;
    scl3(dir, p[0].v, -1); normalize3(dir);
    mink_support(h1, h2, dir, &p[1]);
    dot = dot3(p[1].v, dir);
    if (is_zero(dot) || dot < 0) return -1;
    cross3(dir, p[0].v, p[1].v);
    if (is_zero(dot3(dir, dir))) {
        /* origin on the v0-v1 ray: touching (res 1) or segment (res 2) case of libccd */
        if (is_zero(p[1].v[0]) && is_zero(p[1].v[1]) && is_zero(p[1].v[2])) {
            *depth = 0; dir_out[0] = dir_out[1] = dir_out[2] = 0;
            for (int i = 0; i < 3; i++) pos[i] = 0.5 * (p[1].v1[i] + p[1].v2[i]);
        } else {
            *depth = norm3(p[1].v);
            copy3(dir_out, p[1].v); normalize3(dir_out);
            for (int i = 0; i < 3; i++) pos[i] = 0.5 * (p[1].v1[i] + p[1].v2[i]);
        }
        return 0;
    }
    normalize3(dir);
    mink_support(h1, h2, dir, &p[2]);
    dot = dot3(p[2].v, dir);
    if (is_zero(dot) || dot < 0) return -1;
    sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
    cross3(dir, va, vb); normalize3(dir);
    if (dot3(dir, p[0].v) > 0) { Supp t = p[1]; p[1] = p[2]; p[2] = t; scl3(dir, dir, -1); }
    for (int guard = 0;; guard++) {
        if (guard > 100) return -1;
        mink_support(h1, h2, dir, &p[3]);
        dot = dot3(p[3].v, dir);
        if (is_zero(dot) || dot < 0) return -1;
        int cont = 0;
        cross3(va, p[1].v, p[3].v); dot = dot3(va, p[0].v);
        if (dot < 0 && !is_zero(dot)) { p[2] = p[3]; cont = 1; }
        if (!cont) {
            cross3(va, p[3].v, p[2].v); dot = dot3(va, p[0].v);
            if (dot < 0 && !is_zero(dot)) { p[1] = p[3]; cont = 1; }
        }
        if (!cont) break;
        sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
        cross3(dir, va, vb); normalize3(dir);
    }
    /* refine portal */
    for (int guard = 0;; guard++) {
        if (guard > 1000) return -1;
        portal_dir(p, dir);
        dot = dot3(dir, p[1].v);
        if (is_zero(dot) || dot > 0) break;
        mink_support(h1, h2, dir, &v4);
        dot = dot3(v4.v, dir);
        if (!(is_zero(dot) || dot > 0) || reach_tolerance(p, &v4, dir, tol)) return -1;
        expand_portal(p, &v4);
    }
    /* find penetration */
    for (int it = 0;; it++) {
        portal_dir(p, dir);
        mink_support(h1, h2, dir, &v4);
        if (reach_tolerance(p, &v4, dir, tol) || it > max_iter) {
            double w[3];
            *depth = sqrt(point_tri_dist2(origin, p[1].v, p[2].v, p[3].v, w));
            if (is_zero(w[0]) && is_zero(w[1]) && is_zero(w[2])) copy3(w, dir);
            copy3(dir_out, w); normalize3(dir_out);
            find_pos(p, pos);
            return 0;
        }
        expand_portal(p, &v4);
    }
}

static void make_hull(const KgoModel *m, int mesh, const double *pos, const double *mat, double margin, Hull *h) {
    h->vert = m->hull_vert + 3 * m->mesh[mesh].vert0;
    h->nvert = m->mesh[mesh].nvert;
    h->pos = pos; h->mat = mat; h->margin = margin;
}

int kgo_mpr(const KgoModel *m, int mesh1, const double pos1[3], const double mat1[9], int mesh2,
            const double pos2[3], const double mat2[9], double margin, double *depth, double dir[3], double pos[3]) {
    Hull h1, h2;
    make_hull(m, mesh1, pos1, mat1, 0.5 * margin, &h1);
    make_hull(m, mesh2, pos2, mat2, 0.5 * margin, &h2);
    return mpr_penetration(&h1, &h2, 1e-6, 50, depth, dir, pos);
}

/* mju_makeFrame: complete frame[0:3]=normal to an orthonormal frame */
static void make_frame(double *f) {
    double *x = f, *y = f + 3, *z = f + 6;
    y[0] = y[1] = y[2] = 0;
    if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
    addscl3(y, y, x, -dot3(x, y));
    normalize3(y);
    cross3(z, x, y);
}

/* ------------------------------------------------------------------ P4 + P5 collision */
static void collision(const KgoModel *m, const KgoState *s, KgoData *d) {
    d->ncon = 0; d->n_mpr = 0; d->n_broad = 0;
    const double margin = m->margin;
    for (int pi = 0; pi < m->npair; pi++) {
        int g1 = m->pair[pi].g1, g2 = m->pair[pi].g2;
        double rb2 = (g2 == 8) ? m->obj_rbound[s->shape] : m->geom_rbound[g2];
        const double *p2 = d->geom_xpos + 3 * g2, *R2 = d->geom_xmat + 9 * g2;
        int mesh2 = kgo_mesh_of_geom(m, g2, s->shape);
        if (g1 == 0) {
            /* plane z=0, normal +z (XML:148).  Broadphase: centre height vs bounding sphere */
            if (p2[2] > rb2 + margin) continue;
            d->n_broad++;
            const double *vert = m->hull_vert + 3 * m->mesh[mesh2].vert0;
            int nv = m->mesh[mesh2].nvert, best = 0;
            double ld[3], nneg[3] = {0, 0, -1}, bv = -1e300;
            multmv3(ld, R2, nneg);
            for (int i = 0; i < nv; i++) { double v = dot3(vert + 3 * i, ld); if (v > bv) { bv = v; best = i; } }
            double w[3];
            mulmv3(w, R2, vert + 3 * best); add3(w, w, p2);
            double dist = w[2];
            if (dist > margin) continue;
            int first = d->ncon, cnt = 0;
            double thresh = 0.3 * rb2;
            for (int c = -1; ; c++) {
                /* c == -1: the support vertex; then its hull-graph neighbours (mjc_PlaneConvex) */
                int vi;
                if (c < 0) vi = best;
                else {
                    int a0 = m->adj_start[m->mesh[mesh2].vert0 + best], a1 = m->adj_start[m->mesh[mesh2].vert0 + best + 1];
                    if (a0 + c >= a1) break;
                    vi = m->adj[a0 + c];
                }
                mulmv3(w, R2, vert + 3 * vi); add3(w, w, p2);
                dist = w[2];
                if (dist > margin) continue;
                double cp[3] = {w[0], w[1], w[2] - 0.5 * dist};
                int skip = 0;
                for (int k2 = 0; k2 < cnt; k2++) {
                    double df[3]; sub3(df, cp, d->con[first + k2].pos);
                    if (norm3(df) < thresh) { skip = 1; break; }
                }
                if (skip) continue;
                if (d->ncon >= KGO_MAXCON) break;
                KgoContact *con = &d->con[d->ncon++];
                memset(con, 0, sizeof *con);
                con->geom1 = 0; con->geom2 = g2; con->dist = dist; copy3(con->pos, cp);
                con->frame[0] = 0; con->frame[1] = 0; con->frame[2] = 1;
                make_frame(con->frame);
                con->friction = m->pair[pi].friction; con->includemargin = margin;
                if (++cnt >= 4) break;
            }
        } else {
            double rb1 = (g1 == 8) ? m->obj_rbound[s->shape] : m->geom_rbound[g1];
            const double *p1 = d->geom_xpos + 3 * g1, *R1 = d->geom_xmat + 9 * g1;
            double df[3]; sub3(df, p2, p1);
            if (norm3(df) > rb1 + rb2 + margin) continue;
            d->n_broad++; d->n_mpr++;
            int mesh1 = kgo_mesh_of_geom(m, g1, s->shape);
            double depth, dir[3], pos[3];
            if (kgo_mpr(m, mesh1, p1, R1, mesh2, p2, R2, margin, &depth, dir, pos) != 0) continue;
            if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) continue;
            double dist = margin - depth;
            if (dist >= margin) continue;
            if (d->ncon >= KGO_MAXCON) continue;
            KgoContact *con = &d->con[d->ncon++];
            memset(con, 0, sizeof *con);
            con->geom1 = g1; con->geom2 = g2; con->dist = dist; copy3(con->pos, pos);
            copy3(con->frame, dir);
            make_frame(con->frame);
            con->friction = m->pair[pi].friction; con->includemargin = margin;
        }
    }
}

/* ------------------------------------------------------------------ P6 constraint rows */
static void jac_point(const KgoModel *m, const Kin *k, int body, const double *p, double *jacp /*3 x NV*/) {
    memset(jacp, 0, 3 * KGO_NV * sizeof(double));
    if (body == 0) return;
    /* last dof of the body */
    int last = -1;
    for (int j = 0; j < KGO_NV; j++) if (m->dof_body[j] == body) last = j;
    for (int j = last; j >= 0; j = m->dof_parent[j]) {
        double v[3];
        cross3(v, k->cdof[j], p);
        for (int c = 0; c < 3; c++) jacp[KGO_NV * c + j] = v[c] + k->cdof[j][3 + c];
    }
}

static double impedance(const KgoModel *m, double pos, double margin) {
    double dmin = m->solimp[0], dmax = m->solimp[1], width = m->solimp[2];
    if (dmin < 1e-4) dmin = 1e-4; if (dmin > 0.9999) dmin = 0.9999;
    if (dmax < 1e-4) dmax = 1e-4; if (dmax > 0.9999) dmax = 0.9999;
    if (width <= 1e-15 || dmin == dmax) return 0.5 * (dmin + dmax);
    double x = fabs(pos - margin) / width;
    if (x >= 1) return dmax;
    double y = x <= 0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x);
    return dmin + y * (dmax - dmin);
}

static void add_row(const KgoModel *m, const KgoState *s, KgoData *d, const double *J, double pos, double margin,
                    double diagApprox, int type) {
    if (d->nefc >= KGO_MAXEFC) return;
    int i = d->nefc++;
    memcpy(d->efc_J + KGO_NV * i, J, KGO_NV * sizeof(double));
    d->efc_pos[i] = pos; d->efc_margin[i] = margin; d->efc_type[i] = type;
    double tc = m->solref[0], dr = m->solref[1];
    if (tc < 2 * m->timestep) tc = 2 * m->timestep; /* refsafe */
    double dmax = m->solimp[1];
    double K = 1.0 / (dmax * dmax * tc * tc * dr * dr), B = 2.0 / (dmax * tc);
    double imp = impedance(m, pos, margin);
    double R = (1 - imp) / imp * diagApprox;
    if (R < 1e-15) R = 1e-15;
    d->efc_R[i] = R;
    double vel = 0;
    for (int j = 0; j < KGO_NV; j++) vel += J[j] * s->qvel[j];
    d->efc_aref[i] = -B * vel - K * imp * (pos - margin);
}

static void make_constraints(const KgoModel *m, const KgoState *s, KgoData *d, const Kin *k) {
    d->nefc = 0;
    double J[KGO_NV];
    /* equality: fixed tendons L_k = q_prox - 2 q_dist held at L0 = 0 (XML:171-188) */
    for (int f = 0; f < 3; f++) {
        memset(J, 0, sizeof J);
        J[3 + 2 * f] = 1; J[4 + 2 * f] = -2;
        double L = s->qpos[3 + 2 * f] - 2 * s->qpos[4 + 2 * f];
        add_row(m, s, d, J, L, 0, m->tendon_invw[f], 0);
    }
    /* joint limits: slides +-0.5 (XML:62-64), proximal hinges [0,2] (XML:82,105,126); joint margin 0 */
    for (int a = 0; a < 6; a++) {
        int j = a < 3 ? a : 3 + 2 * (a - 3);
        double lo = a < 3 ? m->slide_range[a][0] : m->finger_range[a - 3][0];
        double hi = a < 3 ? m->slide_range[a][1] : m->finger_range[a - 3][1];
        double q = s->qpos[j];
        for (int side = 0; side < 2; side++) {
            double dist = side == 0 ? q - lo : hi - q;
            if (dist < 0) {
                memset(J, 0, sizeof J);
                J[j] = side == 0 ? 1 : -1;
                add_row(m, s, d, J, dist, 0, m->dof_invw[j], 1);
            }
        }
    }
    /* contacts: pyramidal, condim 3 -> 4 rows */
    for (int c = 0; c < d->ncon; c++) {
        KgoContact *con = &d->con[c];
        int b1 = m->geom_body[con->geom1], b2 = m->geom_body[con->geom2];
        double j1[3 * KGO_NV], j2[3 * KGO_NV], jd[3 * KGO_NV], jf[3][KGO_NV];
        jac_point(m, k, b1, con->pos, j1);
        jac_point(m, k, b2, con->pos, j2);
        for (int i = 0; i < 3 * KGO_NV; i++) jd[i] = j2[i] - j1[i];
        for (int r = 0; r < 3; r++) for (int j = 0; j < KGO_NV; j++)
            jf[r][j] = con->frame[3 * r] * jd[j] + con->frame[3 * r + 1] * jd[KGO_NV + j] + con->frame[3 * r + 2] * jd[2 * KGO_NV + j];
        double tran1 = (b1 == 8) ? m->obj_invw[s->shape][0] : m->body_invw[b1][0];
        double tran2 = (b2 == 8) ? m->obj_invw[s->shape][0] : m->body_invw[b2][0];
        double tran = tran1 + tran2, mu = con->friction;
        double dA = tran + mu * mu * tran;
        int first = d->nefc;
        for (int t = 0; t < 2; t++) for (int sg = 0; sg < 2; sg++) {
            for (int j = 0; j < KGO_NV; j++) J[j] = jf[0][j] + (sg == 0 ? mu : -mu) * jf[1 + t][j];
            add_row(m, s, d, J, con->dist, con->includemargin, dA, 2);
        }
        /* pyramid regulariser: R_py = 2 mu~^2 R, mu~ = mu / sqrt(impratio) (Appendix A "Friction cone") */
        if (d->nefc - first == 4) {
            double mut2 = mu * mu / m->impratio;
            double Rpy = 2 * mut2 * d->efc_R[first];
            if (Rpy < 1e-15) Rpy = 1e-15;
            for (int r = 0; r < 4; r++) d->efc_R[first + r] = Rpy;
        }
    }
    for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1.0 / d->efc_R[i];
}

/* ------------------------------------------------------------------ P7 Newton solver */
typedef struct { double a; double dq; } Brk;
static int brk_cmp(const void *x, const void *y) { double a = ((const Brk *)x)->a, b = ((const Brk *)y)->a; return a < b ? -1 : a > b; }

static double cost_eval(const KgoData *d, const double *qacc, const double *Ma_minus /* M(qacc-qacc_smooth) */, double *jar) {
    double cost = 0;
    for (int j = 0; j < KGO_NV; j++) cost += 0.5 * Ma_minus[j] * (qacc[j] - d->qacc_smooth[j]);
    for (int i = 0; i < d->nefc; i++) {
        double v = -d->efc_aref[i];
        for (int j = 0; j < KGO_NV; j++) v += d->efc_J[KGO_NV * i + j] * qacc[j];
        jar[i] = v;
        if (d->efc_type[i] == 0 || v < 0) cost += 0.5 * d->efc_D[i] * v * v;
    }
    return cost;
}

static void solve_newton(const KgoModel *m, const KgoState *s, KgoData *d) {
    const int nv = KGO_NV, n = d->nefc;
    double qacc[KGO_NV], Mdq[KGO_NV], jar[KGO_MAXEFC], jar2[KGO_MAXEFC];
    double scale = 1.0 / (m->meaninertia[s->shape] * nv);
    const double tol = 1e-8;
    /* warm start: pick the cheaper of qacc_warmstart and qacc_smooth */
    for (int j = 0; j < nv; j++) qacc[j] = s->qacc_warmstart[j];
    for (int i = 0; i < nv; i++) { Mdq[i] = 0; for (int j = 0; j < nv; j++) Mdq[i] += d->M[nv * i + j] * (qacc[j] - d->qacc_smooth[j]); }
    double cost = cost_eval(d, qacc, Mdq, jar);
    {
        double z[KGO_NV] = {0};
        double c0 = cost_eval(d, d->qacc_smooth, z, jar2);
        if (c0 < cost) { memcpy(qacc, d->qacc_smooth, sizeof qacc); memset(Mdq, 0, sizeof Mdq); memcpy(jar, jar2, n * sizeof(double)); cost = c0; }
    }
    int iter = 0;
    for (; iter < 100 && n > 0; iter++) {
        double grad[KGO_NV], H[KGO_NV * KGO_NV], dir[KGO_NV];
        memcpy(grad, Mdq, sizeof grad);
        memcpy(H, d->M, sizeof H);
        for (int i = 0; i < n; i++) {
            if (!(d->efc_type[i] == 0 || jar[i] < 0)) continue;
            const double *J = d->efc_J + nv * i;
            double Dj = d->efc_D[i];
            for (int a = 0; a < nv; a++) {
                if (J[a] == 0) continue;
                grad[a] += J[a] * Dj * jar[i];
                for (int b = 0; b < nv; b++) H[nv * a + b] += Dj * J[a] * J[b];
            }
        }
        double gn = 0;
        for (int j = 0; j < nv; j++) gn += grad[j] * grad[j];
        if (scale * sqrt(gn) < tol) break;
        if (chol(H, nv) != 0) break;
        for (int j = 0; j < nv; j++) dir[j] = -grad[j];
        chol_solve(H, nv, dir);
        /* exact line search on the piecewise-quadratic phi(alpha) */
        double Md[KGO_NV], jv[KGO_MAXEFC];
        for (int i = 0; i < nv; i++) { Md[i] = 0; for (int j = 0; j < nv; j++) Md[i] += d->M[nv * i + j] * dir[j]; }
        double p1 = 0, p2 = 0; /* phi'(0) and phi'' of the Gauss part */
        for (int j = 0; j < nv; j++) { p1 += Mdq[j] * dir[j]; p2 += Md[j] * dir[j]; }
        Brk brk[KGO_MAXEFC]; int nb = 0;
        double d1 = p1, d2 = p2; /* derivative and curvature at alpha=0+ */
        for (int i = 0; i < n; i++) {
            double v = 0;
            for (int j = 0; j < nv; j++) v += d->efc_J[nv * i + j] * dir[j];
            jv[i] = v;
            int act = d->efc_type[i] == 0 || jar[i] < 0;
            if (act) { d1 += d->efc_D[i] * jar[i] * v; d2 += d->efc_D[i] * v * v; }
            if (d->efc_type[i] != 0 && v != 0) {
                double a = -jar[i] / v;
                if (a > 0) { brk[nb].a = a; brk[nb].dq = (double)i; nb++; }
            }
        }
        qsort(brk, nb, sizeof(Brk), brk_cmp);
        double alpha = 0, a_prev = 0;
        int done = 0;
        for (int b = 0; b <= nb && !done; b++) {
            double a_next = b < nb ? brk[b].a : 1e300;
            /* on [a_prev, a_next): derivative d1 + d2 (alpha - a_prev) */
            if (d2 > 0) {
                double root = a_prev - d1 / d2;
                if (root <= a_next) { alpha = root < a_prev ? a_prev : root; done = 1; break; }
            } else if (d1 >= 0) { alpha = a_prev; done = 1; break; }
            if (b == nb) break;
            d1 += d2 * (a_next - a_prev);
            int i = (int)brk[b].dq;
            /* row i toggles at a_next: becomes active if jv<0 (jar decreasing through 0), else inactive */
            double c2 = d->efc_D[i] * jv[i] * jv[i];
            if (jv[i] < 0) d2 += c2; else d2 -= c2;
            a_prev = a_next;
        }
        if (!done) alpha = a_prev;
        for (int j = 0; j < nv; j++) { qacc[j] += alpha * dir[j]; Mdq[j] += alpha * Md[j]; }
        double cost_new = cost_eval(d, qacc, Mdq, jar);
        double improvement = scale * (cost - cost_new);
        cost = cost_new;
        if (improvement < tol) { iter++; break; }
    }
    d->solver_iter = iter; d->solver_cost = cost;
    memcpy(d->qacc, qacc, sizeof qacc);
    memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
    for (int i = 0; i < n; i++) {
        int act = d->efc_type[i] == 0 || jar[i] < 0;
        d->efc_force[i] = act ? -d->efc_D[i] * jar[i] : 0;
        for (int j = 0; j < nv; j++) d->qfrc_constraint[j] += d->efc_J[nv * i + j] * d->efc_force[i];
    }
}

/* ------------------------------------------------------------------ mj_forward / mj_step */
/* given != NULL: the contact list is taken from the caller (given_n x {geom1, geom2, dist, pos[3], normal[3]}) instead of
 * the collision stage -- test infrastructure: it lets a parity test hand the oracle the contacts ANOTHER narrowphase
 * found, so that the constraint rows and the solver are compared on identical input */
static void forward_impl(const KgoModel *m, const KgoState *s, const double ctrl[KGO_NU], int flags, KgoData *d,
                         int given_n, const double *given) {
    Kin k;
    kinematics(m, s, d, &k);
    crba_rne(m, s, d, &k);
    if (!given) collision(m, s, d);
    else {
        d->ncon = 0; d->n_mpr = 0; d->n_broad = 0;
        for (int c = 0; c < given_n && d->ncon < KGO_MAXCON; c++) {
            const double *g = given + 9 * c;
            KgoContact *con = &d->con[d->ncon++];
            memset(con, 0, sizeof *con);
            con->geom1 = (int)g[0]; con->geom2 = (int)g[1]; con->dist = g[2];
            copy3(con->pos, g + 3); copy3(con->frame, g + 6);
            make_frame(con->frame);
            con->includemargin = m->margin;
            for (int pi = 0; pi < m->npair; pi++)
                if (m->pair[pi].g1 == con->geom1 && m->pair[pi].g2 == con->geom2) con->friction = m->pair[pi].friction;
        }
    }
    make_constraints(m, s, d, &k);
    /* sensors (position stage): 9 jointpos then 17 rangefinders (XML:247-288) */
    for (int i = 0; i < 3; i++) d->sensordata[i] = s->qpos[i];
    for (int f = 0; f < 3; f++) { d->sensordata[3 + f] = s->qpos[3 + 2 * f]; d->sensordata[6 + f] = s->qpos[4 + 2 * f]; }
    if (flags & KGO_RAYS) kgo_rangefinders(m, s, d);
    else for (int i = 0; i < KGO_NSITE; i++) d->sensordata[9 + i] = -1;
    actuation(m, s, ctrl, d);
    double L[KGO_NV * KGO_NV];
    memcpy(L, d->M, sizeof L);
    chol(L, KGO_NV);
    for (int j = 0; j < KGO_NV; j++) {
        d->qfrc_smooth[j] = d->qfrc_passive[j] - d->qfrc_bias[j] + d->qfrc_actuator[j];
        d->qacc_smooth[j] = d->qfrc_smooth[j];
    }
    chol_solve(L, KGO_NV, d->qacc_smooth);
    if (d->nefc == 0) {
        memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc);
        memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
        d->solver_iter = 0; d->solver_cost = 0;
    } else {
        solve_newton(m, s, d);
    }
}

void kgo_forward(const KgoModel *m, const KgoState *s, const double ctrl[KGO_NU], int flags, KgoData *d) {
    forward_impl(m, s, ctrl, flags, d, 0, NULL);
}
void kgo_forward_contacts(const KgoModel *m, const KgoState *s, const double ctrl[KGO_NU], int flags, int ncon,
                          const double *con, KgoData *d) {
    forward_impl(m, s, ctrl, flags, d, ncon, con);
}

static void euler(const KgoModel *m, KgoState *s, KgoData *d);
void kgo_step(const KgoModel *m, KgoState *s, const double ctrl[KGO_NU], int flags, KgoData *d) {
    kgo_forward(m, s, ctrl, flags, d);
    euler(m, s, d);
}
/* test hook: one mj_step whose contact list is given by the caller (see kgo_forward_contacts) */
void kgo_step_contacts(const KgoModel *m, KgoState *s, const double ctrl[KGO_NU], int flags, int ncon, const double *con,
                       KgoData *d) {
    forward_impl(m, s, ctrl, flags, d, ncon, con);
    euler(m, s, d);
}
static void euler(const KgoModel *m, KgoState *s, KgoData *d) {
    const double h = m->timestep;
    /* mj_Euler: joint damping implicit, (M + h B) a = qfrc_smooth + qfrc_constraint */
    double A[KGO_NV * KGO_NV], a[KGO_NV];
    memcpy(A, d->M, sizeof A);
    for (int j = 0; j < KGO_NV; j++) { A[KGO_NV * j + j] += h * m->damping; a[j] = d->qfrc_smooth[j] + d->qfrc_constraint[j]; }
    chol(A, KGO_NV);
    chol_solve(A, KGO_NV, a);
    memcpy(s->qacc_warmstart, d->qacc, sizeof d->qacc);
    for (int j = 0; j < KGO_NV; j++) s->qvel[j] += h * a[j];
    for (int j = 0; j < 9; j++) s->qpos[j] += h * s->qvel[j];
    for (int j = 0; j < 3; j++) s->qpos[9 + j] += h * s->qvel[9 + j];
    double w[3] = {s->qvel[12], s->qvel[13], s->qvel[14]};
    double ang = h * norm3(w);
    if (ang > 0) {
        double ax[3]; copy3(ax, w); normalize3(ax);
        double qr[4] = {cos(0.5 * ang), ax[0] * sin(0.5 * ang), ax[1] * sin(0.5 * ang), ax[2] * sin(0.5 * ang)};
        kgo_mulquat(s->qpos + 12, s->qpos + 12, qr);
    }
    normquat(s->qpos + 12);
}
