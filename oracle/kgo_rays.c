/* kgo_rays.c -- rangefinder sensors of the oracle (test infrastructure).
 * Restates MuJoCo's mj_ray for the 17 rangefinders of XML:265-288: ray from the site origin along
 * the site +z axis, geoms of the site's own body excluded, the ground plane hit from above only,
 * mesh geoms culled by the box of half-sizes geom_size and then intersected as raw triangles
 * (no back-face culling); -1 when nothing is hit (SURVEY Appendix A "Sensors"). */
#include "kgo_internal.h"

static double ray_box(const double *size, const double *p, const double *v) {
    /* nearest x >= 0 with p + x v on the surface of the box, -1 if none */
    double best = -1;
    for (int ax = 0; ax < 3; ax++) {
        if (fabs(v[ax]) < 1e-15) continue;
        for (int sg = -1; sg <= 1; sg += 2) {
            double x = (sg * size[ax] - p[ax]) / v[ax];
            if (x < 0) continue;
            int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
            double c1 = p[a1] + x * v[a1], c2 = p[a2] + x * v[a2];
            if (fabs(c1) <= size[a1] && fabs(c2) <= size[a2]) if (best < 0 || x < best) best = x;
        }
    }
    return best;
}

static double ray_tri(const float *t, const double *p, const double *v, const double *b0, const double *b1) {
    double pl[3][2], dif[3];
    for (int i = 0; i < 3; i++) {
        for (int c = 0; c < 3; c++) dif[c] = (double)t[3 * i + c] - p[c];
        pl[i][0] = dot3(b0, dif); pl[i][1] = dot3(b1, dif);
    }
    for (int c = 0; c < 2; c++) {
        if (pl[0][c] > 0 && pl[1][c] > 0 && pl[2][c] > 0) return -1;
        if (pl[0][c] < 0 && pl[1][c] < 0 && pl[2][c] < 0) return -1;
    }
    double A0 = pl[0][0] - pl[2][0], A1 = pl[1][0] - pl[2][0], A2 = pl[0][1] - pl[2][1], A3 = pl[1][1] - pl[2][1];
    double bb0 = -pl[2][0], bb1 = -pl[2][1];
    double det = A0 * A3 - A1 * A2;
    if (fabs(det) < 1e-15) return -1;
    double t0 = (A3 * bb0 - A1 * bb1) / det, t1 = (-A2 * bb0 + A0 * bb1) / det;
    if (t0 < 0 || t1 < 0 || t0 + t1 > 1) return -1;
    double e0[3], e1[3], e2[3], nrm[3];
    for (int c = 0; c < 3; c++) { e0[c] = (double)t[c] - t[6 + c]; e1[c] = (double)t[3 + c] - t[6 + c]; e2[c] = p[c] - t[6 + c]; }
    cross3(nrm, e0, e1);
    double den = dot3(v, nrm);
    if (fabs(den) < 1e-15) return -1;
    double x = -dot3(e2, nrm) / den;
    return x >= 0 ? x : -1;
}

static int ray_aabb(const float *lo, const float *hi, const double *p, const double *inv, double tmax) {
    double t0 = 0, t1 = tmax;
    for (int c = 0; c < 3; c++) {
        double a = ((double)lo[c] - p[c]) * inv[c], b = ((double)hi[c] - p[c]) * inv[c];
        if (a > b) { double s = a; a = b; b = s; }
        if (a > t0) t0 = a;
        if (b < t1) t1 = b;
        if (t0 > t1 * (1 + 1e-12) + 1e-12) return 0;
    }
    return 1;
}

double kgo_ray_mesh(const KgoModel *m, int mesh, const double gpos[3], const double gmat[9], const double size[3],
                    const double pnt[3], const double vec[3], int use_bvh) {
    double dp[3], lp[3], lv[3];
    sub3(dp, pnt, gpos);
    multmv3(lp, gmat, dp); multmv3(lv, gmat, vec);
    if (ray_box(size, lp, lv) < 0) return -1;
    /* basis of the plane normal to the ray */
    double b0[3], b1[3];
    if (fabs(lv[0]) < fabs(lv[1]) && fabs(lv[0]) < fabs(lv[2])) { b0[0] = 0; b0[1] = lv[2]; b0[2] = -lv[1]; }
    else if (fabs(lv[1]) < fabs(lv[2])) { b0[0] = -lv[2]; b0[1] = 0; b0[2] = lv[0]; }
    else { b0[0] = lv[1]; b0[1] = -lv[0]; b0[2] = 0; }
    normalize3(b0);
    cross3(b1, b0, lv);
    normalize3(b1);
    const MeshIndex *mi = &m->mesh[mesh];
    const float *tv = m->tri_vert + 9 * (size_t)mi->tri0;
    double best = -1;
    if (!use_bvh) {
        for (int i = 0; i < mi->ntri; i++) {
            double x = ray_tri(tv + 9 * i, lp, lv, b0, b1);
            if (x >= 0 && (best < 0 || x < best)) best = x;
        }
        return best;
    }
    double inv[3];
    for (int c = 0; c < 3; c++) inv[c] = fabs(lv[c]) > 1e-300 ? 1.0 / lv[c] : (lv[c] >= 0 ? 1e300 : -1e300);
    const float *lo = m->bvh_lo + 3 * (size_t)mi->bvh0, *hi = m->bvh_hi + 3 * (size_t)mi->bvh0;
    const int *meta = m->bvh_meta + 2 * (size_t)mi->bvh0;
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        int n = stack[--sp];
        if (!ray_aabb(lo + 3 * n, hi + 3 * n, lp, inv, best < 0 ? 1e300 : best)) continue;
        if (meta[2 * n + 1] == 0) { stack[sp++] = meta[2 * n]; stack[sp++] = meta[2 * n] + 1; }
        else for (int i = 0; i < meta[2 * n + 1]; i++) {
            double x = ray_tri(tv + 9 * (meta[2 * n] + i), lp, lv, b0, b1);
            if (x >= 0 && (best < 0 || x < best)) best = x;
        }
    }
    return best;
}

void kgo_rangefinders(const KgoModel *m, const KgoState *s, KgoData *d) {
    for (int i = 0; i < KGO_NSITE; i++) {
        const double *p = d->site_xpos + 3 * i, *v = d->site_zaxis + 3 * i;
        int sb = m->site_body[i], own;
        own = sb == 0 ? 1 : (sb <= 3 ? 2 * sb : 2 * (sb - 3) + 1);
        double best = -1;
        /* ground plane: front face only */
        if (v[2] < -1e-15) {
            double x = -p[2] / v[2];
            if (x >= 0) {
                double hx = p[0] + x * v[0], hy = p[1] + x * v[1];
                if (fabs(hx) <= 2 && fabs(hy) <= 2) best = x;
            }
        }
        for (int g = 1; g < KGO_NGEOM; g++) {
            if (m->geom_body[g] == own) continue;
            const double *size = (g == 8) ? m->obj_size[s->shape] : m->geom_size[g];
            double x = kgo_ray_mesh(m, kgo_mesh_of_geom(m, g, s->shape), d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, size, p, v, 1);
            if (x >= 0 && (best < 0 || x < best)) best = x;
        }
        d->sensordata[9 + i] = best;
    }
}
