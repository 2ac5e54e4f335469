/* kgo_model.c -- model-card blob parser for the fp64 oracle (test infrastructure).
 * Restates the structure of kinova_description/j2s7s300_end_effector_v1_<Shape>.xml:54-223. */
#include "kgo_internal.h"
#include <stdint.h>

typedef struct {
    char name[24];
    uint32_t dtype, count;
    uint64_t offset;
} Section;

const void *kgo_section(const KgoModel *m, const char *name, int *count, int *dtype) {
    const Section *tab = (const Section *)(m->blob + 16);
    for (int i = 0; i < m->nsec; i++) {
        if (strncmp(tab[i].name, name, 24) == 0) {
            if (count) *count = (int)tab[i].count;
            if (dtype) *dtype = (int)tab[i].dtype;
            return m->blob + tab[i].offset;
        }
    }
    fprintf(stderr, "kgo: missing section %s\n", name);
    abort();
}

const double *kgo_model_section(const KgoModel *m, const char *name, int *count) {
    int dt;
    const void *p = kgo_section(m, name, count, &dt);
    return dt == 0 ? (const double *)p : NULL;
}

static const double *secd(const KgoModel *m, const char *name, int expect) {
    int c, dt;
    const void *p = kgo_section(m, name, &c, &dt);
    if (dt != 0 || (expect > 0 && c != expect)) { fprintf(stderr, "kgo: bad section %s (%d,%d)\n", name, c, dt); abort(); }
    return (const double *)p;
}
static const int *seci(const KgoModel *m, const char *name, int expect) {
    int c, dt;
    const void *p = kgo_section(m, name, &c, &dt);
    if (dt != 1 || (expect > 0 && c != expect)) { fprintf(stderr, "kgo: bad section %s\n", name); abort(); }
    return (const int *)p;
}
static const float *secf(const KgoModel *m, const char *name) {
    int c, dt;
    const void *p = kgo_section(m, name, &c, &dt);
    if (dt != 2) abort();
    return (const float *)p;
}

int kgo_mesh_of_geom(const KgoModel *m, int geom, int shape) {
    if (geom == 8) return 3 + shape;
    return m->geom_mesh[geom];
}

KgoModel *kgo_model_load(const void *blob, size_t nbytes) {
    if (nbytes < 16 || memcmp(blob, "KGMC0001", 8) != 0) return NULL;
    KgoModel *m = (KgoModel *)calloc(1, sizeof(KgoModel));
    m->blob = (unsigned char *)malloc(nbytes);
    memcpy(m->blob, blob, nbytes);
    m->nbytes = nbytes;
    m->nsec = (int)*(const uint32_t *)(m->blob + 8);

    const double *opt = secd(m, "opt", 12);
    m->timestep = opt[0]; m->impratio = opt[1]; m->margin = opt[2]; m->damping = opt[3];
    m->armature = opt[4]; m->obj_mass = opt[5]; m->gravity = opt[6];
    m->solref[0] = opt[7]; m->solref[1] = opt[8];
    m->solimp[0] = opt[9]; m->solimp[1] = opt[10]; m->solimp[2] = opt[11];

    static const int parent[KGO_NBODY] = {-1, 0, 1, 2, 1, 4, 1, 6, 0};
    memcpy(m->body_parent, parent, sizeof parent);
    for (int b = 0; b < KGO_NBODY; b++) { m->body_quat[b][0] = 1; m->body_iquat[b][0] = 1; }
    const double *l7 = secd(m, "l7_pos", 3), *l7i = secd(m, "l7_inertial", 7);
    copy3(m->body_pos[1], l7);
    copy3(m->body_ipos[1], l7i); m->body_mass[1] = l7i[3]; copy3(m->body_inertia[1], l7i + 4);
    memcpy(m->slide_axes, secd(m, "slide_axes", 9), 9 * sizeof(double));
    memcpy(m->slide_range, secd(m, "slide_range", 6), 6 * sizeof(double));
    memcpy(m->finger_range, secd(m, "finger_range", 6), 6 * sizeof(double));
    const double *fp = secd(m, "finger_pos", 9), *fq = secd(m, "finger_quat", 12), *fi = secd(m, "finger_inertial", 21);
    const double *tp = secd(m, "tip_pos", 9), *ti = secd(m, "tip_inertial", 21);
    for (int k = 0; k < 3; k++) {
        int bf = 2 + 2 * k, bt = 3 + 2 * k;
        copy3(m->body_pos[bf], fp + 3 * k); memcpy(m->body_quat[bf], fq + 4 * k, 4 * sizeof(double));
        copy3(m->body_ipos[bf], fi + 7 * k); m->body_mass[bf] = fi[7 * k + 3]; copy3(m->body_inertia[bf], fi + 7 * k + 4);
        copy3(m->body_pos[bt], tp + 3 * k);
        copy3(m->body_ipos[bt], ti + 7 * k); m->body_mass[bt] = ti[7 * k + 3]; copy3(m->body_inertia[bt], ti + 7 * k + 4);
    }
    m->body_mass[8] = m->obj_mass;
    static const int dof_body[KGO_NV] = {1, 1, 1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8};
    static const int dof_parent[KGO_NV] = {-1, 0, 1, 2, 3, 2, 5, 2, 7, -1, 9, 10, 11, 12, 13};
    memcpy(m->dof_body, dof_body, sizeof dof_body);
    memcpy(m->dof_parent, dof_parent, sizeof dof_parent);

    /* meshes */
    const int *mi = seci(m, "mesh_index", NMESH * 8);
    memcpy(m->mesh, mi, sizeof(MeshIndex) * NMESH);
    m->hull_vert = secd(m, "hull_vert", 0);
    m->adj_start = seci(m, "hull_adj_start", 0);
    m->adj = seci(m, "hull_adj", 0);
    m->tri_vert = secf(m, "tri_vert");
    m->bvh_lo = secf(m, "bvh_lo");
    m->bvh_hi = secf(m, "bvh_hi");
    m->bvh_meta = seci(m, "bvh_meta", 0);
    const double *mg = secd(m, "mesh_geom", NMESH * 15);
    /* geoms */
    static const int gbody[KGO_NGEOM] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
    static const int gmesh[KGO_NGEOM] = {-1, 0, 1, 2, 1, 2, 1, 2, 3};
    memcpy(m->geom_body, gbody, sizeof gbody);
    memcpy(m->geom_mesh, gmesh, sizeof gmesh);
    m->geom_quat[0][0] = 1;
    for (int g = 1; g < 8; g++) {
        const double *c = mg + 15 * gmesh[g];
        copy3(m->geom_pos[g], c); memcpy(m->geom_quat[g], c + 3, 4 * sizeof(double));
        m->geom_rbound[g] = c[7]; copy3(m->geom_size[g], c + 8);
    }
    for (int s = 0; s < KGO_NSHAPE; s++) {
        const double *c = mg + 15 * (3 + s);
        copy3(m->obj_gpos[s], c); memcpy(m->obj_gquat[s], c + 3, 4 * sizeof(double));
        m->obj_rbound[s] = c[7]; copy3(m->obj_size[s], c + 8);
        for (int i = 0; i < 3; i++) m->obj_inertia[s][i] = c[12 + i] * m->obj_mass / c[11];
    }
    memcpy(m->obj_invw, secd(m, "obj_invw0", KGO_NSHAPE * 4), sizeof m->obj_invw);
    memcpy(m->obj_flags, seci(m, "shape_flags", KGO_NSHAPE), sizeof m->obj_flags);
    /* sites */
    memcpy(m->site_body, seci(m, "site_body", KGO_NSITE), sizeof m->site_body);
    memcpy(m->site_pos, secd(m, "site_pos", KGO_NSITE * 3), sizeof m->site_pos);
    memcpy(m->site_z, secd(m, "site_zaxis", KGO_NSITE * 3), sizeof m->site_z);
    memcpy(m->fod_site, seci(m, "fod_site", 12), sizeof m->fod_site);
    /* actuators */
    memcpy(m->act_kv, secd(m, "act_kv", KGO_NU), sizeof m->act_kv);
    memcpy(m->act_gear, secd(m, "act_gear", KGO_NU), sizeof m->act_gear);
    memcpy(m->act_range, secd(m, "act_ctrlrange", KGO_NU * 2), sizeof m->act_range);
    /* constraint weights */
    const double *hb = secd(m, "hand_body_invw0", 14), *hd = secd(m, "hand_dof_invw0", 9);
    for (int b = 0; b < 7; b++) { m->body_invw[1 + b][0] = hb[2 * b]; m->body_invw[1 + b][1] = hb[2 * b + 1]; }
    for (int j = 0; j < 9; j++) m->dof_invw[j] = hd[j];
    memcpy(m->tendon_invw, secd(m, "tendon_invw0", 3), sizeof m->tendon_invw);
    const double *hm = secd(m, "hand_M0", 81);
    double tr = 0;
    for (int j = 0; j < 9; j++) tr += hm[10 * j];
    for (int s = 0; s < KGO_NSHAPE; s++) {
        /* diag of the object's 6x6 block at qpos0: m+armature (x3), body-frame inertia about the joint origin + armature */
        const double *c = m->obj_gpos[s];
        double R[9], I[9] = {0};
        kgo_quat2mat(R, m->obj_gquat[s]);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) I[3 * i + j] += R[3 * i + k] * m->obj_inertia[s][k] * R[3 * j + k];
        double cc = dot3(c, c), t = tr + 3 * (m->obj_mass + m->armature);
        for (int i = 0; i < 3; i++) t += I[4 * i] + m->obj_mass * (cc - c[i] * c[i]) + m->armature;
        m->meaninertia[s] = t / KGO_NV;
    }
    /* pairs: plane first, then the explicit object pairs (XML:159-166), then dynamic pairs */
    const double *pf = secd(m, "pair_friction", 8);
    int n = 0;
    for (int g = 1; g <= 8; g++) { m->pair[n].g1 = 0; m->pair[n].g2 = g; m->pair[n].friction = (g == 8) ? pf[0] : 1.0; n++; }
    for (int g = 2; g <= 7; g++) { m->pair[n].g1 = 8; m->pair[n].g2 = g; m->pair[n].friction = pf[g - 1]; n++; }
    m->pair[n].g1 = 8; m->pair[n].g2 = 1; m->pair[n].friction = pf[7]; n++;
    static const int dyn[15][2] = {{1, 3}, {1, 5}, {1, 7}, {2, 4}, {2, 5}, {3, 4}, {3, 5}, {2, 6}, {2, 7}, {3, 6}, {3, 7}, {4, 6}, {4, 7}, {5, 6}, {5, 7}};
    for (int i = 0; i < 15; i++) { m->pair[n].g1 = dyn[i][0]; m->pair[n].g2 = dyn[i][1]; m->pair[n].friction = 1.0; n++; }
    m->npair = n;
    return m;
}

void kgo_model_free(KgoModel *m) {
    if (!m) return;
    free(m->blob);
    free(m);
}

int kgo_sizeof_data(void) { return (int)sizeof(KgoData); }
int kgo_sizeof_state(void) { return (int)sizeof(KgoState); }
int kgo_sizeof_env(void) { return (int)sizeof(KgoEnv); }
