/* kgo_internal.h -- shared private declarations of the fp64 oracle (test infrastructure). */
#ifndef KGO_INTERNAL_H
#define KGO_INTERNAL_H
#include "kgo.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

#define NMESH 27
typedef struct {
    int vert0, nvert, adj0, tri0, ntri, bvh0, nbvh, nface;
} MeshIndex;

typedef struct {
    int g1, g2;       /* geom ids; for plane pairs g1 == 0 */
    double friction;
} PairDef;

struct KgoModel {
    unsigned char *blob;
    size_t nbytes;
    int nsec;
    double timestep, impratio, margin, damping, armature, obj_mass, gravity, solref[2], solimp[3];
    /* tree */
    int body_parent[KGO_NBODY];
    double body_pos[KGO_NBODY][3], body_quat[KGO_NBODY][4];
    double body_ipos[KGO_NBODY][3], body_iquat[KGO_NBODY][4], body_mass[KGO_NBODY], body_inertia[KGO_NBODY][3];
    double slide_axes[3][3], slide_range[3][2], finger_range[3][2];
    int dof_body[KGO_NV], dof_parent[KGO_NV];
    /* geoms (object entries are per shape) */
    int geom_body[KGO_NGEOM], geom_mesh[KGO_NGEOM];
    double geom_pos[KGO_NGEOM][3], geom_quat[KGO_NGEOM][4], geom_rbound[KGO_NGEOM], geom_size[KGO_NGEOM][3];
    double obj_gpos[KGO_NSHAPE][3], obj_gquat[KGO_NSHAPE][4], obj_rbound[KGO_NSHAPE], obj_size[KGO_NSHAPE][3];
    double obj_inertia[KGO_NSHAPE][3], obj_invw[KGO_NSHAPE][4], meaninertia[KGO_NSHAPE];
    int obj_flags[KGO_NSHAPE];   /* bit 0: _get_obj_size always swaps size[0] and size[2] ('lemon' in self.filename, ENV:721) */
    /* meshes */
    MeshIndex mesh[NMESH];
    const double *hull_vert;
    const int *adj_start, *adj;
    const float *tri_vert, *bvh_lo, *bvh_hi;
    const int *bvh_meta;
    /* sites */
    int site_body[KGO_NSITE], fod_site[12];
    double site_pos[KGO_NSITE][3], site_z[KGO_NSITE][3];
    /* actuators */
    double act_kv[KGO_NU], act_gear[KGO_NU], act_range[KGO_NU][2];
    /* constraint weights */
    double body_invw[KGO_NBODY][2], dof_invw[KGO_NV], tendon_invw[3];
    /* pairs */
    int npair;
    PairDef pair[40];
};

/* small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void add3(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void scl3(double *r, const double *a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void addscl3(double *r, const double *a, const double *b, double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
static inline void copy3(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double *a) {
    double n = norm3(a);
    if (n < 1e-15) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
    a[0] /= n; a[1] /= n; a[2] /= n; return n;
}
/* r = M v (row-major 3x3) */
static inline void mulmv3(double *r, const double *M, const double *v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
/* r = M^T v */
static inline void multmv3(double *r, const double *M, const double *v) {
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
void kgo_quat2mat(double *M, const double *q);
void kgo_mulquat(double *r, const double *a, const double *b);
void kgo_euler2quat(double *q, const double *e);
void kgo_mulmm3(double *r, const double *A, const double *B);

const void *kgo_section(const KgoModel *m, const char *name, int *count, int *dtype);

/* rays */
void kgo_rangefinders(const KgoModel *m, const KgoState *s, KgoData *d);
#endif
