// kg_model.h -- fp32 device-side model card (POD) + host loader from assets/model_card.bin.
// Replaces what mujoco_py.load_model_from_path + MjSim build on every reset
// (kinova_gripper_env.py:878-879, 1002-1003) with one immutable table set per GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define KG_NSHAPE 24
#define KG_NMESH 27
#define KG_NSITE 17

struct KgMesh {
    int vert0, nvert, adj0, tri0, ntri, bvh0, nbvh, nface;
};

struct KgShape {
    float gpos[3];        // object geom centre (mesh centre of mass) in the body frame
    float gR[9];          // geom (mesh inertial) frame in the body frame, row-major
    float rbound;
    float size[3];        // geom_size (inertia-box half sizes): ray culling box
    float obs_size[3];    // _get_obj_size() result (ENV:706-746)
    float Ic[6];          // inertia about the centre of mass, body frame: xx,yy,zz,xy,xz,yz
    float Io[6];          // inertia about the body origin, body frame (+ nothing else)
    float invw_tran;      // body_invweight0 (translational)
    float scale;          // 1 / (meaninertia * nv): Newton tolerance scaling
    float Sinv[2][6];     // inverse Schur complement of the object's 6x6 mass block (rotational part after eliminating the
                          // translation): CONSTANT in the body frame, [0] for M, [1] for M + timestep * damping (kg_minv)
};

struct KgModel {
    float timestep, impratio, margin, damping, armature, obj_mass, gravity;
    float solK, solB, dmin, dmax, width;   // from solref (0.02,1) / solimp (0.9,0.95,0.001), refsafe applied
    float l7_pos[3], slide_axes[9], slide_range[6], finger_range[6];
    float m_hand_total;
    float finger_pos[9], finger_R[27], tip_pos[3], link_com[3];
    float link_mass, link_Izz;             // proximal == distal inertial (XML:81,90)
    float palm_gpos[3], palm_gR[9], prox_gpos[3], prox_gR[9], dist_gpos[3], dist_gR[9];
    float hand_rbound[3], hand_size[9];    // palm, proximal, distal
    float site_pos[KG_NSITE * 3], site_z[KG_NSITE * 3];
    int site_body[KG_NSITE], fod_site[12];
    float body_invw[8];                    // translational invweight0 of geoms 0..7's bodies (ground = 0)
    float dof_invw[9], tendon_invw[3];
    float act_kv_slide, act_gear_slide, act_kv_finger, ctrl_slide, ctrl_finger;
    float pair_mu[30];
    KgShape shape[KG_NSHAPE];
    KgMesh mesh[KG_NMESH];
    // big read-only tables (device pointers on the GPU, host pointers in the emulation build)
    const float *hull_vert;        // xyzw per vertex, 128-bit loads; w = bits of (first OVERFLOW chunk | n overflow chunks << 24); chunk 0 = the vertex's global index
    const float *hull_pad;         // the same vertices, every mesh padded to a multiple of 32 with copies of its vertex 0 (branch-free scans)
    int hpad0[KG_NMESH], hnpad[KG_NMESH];
    // per mesh and direction octant (bit c set = component c negative): the vertices that can be a support vertex for SOME
    // direction of that closed octant, ascending vertex index, xyz + bits(vertex index), padded to a multiple of 32 with
    // copies of the list's first entry.  Exact: the arg-max of a scan over the list equals the arg-max over the mesh.
    const float *hull_oct;
    int oct0[KG_NMESH][8], octn[KG_NMESH][8];
    const unsigned short *adj8;    // neighbour lists in 16-byte chunks of 8 mesh-local vertex ids, padded with the vertex itself; chunk gv = first chunk of vertex gv
    const float *nbr8;             // the same chunks with the neighbours' vertex records (xyzw as in hull_vert) inline: 128 bytes per chunk
    const int *adj_start;
    const unsigned short *adj;
    const float *tri;              // 9 floats per triangle, BVH leaf order
    const float *bvh_node;         // 8 floats per node: lo.xyz, bits(left child | first triangle), hi.xyz, bits(triangle count; 0 = inner)
    const float *pose_train, *pose_test;  // [shape][orient][row][9]
    int pose_rows_train, pose_rows_test;
};

// ---------------------------------------------------------------------------------- host loader
struct KgHostTables {
    std::vector<float> hull_vert, hull_pad, hull_oct, tri, bvh_node, pose_train, pose_test;
    std::vector<int> adj_start;
    std::vector<unsigned short> adj, adj8;
    std::vector<float> nbr8;
};

namespace kgdetail {
struct Sec { char name[24]; uint32_t dtype, count; uint64_t offset; };
inline const Sec *find(const unsigned char *blob, const char *name) {
    uint32_t n = *(const uint32_t *)(blob + 8);
    const Sec *t = (const Sec *)(blob + 16);
    for (uint32_t i = 0; i < n; i++) if (std::strncmp(t[i].name, name, 24) == 0) return &t[i];
    return nullptr;
}
inline void q2m(const double *q, double *M) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
    M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
    M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
}  // namespace kgdetail

// Fills `m` (pointers left null) and `t`.  Returns "" or an error string.
inline std::string kg_load_model(const void *blob_, size_t nbytes, KgModel *m, KgHostTables *t) {
    using namespace kgdetail;
    const unsigned char *blob = (const unsigned char *)blob_;
    if (nbytes < 16 || std::memcmp(blob, "KGMC0001", 8) != 0) return "bad magic";
    std::memset(m, 0, sizeof *m);
    auto D = [&](const char *n, size_t expect) -> const double * {
        const Sec *s = find(blob, n);
        if (!s || s->dtype != 0 || (expect && s->count != expect)) return nullptr;
        return (const double *)(blob + s->offset);
    };
    auto I = [&](const char *n, size_t *cnt) -> const int * {
        const Sec *s = find(blob, n);
        if (!s || s->dtype != 1) return nullptr;
        if (cnt) *cnt = s->count;
        return (const int *)(blob + s->offset);
    };
    auto F = [&](const char *n, size_t *cnt) -> const float * {
        const Sec *s = find(blob, n);
        if (!s || s->dtype != 2) return nullptr;
        if (cnt) *cnt = s->count;
        return (const float *)(blob + s->offset);
    };
#define KG_NEED(p) if (!(p)) return std::string("missing/bad section at ") + #p
    const double *opt = D("opt", 12); KG_NEED(opt);
    m->timestep = (float)opt[0]; m->impratio = (float)opt[1]; m->margin = (float)opt[2]; m->damping = (float)opt[3];
    m->armature = (float)opt[4]; m->obj_mass = (float)opt[5]; m->gravity = (float)opt[6];
    {
        double tc = opt[7], dr = opt[8], dmax = opt[10];
        if (tc < 2 * opt[0]) tc = 2 * opt[0];
        m->solK = (float)(1.0 / (dmax * dmax * tc * tc * dr * dr));
        m->solB = (float)(2.0 / (dmax * tc));
        m->dmin = (float)opt[9]; m->dmax = (float)dmax; m->width = (float)opt[11];
    }
    const double *p;
    p = D("l7_pos", 3); KG_NEED(p); for (int i = 0; i < 3; i++) m->l7_pos[i] = (float)p[i];
    p = D("slide_axes", 9); KG_NEED(p); for (int i = 0; i < 9; i++) m->slide_axes[i] = (float)p[i];
    p = D("slide_range", 6); KG_NEED(p); for (int i = 0; i < 6; i++) m->slide_range[i] = (float)p[i];
    p = D("finger_range", 6); KG_NEED(p); for (int i = 0; i < 6; i++) m->finger_range[i] = (float)p[i];
    const double *l7i = D("l7_inertial", 7); KG_NEED(l7i);
    const double *fi = D("finger_inertial", 21); KG_NEED(fi);
    const double *ti = D("tip_inertial", 21); KG_NEED(ti);
    m->m_hand_total = (float)(l7i[3] + 3 * fi[3] + 3 * ti[3]);
    m->link_mass = (float)fi[3]; m->link_Izz = (float)fi[6];
    for (int i = 0; i < 3; i++) m->link_com[i] = (float)fi[i];
    for (int k = 0; k < 3; k++) {  // the kernels rely on identical proximal / distal inertials
        if (fi[7 * k + 3] != fi[3] || ti[7 * k + 3] != fi[3] || ti[7 * k + 6] != fi[6] || ti[7 * k] != fi[0]) return "finger inertials differ";
    }
    p = D("finger_pos", 9); KG_NEED(p); for (int i = 0; i < 9; i++) m->finger_pos[i] = (float)p[i];
    p = D("finger_quat", 12); KG_NEED(p);
    for (int k = 0; k < 3; k++) { double R[9]; q2m(p + 4 * k, R); for (int i = 0; i < 9; i++) m->finger_R[9 * k + i] = (float)R[i]; }
    p = D("tip_pos", 9); KG_NEED(p); for (int i = 0; i < 3; i++) m->tip_pos[i] = (float)p[i];
    const double *mg = D("mesh_geom", KG_NMESH * 15); KG_NEED(mg);
    float *gp[3] = {m->palm_gpos, m->prox_gpos, m->dist_gpos};
    float *gr[3] = {m->palm_gR, m->prox_gR, m->dist_gR};
    for (int k = 0; k < 3; k++) {
        const double *c = mg + 15 * k; double R[9];
        q2m(c + 3, R);
        for (int i = 0; i < 3; i++) { gp[k][i] = (float)c[i]; m->hand_size[3 * k + i] = (float)c[8 + i]; }
        for (int i = 0; i < 9; i++) gr[k][i] = (float)R[i];
        m->hand_rbound[k] = (float)c[7];
    }
    const double *oinv = D("obj_invw0", KG_NSHAPE * 4); KG_NEED(oinv);
    const int *shape_flags = I("shape_flags", nullptr);      // optional: bit 0 = _get_obj_size always swaps (lemon)
    const double *hm = D("hand_M0", 81); KG_NEED(hm);
    double tr = 0; for (int j = 0; j < 9; j++) tr += hm[10 * j];
    for (int s = 0; s < KG_NSHAPE; s++) {
        const double *c = mg + 15 * (3 + s); double R[9];
        KgShape &sh = m->shape[s];
        q2m(c + 3, R);
        for (int i = 0; i < 3; i++) { sh.gpos[i] = (float)c[i]; sh.size[i] = (float)c[8 + i]; }
        for (int i = 0; i < 9; i++) sh.gR[i] = (float)R[i];
        sh.rbound = (float)c[7];
        double in[3], Ic[9] = {0}, Io[9];
        for (int i = 0; i < 3; i++) in[i] = c[12 + i] * opt[5] / c[11];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) Ic[3 * i + j] += R[3 * i + k] * in[k] * R[3 * j + k];
        double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Io[3 * i + j] = Ic[3 * i + j] + opt[5] * ((i == j ? cc : 0) - c[i] * c[j]);
        const int ix[6] = {0, 4, 8, 1, 2, 5};
        for (int i = 0; i < 6; i++) { sh.Ic[i] = (float)Ic[ix[i]]; sh.Io[i] = (float)Io[ix[i]]; }
        sh.invw_tran = (float)oinv[4 * s];
        double t = tr + 3 * (opt[5] + opt[4]);
        for (int i = 0; i < 3; i++) t += Io[4 * i] + opt[4];
        sh.scale = (float)(1.0 / (t / 15.0 * 15.0));
        // object block of M (+ d I): [mp I, T; T^T, Io + (arm + d) I] with T = -m R [c]x, so T^T T = m^2 (|c|^2 I - c c^T) whatever R is
        for (int which = 0; which < 2; which++) {
            const double d = which ? opt[0] * opt[3] : 0.0, mp = opt[5] + opt[4] + d;
            double S[9], Si[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
                S[3 * i + j] = Io[3 * i + j] + (i == j ? opt[4] + d : 0.0) - opt[5] * opt[5] / mp * ((i == j ? cc : 0) - c[i] * c[j]);
            const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
            Si[0] = (S[4] * S[8] - S[5] * S[7]) / det; Si[1] = (S[2] * S[7] - S[1] * S[8]) / det; Si[2] = (S[1] * S[5] - S[2] * S[4]) / det;
            Si[4] = (S[0] * S[8] - S[2] * S[6]) / det; Si[5] = (S[2] * S[3] - S[0] * S[5]) / det; Si[8] = (S[0] * S[4] - S[1] * S[3]) / det;
            Si[3] = Si[1]; Si[6] = Si[2]; Si[7] = Si[5];
            for (int i = 0; i < 6; i++) sh.Sinv[which][i] = (float)Si[ix[i]];
        }
        // _get_obj_size (ENV:706-746), single-geom object
        double sz[3] = {c[8], c[9], c[10]};
        if (sz[2] == 0) { sz[2] = sz[1]; sz[1] = sz[0]; }
        double df[3] = {std::abs(sz[0] - sz[1]), std::abs(sz[1] - sz[2]), std::abs(sz[0] - sz[2])};
        int am = 0; if (df[1] < df[am]) am = 1; if (df[2] < df[am]) am = 2;
        if (am != 0 || (shape_flags && (shape_flags[s] & 1))) { double x = sz[0]; sz[0] = sz[2]; sz[2] = x; }      // (`'lemon' in self.filename`, ENV:721)
        sh.obs_size[0] = (float)(sz[0] > 0 ? sz[0] : 0); sh.obs_size[1] = (float)(sz[1] > 0 ? sz[1] : 0); sh.obs_size[2] = (float)sz[2];
    }
    p = D("site_pos", KG_NSITE * 3); KG_NEED(p); for (int i = 0; i < KG_NSITE * 3; i++) m->site_pos[i] = (float)p[i];
    p = D("site_zaxis", KG_NSITE * 3); KG_NEED(p); for (int i = 0; i < KG_NSITE * 3; i++) m->site_z[i] = (float)p[i];
    const int *ip;
    ip = I("site_body", nullptr); KG_NEED(ip); for (int i = 0; i < KG_NSITE; i++) m->site_body[i] = ip[i];
    ip = I("fod_site", nullptr); KG_NEED(ip); for (int i = 0; i < 12; i++) m->fod_site[i] = ip[i];
    const double *hb = D("hand_body_invw0", 14); KG_NEED(hb);
    m->body_invw[0] = 0;
    for (int g = 1; g < 8; g++) m->body_invw[g] = (float)hb[2 * (g - 1)];
    p = D("hand_dof_invw0", 9); KG_NEED(p); for (int i = 0; i < 9; i++) m->dof_invw[i] = (float)p[i];
    p = D("tendon_invw0", 3); KG_NEED(p); for (int i = 0; i < 3; i++) m->tendon_invw[i] = (float)p[i];
    const double *kv = D("act_kv", 9), *gear = D("act_gear", 9), *cr = D("act_ctrlrange", 18);
    KG_NEED(kv); KG_NEED(gear); KG_NEED(cr);
    m->act_kv_slide = (float)kv[0]; m->act_gear_slide = (float)gear[1]; m->act_kv_finger = (float)kv[6];
    m->ctrl_slide = (float)cr[1]; m->ctrl_finger = (float)cr[13];
    const double *pf = D("pair_friction", 8); KG_NEED(pf);
    // pair table: 0..7 plane vs geoms 1..8; 8..13 object vs geoms 2..7; 14 object vs palm; 15..29 dynamic
    for (int i = 0; i < 30; i++) m->pair_mu[i] = 1.0f;
    m->pair_mu[7] = (float)pf[0];
    for (int g = 2; g <= 7; g++) m->pair_mu[8 + g - 2] = (float)pf[g - 1];
    m->pair_mu[14] = (float)pf[7];
    // tables
    size_t n;
    ip = I("mesh_index", &n); KG_NEED(ip);
    std::memcpy(m->mesh, ip, sizeof(KgMesh) * KG_NMESH);
    const Sec *hv = find(blob, "hull_vert"); KG_NEED(hv);
    const double *hvd = (const double *)(blob + hv->offset);
    t->hull_vert.assign(hv->count / 3 * 4, 0.f);
    for (size_t i = 0; i < hv->count / 3; i++) for (int c = 0; c < 3; c++) t->hull_vert[4 * i + c] = (float)hvd[3 * i + c];
    ip = I("hull_adj_start", &n); KG_NEED(ip); t->adj_start.assign(ip, ip + n);
    ip = I("hull_adj", &n); KG_NEED(ip); t->adj.resize(n); for (size_t i = 0; i < n; i++) t->adj[i] = (unsigned short)ip[i];
    // chunked neighbour lists: a chunk = 8 neighbour ids (adj8, 16 bytes) + their vertex records (nbr8, 128 bytes), padded
    // with the vertex itself.  The FIRST chunk of vertex gv (global index) is chunk gv, so it is addressed from the vertex
    // index alone and loads together with the vertex record; vertices with more than 8 neighbours continue in overflow
    // chunks appended behind, referenced by hull_vert.w = bits of (first overflow chunk | n overflow chunks << 24).
    {
        size_t nvtot = 0;
        for (int k = 0; k < KG_NMESH; k++) nvtot = std::max(nvtot, (size_t)(m->mesh[k].vert0 + m->mesh[k].nvert));
        t->adj8.assign(8 * nvtot, 0);
        for (int k = 0; k < KG_NMESH; k++) {
            const KgMesh &mk = m->mesh[k];
            for (int v = 0; v < mk.nvert; v++) {
                const size_t gv = (size_t)mk.vert0 + v;
                const int a0 = t->adj_start[gv], a1 = t->adj_start[gv + 1];
                for (int i = 0; i < 8; i++) t->adj8[8 * gv + i] = a0 + i < a1 ? t->adj[a0 + i] : (unsigned short)v;
                const uint32_t novf = a1 - a0 > 8 ? (uint32_t)((a1 - a0 - 8 + 7) / 8) : 0u, first = (uint32_t)(t->adj8.size() / 8);
                for (uint32_t i = 0; i < novf * 8; i++) t->adj8.push_back(a0 + 8 + (int)i < a1 ? t->adj[a0 + 8 + i] : (unsigned short)v);
                const uint32_t info = (novf ? first : 0u) | (novf << 24);
                std::memcpy(&t->hull_vert[4 * gv + 3], &info, 4);
            }
        }
        t->nbr8.assign(4 * t->adj8.size(), 0.f);
        for (int k = 0; k < KG_NMESH; k++) {
            const KgMesh &mk = m->mesh[k];
            for (int v = 0; v < mk.nvert; v++) {
                const size_t gv = (size_t)mk.vert0 + v;
                uint32_t info; std::memcpy(&info, &t->hull_vert[4 * gv + 3], 4);
                const uint32_t first = info & 0xffffffu, novf = info >> 24;
                for (uint32_t ch = 0; ch <= novf; ch++) {
                    const size_t ci = ch == 0 ? gv : (size_t)first + ch - 1;
                    for (int i = 0; i < 8; i++) {
                        const int u = t->adj8[8 * ci + i];
                        for (int c = 0; c < 4; c++) t->nbr8[4 * (8 * ci + i) + c] = t->hull_vert[4 * ((size_t)mk.vert0 + u) + c];
                    }
                }
            }
        }
    }
    for (int k = 0; k < KG_NMESH; k++) {
        const KgMesh &mk = m->mesh[k];
        m->hpad0[k] = (int)(t->hull_pad.size() / 4);
        m->hnpad[k] = (mk.nvert + 31) / 32 * 32;
        for (int v = 0; v < m->hnpad[k]; v++) {
            const float *src = &t->hull_vert[4 * (size_t)(mk.vert0 + (v < mk.nvert ? v : 0))];
            for (int c = 0; c < 3; c++) t->hull_pad.push_back(src[c]);
            t->hull_pad.push_back(0.f);
        }
    }
    // ---- octant candidate lists.  On a convex polytope vertex v maximises d.x iff d.(v - u) >= 0 for every hull
    // neighbour u, so v is a candidate of octant o iff the cone {d : d.(v-u) >= 0 for all u, s_c d_c >= 0} is more than {0}.
    // A non-trivial cone in R^3 contains one of the directions +-(n_i x n_j) of its constraint normals (or a normal
    // itself), so feasibility is a finite check (in double, with a tolerance that can only ADD candidates).
    for (int k = 0; k < KG_NMESH; k++) {
        const KgMesh &mk = m->mesh[k];
        std::vector<std::vector<int>> lists(8);
        std::vector<double> nrm;
        for (int v = 0; v < mk.nvert; v++) {
            const float *pv = &t->hull_vert[4 * (size_t)(mk.vert0 + v)];
            nrm.clear();
            for (int a = t->adj_start[mk.vert0 + v]; a < t->adj_start[mk.vert0 + v + 1]; a++) {
                const float *pu = &t->hull_vert[4 * (size_t)(mk.vert0 + t->adj[a])];
                double e[3] = {(double)pv[0] - pu[0], (double)pv[1] - pu[1], (double)pv[2] - pu[2]};
                double l = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                if (l > 0) { nrm.push_back(e[0] / l); nrm.push_back(e[1] / l); nrm.push_back(e[2] / l); }
            }
            const size_t nedge = nrm.size() / 3;
            for (int o = 0; o < 8; o++) {
                nrm.resize(3 * nedge);
                for (int c = 0; c < 3; c++) { double s3[3] = {0, 0, 0}; s3[c] = ((o >> c) & 1) ? -1.0 : 1.0; nrm.insert(nrm.end(), s3, s3 + 3); }
                const size_t nc = nrm.size() / 3;
                auto ok = [&](const double *d) {
                    double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    if (l < 1e-12) return false;
                    for (size_t i = 0; i < nc; i++) if (nrm[3 * i] * d[0] + nrm[3 * i + 1] * d[1] + nrm[3 * i + 2] * d[2] < -1e-6 * l) return false;
                    return true;
                };
                bool feas = false;
                for (size_t i = 0; i < nc && !feas; i++) {
                    const double *a = &nrm[3 * i];
                    double na[3] = {-a[0], -a[1], -a[2]};
                    if (ok(a) || ok(na)) { feas = true; break; }
                    for (size_t j = i + 1; j < nc && !feas; j++) {
                        const double *b = &nrm[3 * j];
                        double x[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                        double nx[3] = {-x[0], -x[1], -x[2]};
                        if (ok(x) || ok(nx)) feas = true;
                    }
                }
                if (feas || nedge == 0) lists[o].push_back(v);
            }
        }
        for (int o = 0; o < 8; o++) {
            if (lists[o].empty()) lists[o].push_back(0);
            m->oct0[k][o] = (int)(t->hull_oct.size() / 4);
            m->octn[k][o] = ((int)lists[o].size() + 31) / 32 * 32;
            for (int i = 0; i < m->octn[k][o]; i++) {
                const int v = lists[o][i < (int)lists[o].size() ? i : 0];
                const float *src = &t->hull_vert[4 * (size_t)(mk.vert0 + v)];
                for (int c = 0; c < 3; c++) t->hull_oct.push_back(src[c]);
                float w; std::memcpy(&w, &v, 4); t->hull_oct.push_back(w);
            }
        }
        // self-check: on a fixed set of directions the list scan must return the vertex of the exhaustive scan
        uint32_t rs = 12345u + 977u * (uint32_t)k;
        for (int it = 0; it < 3000; it++) {
            float d[3];
            for (int c = 0; c < 3; c++) { rs = rs * 1664525u + 1013904223u; d[c] = (float)((rs >> 8) & 0xffff) / 32768.f - 1.f; }
            if (it < 27) { d[0] = (float)(it % 3 - 1); d[1] = (float)((it / 3) % 3 - 1); d[2] = (float)(it / 9 - 1); }   // axes, face and edge diagonals
            if (d[0] == 0.f && d[1] == 0.f && d[2] == 0.f) continue;
            const int o = (d[0] < 0) | ((d[1] < 0) << 1) | ((d[2] < 0) << 2);
            float best = -3.4e38f, bo = -3.4e38f; int bi = 0, oi = 0;
            for (int v = 0; v < mk.nvert; v++) {
                const float *pv = &t->hull_vert[4 * (size_t)(mk.vert0 + v)];
                float dv = pv[0] * d[0] + pv[1] * d[1] + pv[2] * d[2];
                if (dv > best) { best = dv; bi = v; }
            }
            for (int i = 0; i < m->octn[k][o]; i++) {
                const float *pv = &t->hull_oct[4 * (size_t)(m->oct0[k][o] + i)];
                float dv = pv[0] * d[0] + pv[1] * d[1] + pv[2] * d[2];
                int v; std::memcpy(&v, pv + 3, 4);
                if (dv > bo || (dv == bo && v < oi)) { bo = dv; oi = v; }
            }
            if (oi != bi) return "octant support table disagrees with the exhaustive scan";
        }
    }
    const float *fp;
    fp = F("tri_vert", &n); KG_NEED(fp); t->tri.assign(fp, fp + n);
    {
        size_t nlo = 0, nhi = 0, nmeta = 0;
        const float *lo = F("bvh_lo", &nlo), *hi = F("bvh_hi", &nhi);
        const int *meta = I("bvh_meta", &nmeta);
        KG_NEED(lo); KG_NEED(hi); KG_NEED(meta);
        size_t nn = nlo / 3;
        if (nhi != nlo || nmeta != 2 * nn) return "bvh sections disagree";
        t->bvh_node.assign(8 * nn, 0.f);
        for (size_t i = 0; i < nn; i++) {
            for (int c = 0; c < 3; c++) { t->bvh_node[8 * i + c] = lo[3 * i + c]; t->bvh_node[8 * i + 4 + c] = hi[3 * i + c]; }
            std::memcpy(&t->bvh_node[8 * i + 3], &meta[2 * i], 4);
            std::memcpy(&t->bvh_node[8 * i + 7], &meta[2 * i + 1], 4);
        }
    }
    fp = F("pose_train", &n); KG_NEED(fp); t->pose_train.assign(fp, fp + n); m->pose_rows_train = (int)(n / (KG_NSHAPE * 3 * 9));
    fp = F("pose_test", &n); KG_NEED(fp); t->pose_test.assign(fp, fp + n); m->pose_rows_test = (int)(n / (KG_NSHAPE * 3 * 9));
#undef KG_NEED
    return "";
}
