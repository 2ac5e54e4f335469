// kg_core.cuh -- the batched grasp-env substep, ONE WARP PER ENVIRONMENT (fp32).
//
// Replaces the per-substep call chain kinova_gripper_env.py:1516-1535 -> mujoco_py MjSim.step() ->
// mj_step for this fixed model.  Everything is written as a sequence of
// warp-wide PHASES over a per-warp shared-memory workspace (KgWS):
//
//     KG_PHASE(l) { ... body executed by lane l ... }  KG_SYNC();
//
// A phase body may read anything written in earlier phases and must write only lane-owned
// locations.  On the GPU a phase is straight-line code for the calling lane followed by
// __syncwarp().  The same source also compiles for the host (tests/emu, TEST-ONLY: a debugging
// aid that executes the lanes of a phase one after the other; it is never linked into the package),
// which is how the algorithm is checked against the fp64 oracle inside the GPU-less container.
//
// Formulation notes (how this differs from the literal pipeline the oracle restates):
//   * link_7 only translates, both hinges of a finger share one world-fixed axis  ->  CRBA/RNE are
//     closed-form per finger (planar 2R chain on a translating base), the free body is closed-form;
//   * hull support is hill-climbing on the hull graph, warm-started per pair across substeps;
//   * the 4 pyramid rows of a contact are never stored: 3 frame rows J_n, J_t1, J_t2 are, and
//     row_k = J_n +- mu J_t is expanded on the fly;
//   * the exact line search brackets the root of the piecewise-linear phi'(alpha) with safeguarded
//     Newton steps evaluated warp-wide.
#pragma once
#include "kg_model.h"
#include <math.h>
#include <float.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__)
#define KG_DEV __device__ __forceinline__
#define KG_NOINL __device__ __noinline__
#define KG_PHASE(l) for (int l = kg_lane, kg_once_ = 0; kg_once_ < 1; ++kg_once_)
#define KG_SYNC() __syncwarp()
// a per-lane variable that lives across phases: a register array on the GPU, one slot per lane in the host emulation
#define KG_LANEVAR(type, name, n) type name[n]
#define KG_LV(name, l) name
// a per-lane partial and its warp total: registers and shuffles on the GPU (no trip through shared memory, no warp sync
// between writing and summing), one slot per lane in the host emulation; same butterfly order as kg_total
#define KG_PARTIAL(name) float name = 0.f
#define KG_PARTIAL_I(name) int name = 0
#define KG_PUT(name, l, v) name = (v)
#define KG_SUM(name, l) kg_total_reg(name)
#define KG_SUM_I(name, l) __reduce_add_sync(0xffffffffu, name)
#define KG_LDG4(p) __ldg((const float4 *)(p))
#define KG_LDGU4(p) __ldg((const uint4 *)(p))
#define KG_F2U(x) __float_as_uint(x)
#define KG_LD4S(p) (*reinterpret_cast<const float4 *>(p))
// Block-wide rendezvous between pipeline stages.  Functionally unnecessary (warps never share data); it keeps the
// 8 warps of a CTA inside the same ~10-20 KB code region so that they share instruction-cache lines (the whole
// substep is ~100 KB of SASS, far beyond the 32 KB L1.5 I-cache; ncu showed stall_no_instruction dominating).
#define KG_STAGE(i) do { if ((WS::STAGE_MASK >> (i)) & 1) __syncthreads(); } while (0)
#else
#define KG_STAGE(i) ((void)0)
#if defined(__CUDACC__)
#define KG_DEV __host__ __device__ inline
#define KG_NOINL __host__ __device__
#else
#define KG_DEV inline
#define KG_NOINL inline
#endif
// host emulation: the lanes of a phase run one after the other, forwards or (race check) backwards.  A phase that
// reads something another lane writes in the same phase gives different results in the two orders.
static int kg_emu_lane_order = 0;
#define KG_PHASE(l) for (int kg_i_ = 0, l = kg_emu_lane_order ? 31 : 0; kg_i_ < 32; ++kg_i_, l += kg_emu_lane_order ? -1 : 1)
#define KG_SYNC() ((void)0)
#define KG_LANEVAR(type, name, n) type name[32][n]
#define KG_LV(name, l) name[l]
#define KG_PARTIAL(name) float name[32]
#define KG_PARTIAL_I(name) int name[32]
#define KG_PUT(name, l, v) name[l] = (v)
#define KG_SUM(name, l) kg_total(name, l)
#define KG_SUM_I(name, l) kg_total_int(name, l)
struct kg_float4 { float x, y, z, w; };
struct kg_uint4 { unsigned x, y, z, w; };
#define KG_LDG4(p) (*(const kg_float4 *)(p))
#define KG_LDGU4(p) (*(const kg_uint4 *)(p))
static inline unsigned KG_F2U(float x) { unsigned u; memcpy(&u, &x, 4); return u; }
#define KG_LD4S(p) (*(const kg_float4 *)(p))
#endif

#ifndef KG_STAGE_MASK
#define KG_STAGE_MASK 0x1f   // bit i: rendezvous before stage i (0 kinematics, 1 collision, 2 rows, 3 solve, 4 Euler)
#endif
#define KG_NV 15
#define KG_MS 16                       // row stride of 15x15 matrices
#ifndef KG_MAXCON
#define KG_MAXCON 23                   // contact groups per env in the full workspace (overflow is counted and dropped)
#endif
#ifndef KG_MAXCON_LIGHT
#define KG_MAXCON_LIGHT 8              // ... in the light workspace: envs with more contacts go to the full-workspace kernel
#endif
#ifndef KG_JS
#define KG_JS 20
#endif                                 // row stride of J: rows are 16-byte aligned (128-bit row loads; the eight rows a
                                       // quarter-warp reads together fall into disjoint bank groups: 20 b mod 32 = 0,20,8,28,16,4,24,12)
#define KG_NSTAGE 54                   // contact staging slots: 8 plane lanes x 4 + 22 hull lanes
#define KG_NOBS 82
#define KG_MPR_TOL 1e-6f
#define KG_MPR_ITER 50
#ifndef KG_NEWTON_MAXIT
#define KG_NEWTON_MAXIT 20             // iteration cap; the loop exits early on the (fp32-aware) 1e-8 tests
#endif
#define KG_LS_MAXIT 12
#define KG_NPROF 32

// Profiling builds (-DKG_PROF): KG_PROBE(i) charges the cycles since the previous probe of this warp to slot i,
// KG_COUNT(i, n) adds n to counter slot i (any lane).  Slots: kg_get_prof in include/kg.h.
#if defined(KG_PROF) && defined(__CUDA_ARCH__)
#define KG_PROBE(i) do { long long t_ = clock64(); if (kg_lane == 0) { ws.prof[i] += t_ - ws.prof_t0; ws.prof_t0 = t_; } } while (0)
#define KG_COUNT(i, n) atomicAdd((unsigned long long *)&ws.prof[i], (unsigned long long)(n))
#else
#define KG_PROBE(i) ((void)0)
#define KG_COUNT(i, n) ((void)0)
#endif

// obs staging aliases the constraint Jacobian storage (J is dead once the last substep's solve is over)
#define KG_WSOBS(ws) ((ws).J)
#define KG_RANGE(ws) ((ws).range)
#define KG_GROW(g) ((g) < 9 ? (g) : 9 + 3 * ((g) - 9))      // first J row of constraint group g
// object hull vertices a narrowphase workspace can STAGE (padded to a multiple of 32).  Only builds without KG_MPR_L1 stage
// the object's hull (the product reads the hull tables through the read-only L1 path and has no such limit: the lemon's hull
// has 2434 vertices); the host emulation stages, so it gets room for every shape of the model card.
#ifdef KG_EMU
#define KG_MAXOBJVERT 2560
#else
#define KG_MAXOBJVERT 160
#endif
// Per-environment workspace in shared memory.  An RL step is a WAVEFRONT of kernels (kg_api.cu): the per-env kernels
// keep one KgWS per warp; between kernels the CARRY part of it lives in global memory (one KgCarry record per env,
// env-major so that a warp moves it with coalesced 128-bit accesses).  Carry = everything a later kernel of the same
// RL step reads: state, episode constants, the kinematics of the current substep, collision caches, staged contacts.
template <int MAXCON_, int STAGE_MASK_>
struct alignas(16) KgWST {
    static constexpr int MAXCON = MAXCON_;               // contact groups this workspace has rows for
    static constexpr int NBASE = 9 + 3 * MAXCON_;        // 3 tendon rows, 6 limit rows, 3 frame rows per contact
    static constexpr int STAGE_MASK = STAGE_MASK_;       // CTA-wide rendezvous before the stages in this mask (KG_STAGE)
    // ======================================================================== carry section A (contiguous)
    // ---- state
    float qpos[16], qvel[16], qwarm[16], ctrl[12], sens_q[12];
    // ---- per-episode constants
    float Rh[9], B[3][9], fo[3][3], sax[3][3];
    int shape, ntask;
    // ---- kinematics
    float hp[3], p1[3][3], p2[3][3], c1[3][3], c2[3][3];
    float gpos[9][3], gmat[9][9];
    float oR[9];
    // ---- collision bookkeeping
    int cnt[32];
    // ---- statistics
    int stat_mpr, stat_newton, stat_overflow, stat_pad;
    // ======================================================================== end of carry section A
    // ---- dynamics
    alignas(16) float M[KG_NV * KG_MS];
    // the Newton-solver scratch and the contact staging of the collision stage are never live together; the staging
    // (carry section B) is what the narrowphase kernel fills in global memory.  The rangefinder readings only exist
    // between kg_rays and kg_obs, at the very end of an RL step.
    union {
        struct { float L[KG_NV * KG_MS], bdot[NBASE + 2], bdir[NBASE + 2], bw[NBASE + 2], range[20]; };
        struct { float st_dist[KG_NSTAGE], st_pos[KG_NSTAGE][3], st_n[KG_NSTAGE][3]; unsigned char st_g1[KG_NSTAGE], st_g2[KG_NSTAGE], st_pad[16]; };
    };
    alignas(16) float bias[16], fs[16], qas[16], qacc[16], Mdq[16], grad[16], dir[16], Md[16], ftot[16];
    float sbias[3][3];
    // ---- constraint groups: slot g<3 tendon, 3..8 limits, 9.. contacts.  Their per-group scalars (type, aref, D, mu,
    // staging slot) live in the 5 pad words of the group's J rows (kg_g* accessors below).
    int ngroup, ncon_total;
    // J rows have 15 entries at stride KG_JS = 20.  The same storage is, at other times of a substep: the COLLISION CACHE
    // of the env (carry section C: warm-start hull vertices and cached separating axes per pair lane; only the collision
    // stage reads it, long after the solve has finished with J), the obs staging, the rangefinder scratch.
    union {
        alignas(16) float J[NBASE * KG_JS];
        struct {
            unsigned short cv1[32], cv2[32];   // warm-start hull vertices per pair lane
            float sep[32][3];                  // cached separating axis per hull-pair lane (valid while sepok[l] != 0)
            int sepok[32];
            // separation BUDGET per pair lane: the gap measured along the cached axis (hull pairs) or above the contact margin
            // (plane lanes) when the pair was last tested, minus an upper bound of how far the two geoms' surfaces have moved
            // since (cpos = geom centres at the last collision stage).  While it stays positive the pair cannot have come
            // into contact range, so the support queries of the re-test are skipped -- the contact set is unchanged by this.
            float slack[32];
            float cpos[9][3], cpad;
        };
    };
    float red[32], red2[32];
    // ---- solver scalars (written by lane 0, read by all after a sync)
    float s_alpha, s_cost, s_flag, s_lo, s_hi, s_glo, s_ghi, s_cost2;
    int s_iter, s_done, s_exact;
    // ---- closed-form M^-1 scratch (per finger lane)
    float mv_S[3][6], mv_r[3][3], mv_y[3][2], mv_i[3][3], mv_xs[4], mv_red3[32];
#ifdef KG_PROF
    long long prof[KG_NPROF], prof_t0;  // see KG_PROBE / kg_get_prof
#endif
};
// Two instances: the FULL workspace (KG_MAXCON contact groups; the host emulation and the debug entry points use only
// this one) and the LIGHT one the product kernels run first: fewer Jacobian rows -> more warps per SM.  The carry
// sections do not depend on the row count, so a record written through one workspace loads into the other.
typedef KgWST<KG_MAXCON, KG_STAGE_MASK> KgWS;
typedef KgWST<KG_MAXCON_LIGHT, KG_STAGE_MASK> KgWSLight;
// The HEAVY LANE's workspace: rows for every contact this model can have at once (8 plane lanes x 4 + 22 hull pairs = 54,
// i.e. 63 constraint groups, two per lane), no stage rendezvous (its warps loop over envs independently).  Whatever exceeds
// the workspace of the kernel it runs in -- more than KG_MAXCON_LIGHT contacts in split mode, more than KG_MAXCON in full
// mode -- is run there, so no contact is ever dropped (MuJoCo never drops one for this model: nconmax >> 54).
#define KG_MAXCON_ALL 54
typedef KgWST<KG_MAXCON_ALL, 0> KgWSNoSync;
// the collision caches: contiguous in the workspace (aliased onto J), kept per env across RL steps (a cached axis is only
// ever RE-TESTED, a cached vertex is only a starting point, so stale entries after kg_set_state cost time, not correctness)
#define KG_CCACHE_OFF   (offsetof(KgWS, cv1))
#define KG_CCACHE_BYTES (offsetof(KgWS, cpad) + sizeof(float) - offsetof(KgWS, cv1))
// carry record of an env in global memory: [section A: state, constants, kinematics, contact counts][section C: collision
// cache][section B: staged contacts]
#define KG_CARRY_A_BYTES (offsetof(KgWS, M))
#define KG_CARRY_B_OFF   (offsetof(KgWS, st_dist))
#define KG_CARRY_B_BYTES ((sizeof(float) * 7 * KG_NSTAGE + 2 * KG_NSTAGE + 15) / 16 * 16)
#define KG_CARRY_BYTES   ((KG_CARRY_A_BYTES + KG_CCACHE_BYTES + KG_CARRY_B_BYTES + 127) / 128 * 128)
// offsets of the fields the narrowphase kernel touches inside a carry record
#define KG_COFF(field)   (offsetof(KgWS, field))
#define KG_COFF_C(field) (KG_CARRY_A_BYTES + offsetof(KgWS, field) - KG_CCACHE_OFF)
#define KG_COFF_B(field) (KG_CARRY_A_BYTES + KG_CCACHE_BYTES + offsetof(KgWS, field) - KG_CARRY_B_OFF)

static_assert(offsetof(KgWS, M) % 16 == 0 && offsetof(KgWS, L) % 16 == 0 && offsetof(KgWS, J) % 16 == 0 && sizeof(KgWS) % 16 == 0, "128-bit access needs 16-byte aligned members");
static_assert(KG_NOBS * 4 <= KgWSLight::NBASE * KG_JS * 4, "obs staging must fit into J");
static_assert(KG_CARRY_B_OFF % 16 == 0 && offsetof(KgWS, st_g2) + KG_NSTAGE - KG_CARRY_B_OFF <= KG_CARRY_B_BYTES, "carry section B");
static_assert(KG_CARRY_B_BYTES <= sizeof(float) * 7 * KG_NSTAGE + 2 * KG_NSTAGE + 16, "carry section B must stay inside the union (st_pad absorbs the 16-byte round-up)");
static_assert(offsetof(KgWSLight, M) == offsetof(KgWS, M) && offsetof(KgWSLight, st_dist) == offsetof(KgWS, st_dist) && offsetof(KgWSLight, st_g2) == offsetof(KgWS, st_g2),
              "carry sections must sit at the same offsets in both workspaces");
static_assert(KG_CCACHE_OFF % 16 == 0 && KG_CCACHE_BYTES % 16 == 0 && KG_CCACHE_BYTES == 880 && KG_CARRY_A_BYTES % 16 == 0, "collision cache block");
static_assert(offsetof(KgWS, cv1) == offsetof(KgWS, J) && offsetof(KgWSLight, cv1) == offsetof(KgWSLight, J) && KG_CCACHE_BYTES <= KgWSLight::NBASE * KG_JS * 4, "collision cache aliases J");
static_assert(offsetof(KgWSLight, J) % 16 == 0 && sizeof(KgWSLight) % 16 == 0, "128-bit access needs 16-byte aligned members");

// Per-group scalars in the pad words 15..19 of the group's J rows: row b = KG_GROW(g): [aref0 | aref1 (tendon / limit rows:
// D * active of the current Newton iterate) | D | type | mu]; contact groups also use row b + 1: [aref2 | aref3 | slot | A | B1]
// and row b + 2: [C1 | B2 | C2] -- the Hessian coefficients of the group's active rows at the current Newton iterate (KG_GH).
// type: 0 none, 1 equality, 2 limit, 3 contact; type and the staging slot are small integers stored as floats.
#define KG_GP(ws, g, k) ((ws).J[(KG_GROW(g) + (k) / 5) * KG_JS + 15 + (k) % 5])
#define KG_GH(ws, g, k) KG_GP(ws, g, 8 + (k))      // k = 0..4: A, B1, C1, B2, C2 (contact groups only)
#define KG_GAREF(ws, g, k) KG_GP(ws, g, (k) < 2 ? (k) : (k) + 3)
#define KG_GD(ws, g)       KG_GP(ws, g, 2)
#define KG_GMU(ws, g)      KG_GP(ws, g, 4)
template <class WS> KG_DEV int kg_gtype(const WS &ws, int g) { return g < ws.ngroup ? (int)KG_GP(ws, g, 3) : 0; }
template <class WS> KG_DEV int kg_gslot(const WS &ws, int g) { return (int)KG_GP(ws, g, 7); }

// Workspace of the narrowphase (MPR) kernel: the geom frames of one env, the object's hull, the finger hulls of the CTA.
struct alignas(16) KgMprWS {
    // frames of the task's two geoms, laid out for 128-bit loads: centre (x, y, z, -) and the three rows of R (x, y, z, -)
    alignas(16) float gp[2][4], gr[2][12];
#ifndef KG_MPR_L1
    alignas(16) float objhull[KG_MAXOBJVERT * 4];
#endif
    const float *fhull;
    int shape, hull_shape;
#ifdef KG_PROF
    long long prof[KG_NPROF];
#endif
};

// ------------------------------------------------------------------------------------------------ small math
KG_DEV float kg_dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
KG_DEV void kg_cross(float *r, const float *a, const float *b) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
KG_DEV void kg_sub(float *r, const float *a, const float *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
KG_DEV void kg_add(float *r, const float *a, const float *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
KG_DEV void kg_addscl(float *r, const float *a, const float *b, float s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
KG_DEV void kg_copy(float *r, const float *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
// direction-only normalisation for the MPR search directions (the result steers support queries; rsqrt's 2 ulp is
// far below the 1e-6 portal tolerance).  Output normals and frames keep the exact kg_normalize below.
KG_DEV void kg_normalize_dir(float *a) {
    float n2 = kg_dot(a, a);
    if (n2 < 1e-38f) { a[0] = 1; a[1] = 0; a[2] = 0; return; }
#if defined(__CUDA_ARCH__)
    float inv = rsqrtf(n2);
#else
    float inv = 1.0f / sqrtf(n2);
#endif
    a[0] *= inv; a[1] *= inv; a[2] *= inv;
}
KG_DEV float kg_normalize(float *a) {
    float n = sqrtf(kg_dot(a, a));
    if (n < 1e-30f) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
    float inv = 1.0f / n;
    a[0] *= inv; a[1] *= inv; a[2] *= inv;
    return n;
}
KG_DEV void kg_mulmv(float *r, const float *M, const float *v) {
    float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    float y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    float z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
KG_DEV void kg_multmv(float *r, const float *M, const float *v) {
    float x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    float y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    float z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
KG_DEV void kg_mulmm(float *r, const float *A, const float *Bm) {
    float t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * Bm[j] + A[3 * i + 1] * Bm[3 + j] + A[3 * i + 2] * Bm[6 + j];
    for (int i = 0; i < 9; i++) r[i] = t[i];
}
KG_DEV void kg_quat2mat(float *M, const float *q) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
    M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
    M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}

// 16 consecutive floats from a 16-byte aligned shared-memory address with four 128-bit loads
KG_DEV void kg_ld16(const float *p, float *v) {
#if defined(__CUDA_ARCH__)
    const float4 *q = reinterpret_cast<const float4 *>(p);
    float4 a = q[0], b = q[1], c = q[2], d = q[3];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
#else
    for (int i = 0; i < 16; i++) v[i] = p[i];
#endif
}

// 8 consecutive floats from / to a 16-byte aligned shared-memory address with two 128-bit accesses
KG_DEV void kg_ld8(const float *p, float *v) {
#if defined(__CUDA_ARCH__)
    const float4 *q = reinterpret_cast<const float4 *>(p);
    float4 a = q[0], b = q[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
    for (int i = 0; i < 8; i++) v[i] = p[i];
#endif
}
KG_DEV void kg_st8(float *p, const float *v) {
#if defined(__CUDA_ARCH__)
    float4 *q = reinterpret_cast<float4 *>(p);
    q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]);
#else
    for (int i = 0; i < 8; i++) p[i] = v[i];
#endif
}

// warp total / exclusive prefix of a per-lane int array written in the previous phase (every lane calls these)
KG_DEV int kg_total_int(const int *v, int l) {
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xffffffffu, v[l]);
#else
    (void)l;
    int s = 0;
    for (int i = 0; i < 32; i++) s += v[i];
    return s;
#endif
}
KG_DEV void kg_prefix_int(const int *v, int l, int &off, int &total) {
#if defined(__CUDA_ARCH__)
    const int x = v[l];
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += y; }
    off = inc - x;
    total = __shfl_sync(0xffffffffu, inc, 31);
#else
    off = 0; total = 0;
    for (int k = 0; k < 32; k++) { int c = v[k]; if (k < l) off += c; total += c; }
#endif
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float kg_total_reg(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
#endif
// warp total of ws.red[] (every lane wrote its partial in the previous phase)
KG_DEV float kg_total(const float *red, int l) {
#if defined(__CUDA_ARCH__)
    float v = red[l];
    v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
#else
    (void)l;
    float p[32];
    for (int i = 0; i < 32; i++) p[i] = red[i];
    for (int o = 16; o >= 1; o >>= 1) { float q[32]; for (int i = 0; i < 32; i++) q[i] = p[i] + p[i ^ o]; for (int i = 0; i < 32; i++) p[i] = q[i]; }
    return p[0];
#endif
}

// ------------------------------------------------------------------------------------------------ episode constants
// hand_quat is what the reference writes into the XML at every reset (ENV:851-881); here it is data.
template <class WS> KG_DEV void kg_episode_constants(const KgModel &m, WS &ws, const float *hand_quat, int shape, int kg_lane) {
    KG_PHASE(l) {
        if (l == 0) {
            kg_quat2mat(ws.Rh, hand_quat);
            ws.shape = shape;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l < 3) {
            kg_mulmm(ws.B[l], ws.Rh, m.finger_R + 9 * l);
            kg_mulmv(ws.fo[l], ws.Rh, m.finger_pos + 3 * l);
            kg_mulmv(ws.sax[l], ws.Rh, m.slide_axes + 3 * l);
        }
        if (l == 3) kg_mulmm(ws.gmat[1], ws.Rh, m.palm_gR);
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P1 kinematics
template <class WS> KG_DEV void kg_kinematics(const KgModel &m, WS &ws, int kg_lane) {
    KG_PHASE(l) {
        float hp[3];
        for (int c = 0; c < 3; c++) hp[c] = m.l7_pos[c] + ws.sax[0][c] * ws.qpos[0] + ws.sax[1][c] * ws.qpos[1] + ws.sax[2][c] * ws.qpos[2];
        if (l < 6) {
            // one lane per finger LINK (0..2 proximal, 3..5 distal of finger k), the same instruction stream for both kinds:
            // R = the link's rotation (R1 or R2), P = its joint position (p1 or p2)
            const int k = l < 3 ? l : l - 3;
            const bool distal = l >= 3;
            const float *Bk = ws.B[k];
            float q1 = ws.qpos[3 + 2 * k], q2 = ws.qpos[4 + 2 * k];
            float s1, c1, s2, c2;
            sincosf(q1, &s1, &c1);
            sincosf(distal ? q1 + q2 : q1, &s2, &c2);
            float R1[9], R[9];
            for (int r = 0; r < 3; r++) {
                float b0 = Bk[3 * r], b1 = Bk[3 * r + 1], b2 = Bk[3 * r + 2];
                R1[3 * r] = c1 * b0 + s1 * b1; R1[3 * r + 1] = -s1 * b0 + c1 * b1; R1[3 * r + 2] = b2;
                R[3 * r] = c2 * b0 + s2 * b1; R[3 * r + 1] = -s2 * b0 + c2 * b1; R[3 * r + 2] = b2;
            }
            float t[3], p1[3], p2[3];
            kg_add(p1, hp, ws.fo[k]);
            kg_mulmv(t, R1, m.tip_pos); kg_add(p2, p1, t);
            if (!distal) { kg_copy(ws.p1[k], p1); kg_copy(ws.p2[k], p2); }
            float Pv[3] = {distal ? p2[0] : p1[0], distal ? p2[1] : p1[1], distal ? p2[2] : p1[2]};
            kg_mulmv(t, R, m.link_com); kg_add(distal ? ws.c2[k] : ws.c1[k], Pv, t);
            kg_mulmv(t, R, distal ? m.dist_gpos : m.prox_gpos); kg_add(ws.gpos[2 + 2 * k + (distal ? 1 : 0)], Pv, t);
            kg_mulmm(ws.gmat[2 + 2 * k + (distal ? 1 : 0)], R, distal ? m.dist_gR : m.prox_gR);
        } else if (l == 6) {
            float q[4] = {ws.qpos[12], ws.qpos[13], ws.qpos[14], ws.qpos[15]};
            float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            float inv = 1.0f / n;
            for (int c = 0; c < 4; c++) q[c] *= inv;
            kg_quat2mat(ws.oR, q);
            const KgShape &sh = m.shape[ws.shape];
            float t[3];
            kg_mulmv(t, ws.oR, sh.gpos);
            for (int c = 0; c < 3; c++) ws.gpos[8][c] = ws.qpos[9 + c] + t[c];
            kg_mulmm(ws.gmat[8], ws.oR, sh.gR);
        } else if (l == 7) {
            float t[3];
            kg_mulmv(t, ws.Rh, m.palm_gpos);
            kg_add(ws.gpos[1], hp, t);
            kg_copy(ws.hp, hp);
        } else if (l >= 8 && l < 17) {
            const int i = l - 8;
            ws.sens_q[i] = i < 3 ? ws.qpos[i] : (i < 6 ? ws.qpos[3 + 2 * (i - 3)] : ws.qpos[4 + 2 * (i - 6)]);
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P2 + P3: M, bias, smooth force
template <class WS> KG_DEV void kg_dynamics(const KgModel &m, WS &ws, int kg_lane) {
    KG_PHASE(l) {
        for (int i = l; i < KG_NV * KG_MS; i += 32) ws.M[i] = 0.f;
    }
    KG_SYNC();
    KG_PHASE(l) {
        const float g = m.gravity;
        if (l < 3) {
            const float mk = m.link_mass, Izz = m.link_Izz;
            float a[3] = {ws.B[l][2], ws.B[l][5], ws.B[l][8]};
            float r1[3], r2[3], r22[3], r12[3], d11[3], d12[3], d22[3];
            kg_sub(r1, ws.c1[l], ws.p1[l]); kg_sub(r2, ws.c2[l], ws.p1[l]); kg_sub(r22, ws.c2[l], ws.p2[l]);
            kg_sub(r12, ws.p2[l], ws.p1[l]);
            kg_cross(d11, a, r1); kg_cross(d12, a, r2); kg_cross(d22, a, r22);
            int h1 = 3 + 2 * l, h2 = 4 + 2 * l;
            ws.M[h1 * KG_MS + h1] = 2 * Izz + mk * (kg_dot(d11, d11) + kg_dot(d12, d12)) + m.armature;
            float m12 = Izz + mk * kg_dot(d12, d22);
            ws.M[h1 * KG_MS + h2] = m12; ws.M[h2 * KG_MS + h1] = m12;
            ws.M[h2 * KG_MS + h2] = Izz + mk * kg_dot(d22, d22) + m.armature;
            float w1 = ws.qvel[h1], w2 = ws.qvel[h1] + ws.qvel[h2];
            float ac1[3], ac2[3];
            for (int c = 0; c < 3; c++) { ac1[c] = -w1 * w1 * r1[c]; ac2[c] = -w1 * w1 * r12[c] - w2 * w2 * r22[c]; }
            float lin1[3], lin2[3];   // m * d(com)/dq summed over the links a hinge moves
            for (int c = 0; c < 3; c++) { lin1[c] = mk * (d11[c] + d12[c]); lin2[c] = mk * d22[c]; }
            for (int i = 0; i < 3; i++) {
                float v1 = kg_dot(ws.sax[i], lin1), v2 = kg_dot(ws.sax[i], lin2);
                ws.M[i * KG_MS + h1] = v1; ws.M[h1 * KG_MS + i] = v1;
                ws.M[i * KG_MS + h2] = v2; ws.M[h2 * KG_MS + i] = v2;
                ws.sbias[l][i] = mk * (kg_dot(ws.sax[i], ac1) + kg_dot(ws.sax[i], ac2));
            }
            ws.bias[h1] = mk * (kg_dot(d11, ac1) + g * d11[2]) + mk * (kg_dot(d12, ac2) + g * d12[2]);
            ws.bias[h2] = mk * (kg_dot(d22, ac2) + g * d22[2]);
        } else if (l == 3) {
            const KgShape &sh = m.shape[ws.shape];
            const float mo = m.obj_mass;
            const float *R = ws.oR, *rl = sh.gpos;
            for (int i = 0; i < 3; i++) ws.M[(9 + i) * KG_MS + 9 + i] = mo + m.armature;
            // M_tr = -m R [r_l]x
            float rx[9] = {0, -rl[2], rl[1], rl[2], 0, -rl[0], -rl[1], rl[0], 0};
            float T[9];
            kg_mulmm(T, R, rx);
            for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) {
                float v = -mo * T[3 * j + i];
                ws.M[(9 + j) * KG_MS + 12 + i] = v; ws.M[(12 + i) * KG_MS + 9 + j] = v;
            }
            const float *Io = sh.Io, *Ic = sh.Ic;
            ws.M[12 * KG_MS + 12] = Io[0] + m.armature; ws.M[13 * KG_MS + 13] = Io[1] + m.armature; ws.M[14 * KG_MS + 14] = Io[2] + m.armature;
            ws.M[12 * KG_MS + 13] = ws.M[13 * KG_MS + 12] = Io[3];
            ws.M[12 * KG_MS + 14] = ws.M[14 * KG_MS + 12] = Io[4];
            ws.M[13 * KG_MS + 14] = ws.M[14 * KG_MS + 13] = Io[5];
            float w[3] = {ws.qvel[12], ws.qvel[13], ws.qvel[14]};
            float wr[3], wwr[3], Fl[3], Iw[3], Nc[3], rF[3], Fw[3];
            kg_cross(wr, w, rl); kg_cross(wwr, w, wr);
            float gz[3] = {R[6] * g, R[7] * g, R[8] * g};   // R^T (0,0,g)
            for (int c = 0; c < 3; c++) Fl[c] = mo * (wwr[c] + gz[c]);
            Iw[0] = Ic[0] * w[0] + Ic[3] * w[1] + Ic[4] * w[2];
            Iw[1] = Ic[3] * w[0] + Ic[1] * w[1] + Ic[5] * w[2];
            Iw[2] = Ic[4] * w[0] + Ic[5] * w[1] + Ic[2] * w[2];
            kg_cross(Nc, w, Iw); kg_cross(rF, rl, Fl);
            kg_mulmv(Fw, R, Fl);
            for (int c = 0; c < 3; c++) { ws.bias[9 + c] = Fw[c]; ws.bias[12 + c] = Nc[c] + rF[c]; }
        } else if (l == 4) {
            for (int i = 0; i < 3; i++) ws.M[i * KG_MS + i] = m.m_hand_total + m.armature;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l < KG_NV) {
            float b = ws.bias[l], act = 0.f, v = ws.qvel[l];
            if (l < 3) {
                b = ws.sbias[0][l] + ws.sbias[1][l] + ws.sbias[2][l] + m.m_hand_total * m.gravity * ws.sax[l][2];
                ws.bias[l] = b;
                float c = fminf(fmaxf(ws.ctrl[2 * l], -m.ctrl_slide), m.ctrl_slide);
                act = m.act_kv_slide * (c - v) + m.act_gear_slide * ws.ctrl[2 * l + 1];
            } else if (l < 9 && ((l - 3) & 1) == 0) {
                float c = fminf(fmaxf(ws.ctrl[6 + (l - 3) / 2], -m.ctrl_finger), m.ctrl_finger);
                act = m.act_kv_finger * (c - v);
            }
            ws.fs[l] = -m.damping * v - b + act;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ dense 15x15 Cholesky
// x <- (A)^-1 x for the SPD Newton Hessian A (rows of stride KG_MS in shared memory, written by the lanes that own
// them); on return A holds the lower factor.  GPU: every lane keeps its row in registers and the right-looking
// factorisation, the forward and the backward substitution exchange pivots / columns with warp shuffles
// (about 400 warp instructions, no serial lane-0 section).  The host emulation runs the same arithmetic in the same order.
KG_DEV void kg_chol_factor_solve(float *A, float *x, int kg_lane) {
#if defined(__CUDA_ARCH__)
    const unsigned FULL = 0xffffffffu;
    const int l = kg_lane, r = l < KG_NV ? l : KG_NV - 1;
    float a[16], invd[KG_NV];
    // lanes 0..14 hold the rows of A, lane 15 holds the right-hand side as a 16th row: the right-looking updates then
    // perform the forward substitution L y = x on it for free (y_j appears in lane 15 when column j is scaled)
    kg_ld16(l == KG_NV ? x : A + r * KG_MS, a);
    float xi = 0.f;
#pragma unroll
    for (int j = 0; j < KG_NV; j++) {
        float djj = __shfl_sync(FULL, a[j], j);
#ifndef KG_CHOL_EXACT
        float inv = rsqrtf(fmaxf(djj, 1e-20f)), d = djj * inv;   // the factor only shapes the Newton direction; 2 ulp is plenty
#else
        float d = sqrtf(fmaxf(djj, 1e-20f)), inv = 1.0f / d;
#endif
        invd[j] = inv;
        float lj = (l == j) ? d : a[j] * inv;      // L[l][j] for l >= j (rows above j hold don't-care values)
        a[j] = lj;
        float yj = __shfl_sync(FULL, lj, KG_NV);
        if (l == j) xi = yj;
#pragma unroll
        for (int k = j + 1; k < KG_NV; k++) {
            float lkj = __shfl_sync(FULL, lj, k);
            a[k] = fmaf(-lj, lkj, a[k]);
        }
    }
    // the backward pass needs column l of L: transpose through shared memory
    __syncwarp();
    if (l < KG_NV) {
        float4 *row = reinterpret_cast<float4 *>(A + l * KG_MS);
        row[0] = make_float4(a[0], a[1], a[2], a[3]); row[1] = make_float4(a[4], a[5], a[6], a[7]);
        row[2] = make_float4(a[8], a[9], a[10], a[11]); row[3] = make_float4(a[12], a[13], a[14], 0.f);
    }
    __syncwarp();
    float c[KG_NV];
#pragma unroll
    for (int k = 0; k < KG_NV; k++) c[k] = A[k * KG_MS + r];
#pragma unroll
    for (int k = KG_NV - 1; k >= 0; k--) {
        float t = __shfl_sync(FULL, xi, k) * invd[k];
        xi = (l == k) ? t : (l < k ? fmaf(-c[k], t, xi) : xi);
    }
    if (l < KG_NV) x[l] = xi;
    __syncwarp();
#else
    (void)kg_lane;
    float invd[KG_NV];
    for (int j = 0; j < KG_NV; j++) {
        float d = sqrtf(fmaxf(A[j * KG_MS + j], 1e-20f)), inv = 1.0f / d;
        invd[j] = inv;
        A[j * KG_MS + j] = d;
        for (int l = j + 1; l < KG_NV; l++) A[l * KG_MS + j] *= inv;
        for (int l = j + 1; l < KG_NV; l++) for (int k = j + 1; k <= l; k++) A[l * KG_MS + k] -= A[l * KG_MS + j] * A[k * KG_MS + j];
    }
    for (int k = 0; k < KG_NV; k++) {
        float t = x[k] * invd[k];
        x[k] = t;
        for (int l = k + 1; l < KG_NV; l++) x[l] -= A[l * KG_MS + k] * t;
    }
    for (int k = KG_NV - 1; k >= 0; k--) {
        float t = x[k] * invd[k];
        x[k] = t;
        for (int l = 0; l < k; l++) x[l] -= A[k * KG_MS + l] * t;
    }
#endif
}

// ------------------------------------------------------------------------------------------------ P5 hull support (hill climbing)
// Support point of geom g's convex hull, inflated by `margin`, along the unit world direction dir, expressed relative
// to `origin`.  The geom frame comes from the workspace (shared memory), the vertices from the read-only hull table;
// `cur` is the warm-start vertex (kept per pair across calls and substeps).
template <class WS> KG_DEV void kg_support(const KgModel &m, const WS &ws, int g, int mesh, int &cur, const float *dir, const float *origin, float margin, float *out) {
    const KgMesh &mi = m.mesh[mesh];
    const float *vert = m.hull_vert + 4 * (size_t)mi.vert0;
    const unsigned short *adj8 = m.adj8;
    const float *nbr8 = m.nbr8;
    const float *R = ws.gmat[g];
    float ld[3];
    kg_multmv(ld, R, dir);
    int c = cur;
    auto v = KG_LDG4(vert + 4 * c);
    float best = v.x * ld[0] + v.y * ld[1] + v.z * ld[2];
    float bx = v.x, by = v.y, bz = v.z;
    unsigned info = KG_F2U(v.w);
    for (int guard = 0; guard < 1024; guard++) {
        // one hill-climbing move: the ids and the vertex records of a chunk's 8 neighbours are addressed from the chunk index
        // alone, and the first chunk of a vertex from the vertex index alone -- so the evaluation of the warm-start vertex
        // (its record and its neighbours load together) and every move cost ONE memory latency whatever the degree;
        // ties keep the first neighbour in list order
        const int ovf0 = (int)(info & 0xffffffu), novf = (int)(info >> 24);
        int nxt = c;
        for (int k = 0; k <= novf; k++) {
            const size_t ci = k == 0 ? (size_t)(mi.vert0 + c) : (size_t)(ovf0 + k - 1);
            auto q = KG_LDGU4(adj8 + 8 * ci);
            const unsigned id[8] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16, q.z & 0xffffu, q.z >> 16, q.w & 0xffffu, q.w >> 16};
            decltype(v) w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = KG_LDG4(nbr8 + 4 * (8 * ci + i));
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float d = w[i].x * ld[0] + w[i].y * ld[1] + w[i].z * ld[2];
                if (d > best) { best = d; nxt = (int)id[i]; bx = w[i].x; by = w[i].y; bz = w[i].z; info = KG_F2U(w[i].w); }
            }
        }
        if (nxt == c) break;
        c = nxt;
        KG_COUNT(21, 1);
    }
    KG_COUNT(20, 1);
    cur = c;
    float lv[3] = {bx, by, bz}, wv[3];
    kg_mulmv(wv, R, lv);
    for (int k = 0; k < 3; k++) out[k] = wv[k] + (ws.gpos[g][k] - origin[k]) + margin * dir[k];
}

// Narrowphase team size: a full MPR is run by a TEAM of lanes that share the two support scans of an iteration and execute
// the (scalar) portal algebra redundantly.  KG_MPR_TEAMS = 2 or 4: that many teams (16 / 8 lanes) per warp, each with its own
// task, advancing ONE MPR iteration per trip of the kernel's loop with a warp-wide rendezvous after it -- teams in the same
// state of the algorithm then issue the algebra, most of an iteration's instructions, once for all of them; the scans take
// more trips.  (Teams that merely share a warp without the rendezvous drift apart and issue everything separately: measured,
// no gain.)
#ifndef KG_MPR_TEAMS
#define KG_MPR_TEAMS 1
#endif
#if defined(__CUDA_ARCH__)
#define KG_MPR_W (32 / KG_MPR_TEAMS)
#define KG_MPR_MASK(lane) (KG_MPR_TEAMS == 1 ? 0xffffffffu : (((1u << (KG_MPR_W & 31)) - 1u) << ((lane) & ~(KG_MPR_W - 1))))
#else
#define KG_MPR_W 32
#define KG_MPR_MASK(lane) 0xffffffffu
#endif
// Warp-cooperative support of geom g's hull along the (warp-uniform) direction dir, inflated by `margin`, relative to
// `origin`.  The 32 lanes scan the whole vertex table with independent 128-bit loads (no hill-climbing chain).  The
// tables are padded to a multiple of 32 vertices with copies of vertex 0, so the scan is branch-free; each lane tracks
// only (best dot, index), the arg-max is reduced with redux.sync on an order-preserving integer key and, among equal
// maxima, the lowest vertex index wins (a pad entry can therefore never be returned); the winning vertex is then
// fetched once.  Every lane returns the same result.  The host emulation scans the vertices in index order (same tie rule).
KG_DEV const float *kg_hull_table(const KgModel &m, const KgMprWS &ws, int mesh) {
    // object -> this warp's shared-memory copy, fingers -> the CTA's shared-memory copy, palm -> global memory
#ifdef KG_MPR_L1
    (void)ws;
    return m.hull_pad + 4 * (size_t)m.hpad0[mesh];     // every table through the read-only L1 path (KG_MPR_L1)
#else
    return mesh >= 3 ? ws.objhull : (mesh == 0 ? m.hull_pad : ws.fhull + (mesh == 2 ? 4 * m.hnpad[1] : 0));
#endif
}
KG_DEV void kg_support_coop(const KgModel &m, const KgMprWS &ws, int slot, int mesh, const float *dir, const float *origin, float margin,
                            float *out, int *vidx, int kg_lane) {
    const float *vert = kg_hull_table(m, ws, mesh);
    const auto r0 = KG_LD4S(ws.gr[slot]), r1 = KG_LD4S(ws.gr[slot] + 4), r2 = KG_LD4S(ws.gr[slot] + 8), gc = KG_LD4S(ws.gp[slot]);
    const float R[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z}, gpos[3] = {gc.x, gc.y, gc.z};
    float ld[3];
    kg_multmv(ld, R, dir);
    float best = -FLT_MAX;
    int bi = 0;
#if defined(__CUDA_ARCH__) && defined(KG_MPR_L1)
    // scan only the vertices that can be a support vertex for a direction of ld's octant (kg_model.h: exact candidate
    // lists, 25-40 % of a finger or palm hull); an entry carries its vertex index in w
    const int oct = (ld[0] < 0.f ? 1 : 0) | (ld[1] < 0.f ? 2 : 0) | (ld[2] < 0.f ? 4 : 0);
    const int npad = m.octn[mesh][oct];
    const float4 *v4 = reinterpret_cast<const float4 *>(m.hull_oct) + m.oct0[mesh][oct];
    bi = 0x7fffffff;
#pragma unroll 4
    for (int i = kg_lane & (KG_MPR_W - 1); i < npad; i += KG_MPR_W) {
        float4 v = __ldg(v4 + i);
        float d = v.x * ld[0] + v.y * ld[1] + v.z * ld[2];
        if (d > best) { best = d; bi = __float_as_int(v.w); }
    }
    unsigned u = __float_as_uint(best);
    unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    const unsigned team = KG_MPR_MASK(kg_lane);
    unsigned mx = __reduce_max_sync(team, key);
    bi = __reduce_min_sync(team, key == mx ? bi : 0x7fffffff);
    float4 w = __ldg(reinterpret_cast<const float4 *>(vert) + bi);
    float lv[3] = {w.x, w.y, w.z};
    (void)vert;
#elif defined(__CUDA_ARCH__)
    const int npad = m.hnpad[mesh];
    const float4 *v4 = reinterpret_cast<const float4 *>(vert);
#pragma unroll 4
    for (int i = kg_lane; i < npad; i += 32) {
#ifdef KG_MPR_L1
        float4 v = __ldg(v4 + i);
#else
        float4 v = v4[i];
#endif
        float d = v.x * ld[0] + v.y * ld[1] + v.z * ld[2];
        if (d > best) { best = d; bi = i; }
    }
    unsigned u = __float_as_uint(best);
    unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    unsigned mx = __reduce_max_sync(0xffffffffu, key);
    bi = __reduce_min_sync(0xffffffffu, key == mx ? bi : 0x7fffffff);
#ifdef KG_MPR_L1
    float4 w = __ldg(v4 + bi);
#else
    float4 w = v4[bi];
#endif
    float lv[3] = {w.x, w.y, w.z};
#else
    (void)kg_lane;
    const int nv = m.mesh[mesh].nvert;
    for (int i = 0; i < nv; i++) {
        const float *v = vert + 4 * i;
        float d = v[0] * ld[0] + v[1] * ld[1] + v[2] * ld[2];
        if (d > best) { best = d; bi = i; }
    }
    float lv[3] = {vert[4 * bi], vert[4 * bi + 1], vert[4 * bi + 2]};
#endif
    *vidx = bi;
    float wv[3];
    kg_mulmv(wv, R, lv);
    for (int k = 0; k < 3; k++) out[k] = wv[k] + (gpos[k] - origin[k]) + margin * dir[k];
}

// ------------------------------------------------------------------------------------------------ P5 MPR (libccd's algorithm, fp32)
struct KgSupp { float v[3], v1[3]; };   // Minkowski point and the witness on hull 1 (witness 2 = v1 - v)
// libccd's ccdIsZero is an ABSOLUTE test against DBL_EPSILON; the reference engine runs it in double.  Keeping the
// same absolute threshold in fp32 (2.2e-16 is representable) reproduces its branch decisions, whereas FLT_EPSILON
// would classify every small-but-regular quantity (|v0 x v1|^2, the Gram determinant of a refined portal) as zero.
#define KG_CCD_EPS 2.220446e-16f
KG_DEV bool kg_iszero(float x) { return fabsf(x) < KG_CCD_EPS; }
KG_DEV bool kg_feq(float a, float b) {
    float ab = fabsf(a - b);
    if (ab < FLT_EPSILON) return true;
    float fa = fabsf(a), fb = fabsf(b);
    return ab < FLT_EPSILON * (fb > fa ? fb : fa);
}
KG_DEV void kg_portal_dir(const KgSupp *p, float *dir) {
    float a[3], b[3];
    kg_sub(a, p[2].v, p[1].v); kg_sub(b, p[3].v, p[1].v);
    kg_cross(dir, a, b);
    kg_normalize_dir(dir);
}
KG_DEV bool kg_reach_tol(const KgSupp *p, const KgSupp &v4, const float *dir) {
    float dv4 = kg_dot(v4.v, dir);
    float d1 = dv4 - kg_dot(p[1].v, dir), d2 = dv4 - kg_dot(p[2].v, dir), d3 = dv4 - kg_dot(p[3].v, dir);
    float mn = fminf(d1, fminf(d2, d3));
    return kg_feq(mn, KG_MPR_TOL) || mn < KG_MPR_TOL;
}
KG_DEV void kg_expand_portal(KgSupp *p, const KgSupp &v4) {
    float v4v0[3];
    kg_cross(v4v0, v4.v, p[0].v);
    if (kg_dot(p[1].v, v4v0) > 0) {
        if (kg_dot(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4;
    } else {
        if (kg_dot(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4;
    }
}
KG_DEV float kg_seg_dist2(const float *a, const float *b, float *w) {   // closest point to the origin on [a,b]
    float d[3];
    kg_sub(d, b, a);
    float t = -kg_dot(a, d) / kg_dot(d, d);
    if (t < 0 || kg_iszero(t)) kg_copy(w, a);
    else if (t > 1 || kg_feq(t, 1)) kg_copy(w, b);
    else kg_addscl(w, a, d, t);
    return kg_dot(w, w);
}
KG_DEV float kg_tri_dist2(const float *x0, const float *Bv, const float *C, float *w) {   // origin vs triangle
    float d1[3], d2[3];
    kg_sub(d1, Bv, x0); kg_sub(d2, C, x0);
    float v = kg_dot(d1, d1), ww = kg_dot(d2, d2), p = kg_dot(x0, d1), q = kg_dot(x0, d2), r = kg_dot(d1, d2);
    float div = ww * v - r * r, s, t = 0;
    if (kg_iszero(div)) s = -1; else { s = (q * r - ww * p) / div; t = (-s * r - q) / ww; }
    if ((kg_iszero(s) || s > 0) && (kg_feq(s, 1) || s < 1) && (kg_iszero(t) || t > 0) && (kg_feq(t, 1) || t < 1) && (kg_feq(t + s, 1) || t + s < 1)) {
        kg_addscl(w, x0, d1, s); kg_addscl(w, w, d2, t);
        return kg_dot(w, w);
    }
    float w2[3], dist = kg_seg_dist2(x0, Bv, w), dv;
    dv = kg_seg_dist2(x0, C, w2); if (dv < dist) { dist = dv; kg_copy(w, w2); }
    dv = kg_seg_dist2(Bv, C, w2); if (dv < dist) { dist = dv; kg_copy(w, w2); }
    return dist;
}
#if !defined(__CUDA_ARCH__) && defined(KG_EMU_HIST)
static long kg_emu_mpr_iters = 0;
#endif
// ccdMPRPenetration on the margin-inflated hulls of geoms g1, g2; coordinates are relative to geom g1's centre.
// WARP-UNIFORM: all 32 lanes call it with the same arguments, run the same scalar portal algebra and share the two
// cooperative support scans of every iteration; all lanes return the same result.
// Written as ONE loop around ONE support evaluation (state machine over libccd's discoverPortal / refinePortal /
// findPenetr) so that the hill-climbing support code exists once in the instruction stream.
// On a "no overlap" exit that PROVES separation (support of A-B along dir is negative) *sep_found is set and dir is
// returned in nrm: the caller re-tests that axis first in the next substep (two supports instead of a full MPR).
struct KgMprIt { KgSupp p[4]; float dir[3]; int state, it, guard; };
KG_DEV void kg_mpr_init(const KgMprWS &ws, KgMprIt &q) {
    q.p[0].v1[0] = q.p[0].v1[1] = q.p[0].v1[2] = 0.f;
    kg_sub(q.p[0].v, ws.gp[0], ws.gp[1]);
    if (kg_iszero(q.p[0].v[0]) && kg_iszero(q.p[0].v[1]) && kg_iszero(q.p[0].v[2])) q.p[0].v[0] += KG_CCD_EPS * 10;
    for (int c = 0; c < 3; c++) q.dir[c] = -q.p[0].v[c];
    kg_normalize_dir(q.dir);
    q.state = 0; q.it = 0; q.guard = 0;
}
// ONE iteration of the state machine (one support evaluation of the Minkowski difference): 0 = go on, 1 = penetration found
// (depth, nrm, pos set), 2 = no contact (*sep_found / *gap / nrm set when separation along nrm was proven).
KG_DEV int kg_mpr_iter(const KgModel &m, const KgMprWS &ws, KgMprIt &q, int mesh1, int mesh2, float hmargin, int &cur1, int &cur2,
                       float *depth, float *nrm, float *pos, int *sep_found, float *gap, int kg_lane) {
    KgSupp *p = q.p;
    float *dir = q.dir;
    float va[3], vb[3], dot;
    const float *org = ws.gp[0];
    if (++q.guard > 400) return 2;
    KgSupp s;
    if (kg_lane == 0) KG_COUNT(23, 1);
#if !defined(__CUDA_ARCH__) && defined(KG_EMU_HIST)
    kg_emu_mpr_iters++;
#endif
    {
        float nd[3] = {-dir[0], -dir[1], -dir[2]}, v2[3];
        kg_support_coop(m, ws, 0, mesh1, dir, org, hmargin, s.v1, &cur1, kg_lane);
        kg_support_coop(m, ws, 1, mesh2, nd, org, hmargin, v2, &cur2, kg_lane);
        kg_sub(s.v, s.v1, v2);
    }
    bool check_portal = false;
    if (q.state == 0) {
        p[1] = s;
        dot = kg_dot(s.v, dir);
        if (kg_iszero(dot) || dot < 0) { if (dot < 0) { *sep_found = 1; *gap = -dot; kg_copy(nrm, dir); } return 2; }
        kg_cross(dir, p[0].v, p[1].v);
        if (kg_iszero(kg_dot(dir, dir))) {
            if (kg_iszero(p[1].v[0]) && kg_iszero(p[1].v[1]) && kg_iszero(p[1].v[2])) return 2;   // touching, no normal
            *depth = sqrtf(kg_dot(p[1].v, p[1].v));
            kg_copy(nrm, p[1].v); kg_normalize(nrm);
            for (int c = 0; c < 3; c++) pos[c] = p[1].v1[c] - 0.5f * p[1].v[c];
            return 1;
        }
        kg_normalize_dir(dir);
        q.state = 1;
    } else if (q.state == 1) {
        p[2] = s;
        dot = kg_dot(s.v, dir);
        if (kg_iszero(dot) || dot < 0) { if (dot < 0) { *sep_found = 1; *gap = -dot; kg_copy(nrm, dir); } return 2; }
        kg_sub(va, p[1].v, p[0].v); kg_sub(vb, p[2].v, p[0].v);
        kg_cross(dir, va, vb); kg_normalize_dir(dir);
        if (kg_dot(dir, p[0].v) > 0) { KgSupp t = p[1]; p[1] = p[2]; p[2] = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
        q.state = 2;
    } else if (q.state == 2) {
        p[3] = s;
        dot = kg_dot(s.v, dir);
        if (kg_iszero(dot) || dot < 0) { if (dot < 0) { *sep_found = 1; *gap = -dot; kg_copy(nrm, dir); } return 2; }
        bool cont = false;
        kg_cross(va, p[1].v, p[3].v); dot = kg_dot(va, p[0].v);
        if (dot < 0 && !kg_iszero(dot)) { p[2] = p[3]; cont = true; }
        if (!cont) {
            kg_cross(va, p[3].v, p[2].v); dot = kg_dot(va, p[0].v);
            if (dot < 0 && !kg_iszero(dot)) { p[1] = p[3]; cont = true; }
        }
        if (cont) {
            kg_sub(va, p[1].v, p[0].v); kg_sub(vb, p[2].v, p[0].v);
            kg_cross(dir, va, vb); kg_normalize_dir(dir);
        } else { q.state = 3; check_portal = true; }
    } else if (q.state == 3) {
        dot = kg_dot(s.v, dir);
        if (!(kg_iszero(dot) || dot > 0)) { *sep_found = 1; *gap = -dot; kg_copy(nrm, dir); return 2; }
        if (kg_reach_tol(p, s, dir)) return 2;
        kg_expand_portal(p, s);
        check_portal = true;
    } else {
        if (kg_reach_tol(p, s, dir) || q.it > KG_MPR_ITER) {
            float w[3];
            *depth = sqrtf(kg_tri_dist2(p[1].v, p[2].v, p[3].v, w));
            if (kg_iszero(w[0]) && kg_iszero(w[1]) && kg_iszero(w[2])) kg_copy(w, dir);
            kg_copy(nrm, w); kg_normalize(nrm);
            // findPos: barycentric coordinates of the origin in the portal tetrahedron
            float b[4], t[3], sum;
            kg_cross(t, p[1].v, p[2].v); b[0] = kg_dot(t, p[3].v);
            kg_cross(t, p[3].v, p[2].v); b[1] = kg_dot(t, p[0].v);
            kg_cross(t, p[0].v, p[1].v); b[2] = kg_dot(t, p[3].v);
            kg_cross(t, p[2].v, p[1].v); b[3] = kg_dot(t, p[0].v);
            sum = b[0] + b[1] + b[2] + b[3];
            if (kg_iszero(sum) || sum < 0) {
                kg_portal_dir(p, dir);
                b[0] = 0;
                kg_cross(t, p[2].v, p[3].v); b[1] = kg_dot(t, dir);
                kg_cross(t, p[3].v, p[1].v); b[2] = kg_dot(t, dir);
                kg_cross(t, p[1].v, p[2].v); b[3] = kg_dot(t, dir);
                sum = b[1] + b[2] + b[3];
            }
            float inv = 1.0f / sum;
            for (int c = 0; c < 3; c++) {
                float s1 = 0, sv = 0;
                for (int i = 0; i < 4; i++) { s1 += b[i] * p[i].v1[c]; sv += b[i] * p[i].v[c]; }
                pos[c] = (s1 - 0.5f * sv) * inv;   // (p1 + p2)/2 with p2 = p1 - v
            }
            return 1;
        }
        kg_expand_portal(p, s);
        q.it++;
        kg_portal_dir(p, dir);
    }
    if (check_portal) {
        // refinePortal's loop head: does the portal already enclose the origin ray?
        kg_portal_dir(p, dir);
        dot = kg_dot(dir, p[1].v);
        if (kg_iszero(dot) || dot > 0) q.state = 4;   // findPenetr starts with this same portal direction
    }
    return 0;
}
KG_DEV bool kg_mpr(const KgModel &m, const KgMprWS &ws, int g1, int mesh1, int g2, int mesh2, float hmargin, int &cur1, int &cur2,
                   float *depth, float *nrm, float *pos, int *sep_found, float *gap, int kg_lane) {
    (void)g1; (void)g2;
    *sep_found = 0; *gap = 0.f;
    KgMprIt q;
    kg_mpr_init(ws, q);
    if (kg_lane == 0) KG_COUNT(22, 1);
    for (;;) {
        const int r = kg_mpr_iter(m, ws, q, mesh1, mesh2, hmargin, cur1, cur2, depth, nrm, pos, sep_found, gap, kg_lane);
        if (r) return r == 1;
    }
}

// mju_makeFrame on frame[0:3] = normal
KG_DEV void kg_make_frame(const float *n, float *t1, float *t2) {
    t1[0] = 0; t1[1] = 0; t1[2] = 0;
    if (n[1] < 0.5f && n[1] > -0.5f) t1[1] = 1; else t1[2] = 1;
    float d = kg_dot(n, t1);
    kg_addscl(t1, t1, n, -d);
    kg_normalize(t1);
    kg_cross(t2, n, t1);
}

// ------------------------------------------------------------------------------------------------ P4 + P5 collision
// pair table: lanes 0..7 plane vs geoms 1..8; lanes 8..13 object vs geoms 2..7; lane 14 object vs palm;
// lanes 15..29 the dynamic pairs (palm-distal x3, finger-finger x12).  geom1 of the explicit pairs is
// the object (XML:159-166), so their normals point from the object to the hand geom.
KG_DEV void kg_pair_of_lane(int l, int &g1, int &g2) {
    if (l < 8) { g1 = 0; g2 = l + 1; }
    else if (l < 14) { g1 = 8; g2 = l - 6; }
    else if (l == 14) { g1 = 8; g2 = 1; }
    else {
        // nibble k of T1 / T2 = geom1 / geom2 of dynamic pair k: (1,3)(1,5)(1,7)(2,4)(2,5)(3,4)(3,5)(2,6)(2,7)(3,6)(3,7)(4,6)(4,7)(5,6)(5,7)
        const unsigned long long T1 = 0x554433223322111ull, T2 = 0x767676765454753ull;
        const int sh = 4 * (l - 15);
        g1 = (int)((T1 >> sh) & 15); g2 = (int)((T2 >> sh) & 15);
    }
}
KG_DEV int kg_mesh_of_geom_s(int shape, int g) { return g == 8 ? 3 + shape : (g == 1 ? 0 : ((g & 1) ? 2 : 1)); }
template <class WS> KG_DEV int kg_mesh_of_geom(const WS &ws, int g) { return kg_mesh_of_geom_s(ws.shape, g); }
template <class WS> KG_DEV float kg_rbound(const KgModel &m, const WS &ws, int g) { return g == 8 ? m.shape[ws.shape].rbound : (g == 1 ? m.hand_rbound[0] : ((g & 1) ? m.hand_rbound[2] : m.hand_rbound[1])); }

// One narrowphase task: full MPR of pair lane lp of env `env` (warp-uniform: all 32 lanes run the portal algebra and
// share the cooperative support scans).  The geom frames come from the env's carry record, the result is filed there
// by lane 0: separating axis / warm-start vertices, and on a hit the contact in staging slot 32 + (lp - 8).
// the frames of a task's two geoms (2 x (3 + 9) words) from the env's carry record into the team's workspace
KG_DEV void kg_mpr_stage(KgMprWS &mw, const unsigned char *carry, int g1, int g2, int kg_lane) {
#if defined(__CUDA_ARCH__)
    const unsigned team = KG_MPR_MASK(kg_lane);
    __syncwarp(team);
    for (int l = kg_lane & (KG_MPR_W - 1); l < 25; l += KG_MPR_W) {
#else
    KG_PHASE(l) {
#endif
        if (l < 24) {
            const int which = l >= 12, j = l - 12 * which, g = which ? g2 : g1;
            if (j < 3) mw.gp[which][j] = reinterpret_cast<const float *>(carry + KG_COFF(gpos))[3 * g + j];
            else mw.gr[which][4 * ((j - 3) / 3) + (j - 3) % 3] = reinterpret_cast<const float *>(carry + KG_COFF(gmat))[9 * g + j - 3];
        }
        if (l == 24) mw.shape = *reinterpret_cast<const int *>(carry + KG_COFF(shape));
    }
#if defined(__CUDA_ARCH__)
    __syncwarp(team);
#else
    KG_SYNC();
#endif
}
// file the result of pair lane lp in the env's carry record (the team's first lane): separating axis / budget / warm-start
// vertices, and on a hit the contact in staging slot 32 + (lp - 8)
KG_DEV void kg_mpr_file(const KgModel &m, const KgMprWS &mw, unsigned char *carry, int lp, int g1, int g2, bool hit, float depth, const float *nrm, const float *pos,
                        int sepf, float gap, int c1, int c2, int kg_lane) {
#if defined(__CUDA_ARCH__)
    if ((kg_lane & (KG_MPR_W - 1)) == 0) {
#else
    (void)kg_lane;
    {
#endif
        const float margin = m.margin;
        reinterpret_cast<int *>(carry + KG_COFF_C(sepok))[lp] = sepf;
        if (sepf) {
            float *sp = reinterpret_cast<float *>(carry + KG_COFF_C(sep)) + 3 * lp; sp[0] = nrm[0]; sp[1] = nrm[1]; sp[2] = nrm[2];
            reinterpret_cast<float *>(carry + KG_COFF_C(slack))[lp] = gap;      // the collision stage of this very substep has already charged the motion up to here
        }
        reinterpret_cast<unsigned short *>(carry + KG_COFF_C(cv1))[lp] = (unsigned short)c1;
        reinterpret_cast<unsigned short *>(carry + KG_COFF_C(cv2))[lp] = (unsigned short)c2;
        if (hit && depth > 0.f) {
            const int sl = 32 + (lp - 8);
            reinterpret_cast<float *>(carry + KG_COFF_B(st_dist))[sl] = margin - depth;
            float *sp = reinterpret_cast<float *>(carry + KG_COFF_B(st_pos)) + 3 * sl, *sn = reinterpret_cast<float *>(carry + KG_COFF_B(st_n)) + 3 * sl;
            for (int c = 0; c < 3; c++) { sp[c] = pos[c] + mw.gp[0][c]; sn[c] = nrm[c]; }
            (carry + KG_COFF_B(st_g1))[sl] = (unsigned char)g1; (carry + KG_COFF_B(st_g2))[sl] = (unsigned char)g2;
            reinterpret_cast<int *>(carry + KG_COFF(cnt))[lp] = 1;
        }
    }
}
// One narrowphase task from start to end: full MPR of pair lane lp of an env, by one team.
KG_DEV void kg_mpr_task(const KgModel &m, KgMprWS &mw, unsigned char *carry, int lp, int kg_lane) {
    int g1, g2;
    kg_pair_of_lane(lp, g1, g2);
    kg_mpr_stage(mw, carry, g1, g2, kg_lane);
#ifndef KG_MPR_L1
    if (mw.shape != mw.hull_shape) {      // this warp's object-hull copy holds another shape
        KG_SYNC();
        KG_PHASE(l) {
            const float *src = m.hull_pad + 4 * (size_t)m.hpad0[3 + mw.shape];
            const int npad = m.hnpad[3 + mw.shape];
            for (int i = l; i < npad; i += 32) {
                auto v = KG_LDG4(src + 4 * i);
                mw.objhull[4 * i] = v.x; mw.objhull[4 * i + 1] = v.y; mw.objhull[4 * i + 2] = v.z; mw.objhull[4 * i + 3] = v.w;
            }
        }
        KG_SYNC();
        KG_PHASE(l) { if (l == 0) mw.hull_shape = mw.shape; }
        KG_SYNC();
    }
#endif
    const int mesh1 = kg_mesh_of_geom_s(mw.shape, g1), mesh2 = kg_mesh_of_geom_s(mw.shape, g2);
    int c1 = 0, c2 = 0, sepf = 0;
    float depth = 0.f, nrm[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, gap = 0.f;
#if !defined(__CUDA_ARCH__) && defined(KG_EMU_HIST)
    long it0_ = kg_emu_mpr_iters;
#endif
    bool hit = kg_mpr(m, mw, g1, mesh1, g2, mesh2, 0.5f * m.margin, c1, c2, &depth, nrm, pos, &sepf, &gap, kg_lane);
#if !defined(__CUDA_ARCH__) && defined(KG_EMU_HIST)
    fprintf(stderr, "MPR %d %d %ld %d %d\n", g1, g2, kg_emu_mpr_iters - it0_, (int)hit, sepf);
#endif
    kg_mpr_file(m, mw, carry, lp, g1, g2, hit, depth, nrm, pos, sepf, gap, c1, c2, kg_lane);
#if defined(__CUDA_ARCH__)
    __syncwarp(KG_MPR_MASK(kg_lane));
#else
    KG_SYNC();
#endif
}

// Lane-parallel part of the collision stage: broadphase, plane contacts, re-test of the cached separating axes.
// Returns (every lane) the mask of the pair lanes 8..29 that need a full MPR; those become narrowphase tasks.
template <class WS> KG_DEV unsigned kg_collision_lane(const KgModel &m, WS &ws, int kg_lane) {
#ifndef KG_NO_SLACK
    // how far any surface point of geom g can have moved since the previous collision stage: |centre displacement| +
    // bounding radius x rotation angle of the last integration (hinges: h |qvel|, free body: h |omega|; the palm only
    // translates; the ground does not move) + rounding allowance
    KG_PHASE(l) {
        if (l < 9) {
            float mv = 0.f;
            if (l >= 1) {
                float d[3]; kg_sub(d, ws.gpos[l], ws.cpos[l]);
                const float h = m.timestep;
                float th = 0.f;
                if (l == 8) th = h * sqrtf(ws.qvel[12] * ws.qvel[12] + ws.qvel[13] * ws.qvel[13] + ws.qvel[14] * ws.qvel[14]);
                else if (l >= 2) { const int k = (l - 2) >> 1; th = h * (fabsf(ws.qvel[3 + 2 * k]) + ((l & 1) ? fabsf(ws.qvel[4 + 2 * k]) : 0.f)); }
                mv = 1.0001f * (sqrtf(kg_dot(d, d)) + kg_rbound(m, ws, l) * th) + 2e-6f;
                kg_copy(ws.cpos[l], ws.gpos[l]);
            }
            ws.red[l] = mv;
        }
    }
    KG_SYNC();
#endif
    KG_PHASE(l) {
        int n = 0;
        float tried = 0.f;
        if (l < 30) {
            int g1, g2;
            kg_pair_of_lane(l, g1, g2);
            const float margin = m.margin;
            const float rb2 = kg_rbound(m, ws, g2);
            const int mesh2 = kg_mesh_of_geom(ws, g2);
#ifndef KG_NO_SLACK
            const float budget = ws.sepok[l] ? ws.slack[l] - ws.red[g1] - ws.red[g2] : -1.f;
#else
            const float budget = -1.f;
#endif
            // Stage 0: what this lane has to look up.  job 1 = plane lane: lowest hull vertex of geom g2; job 2 = hull pair
            // whose cached separating axis gets re-tested: first the support of g1 along it.  Both kinds of lanes then run
            // the SAME hill-climbing call side by side (one call site, no serialisation between plane and pair lanes).
            int job = 0, gA = g2, meshA = mesh2, curA = 0, mesh1 = 0, c2 = 0;
            float dirA[3] = {0.f, 0.f, -1.f}, orgA[3] = {0.f, 0.f, 0.f}, marginA = 0.f;
            if (l < 8) {
                if (ws.gpos[g2][2] <= rb2 + margin && !(budget > 0.f)) {
                    job = 1;
                    curA = ws.cv2[l];
                    if (curA >= m.mesh[mesh2].nvert) curA = 0;
                } else if (budget > 0.f) ws.slack[l] = budget;
                else ws.sepok[l] = 0;                                   // beyond the bounding sphere: nothing cached
            } else {
                float rb1 = kg_rbound(m, ws, g1), df[3];
                kg_sub(df, ws.gpos[g2], ws.gpos[g1]);
                float lim = rb1 + rb2 + margin;
                if (kg_dot(df, df) <= lim * lim) {
                    if (budget > 0.f) ws.slack[l] = budget;             // still provably separated along the cached axis
                    else {
                        mesh1 = kg_mesh_of_geom(ws, g1);
                        int c1 = ws.cv1[l];
                        c2 = ws.cv2[l];
                        if (c1 >= m.mesh[mesh1].nvert) c1 = 0;
                        if (c2 >= m.mesh[mesh2].nvert) c2 = 0;
                        if (ws.sepok[l]) {
                            // the axis that separated the inflated hulls in the previous substep: still separating?
                            job = 2; gA = g1; meshA = mesh1; curA = c1; marginA = 0.5f * margin;
                            for (int k = 0; k < 3; k++) { dirA[k] = ws.sep[l][k]; orgA[k] = ws.gpos[g1][k]; }
                        } else {
                            tried = 1.f;                                // nothing cached: full MPR, a narrowphase task
                            ws.cv1[l] = (unsigned short)c1; ws.cv2[l] = (unsigned short)c2;
                        }
                    }
                } else ws.sepok[l] = 0;
            }
            // Stage 1
            float sA[3] = {0.f, 0.f, 0.f};
            if (job) kg_support(m, ws, gA, meshA, curA, dirA, orgA, marginA, sA);
            // Stage 2
            if (job == 1) {
                const int base = 4 * l;
                float w[3] = {sA[0], sA[1], sA[2]};
                const int cur = curA;
                KG_COUNT(26, 1);
                ws.cv2[l] = (unsigned short)cur;
                ws.sepok[l] = w[2] > margin; ws.slack[l] = w[2] - margin;      // clear of the floor by this much
                if (w[2] <= margin) {
                    const float thresh = 0.3f * rb2;
                    const float *vert = m.hull_vert + 4 * (size_t)m.mesh[mesh2].vert0;
                    const int *adjs = m.adj_start + m.mesh[mesh2].vert0;
                    int a0 = adjs[cur], a1 = adjs[cur + 1];
                    for (int c = -1; c < a1 - a0 && n < 4; c++) {
                        if (c >= 0) {
                            auto lv = KG_LDG4(vert + 4 * m.adj[a0 + c]);
                            float lvv[3] = {lv.x, lv.y, lv.z};
                            kg_mulmv(w, ws.gmat[g2], lvv); kg_add(w, w, ws.gpos[g2]);
                        }
                        float dist = w[2];
                        if (dist > margin) continue;
                        float cp[3] = {w[0], w[1], w[2] - 0.5f * dist};
                        bool skip = false;
                        for (int k = 0; k < n; k++) {
                            float df[3]; kg_sub(df, cp, ws.st_pos[base + k]);
                            if (kg_dot(df, df) < thresh * thresh) skip = true;
                        }
                        if (skip) continue;
                        int s = base + n;
                        ws.st_dist[s] = dist; kg_copy(ws.st_pos[s], cp);
                        ws.st_n[s][0] = 0; ws.st_n[s][1] = 0; ws.st_n[s][2] = 1;
                        ws.st_g1[s] = 0; ws.st_g2[s] = (unsigned char)g2;
                        n++;
                    }
                }
            } else if (job == 2) {
                const float nn[3] = {-dirA[0], -dirA[1], -dirA[2]};
                float s2[3];
                kg_support(m, ws, g2, mesh2, c2, nn, orgA, 0.5f * margin, s2);
                const float gap = kg_dot(s2, dirA) - kg_dot(sA, dirA);
                const bool separated = gap > 0.f;
                ws.slack[l] = gap;
                KG_COUNT(24, 1); KG_COUNT(25, separated ? 1 : 0);
                if (!separated) tried = 1.f;           // full MPR: a narrowphase task
                ws.cv1[l] = (unsigned short)curA; ws.cv2[l] = (unsigned short)c2;
            }
        }
        ws.red2[l] = tried;
        ws.cnt[l] = n;
    }
    KG_SYNC();
    unsigned mask = 0;
#if defined(__CUDA_ARCH__)
    mask = __ballot_sync(0xffffffffu, ws.red2[kg_lane] != 0.f);
#else
    (void)kg_lane;
    for (int k = 0; k < 32; k++) if (ws.red2[k] != 0.f) mask |= 1u << k;
#endif
    KG_PHASE(l) {
        if (l == 0) {
            int nt = 0;
            for (unsigned mm = mask; mm; mm &= mm - 1) nt++;
            ws.ntask = nt; ws.stat_mpr += nt;
        }
    }
    KG_SYNC();
    return mask;
}

// Contact compaction (after the narrowphase kernel has filed its hits): contact groups 9.. reference staging slots
template <class WS> KG_DEV void kg_compaction(const KgModel &m, WS &ws, int kg_lane) {
    (void)m;
    KG_PHASE(l) {
        int off, total;
        kg_prefix_int(ws.cnt, l, off, total);
        int base = l < 8 ? 4 * l : 32 + (l - 8);
        for (int i = 0; i < ws.cnt[l]; i++) {
            int g = 9 + off + i;
            if (g < 9 + WS::MAXCON) { KG_GP(ws, g, 7) = (float)(base + i); KG_GP(ws, g, 3) = 3.f; }
        }
        if (l == 0) {
            ws.ncon_total = total;
            ws.stat_pad += total;              // contacts of this substep (reported per substep by kg_get_stats)
            int nc = total > WS::MAXCON ? WS::MAXCON : total;
            ws.ngroup = 9 + nc;
            if (total > WS::MAXCON) ws.stat_overflow += 1;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P6 constraint rows
KG_DEV float kg_impedance(const KgModel &m, float r) {
    float x = fabsf(r) / m.width;
    if (x >= 1.f) return m.dmax;
    float y = x <= 0.5f ? 2.f * x * x : 1.f - 2.f * (1.f - x) * (1.f - x);
    return m.dmin + y * (m.dmax - m.dmin);
}
// translational Jacobian row of the point p on the body carrying geom g, projected on f, ADDED with sign sg
// (f by value: a pointer into the caller's registers would put them on the stack)
template <class WS> KG_NOINL void kg_jac_add(const WS &ws, int g, const float *p, float fx, float fy, float fz, float sg, float *row) {
    if (g == 0) return;
    const float f[3] = {fx, fy, fz};
    if (g == 8) {
        float r[3], t[3];
        for (int c = 0; c < 3; c++) { row[9 + c] += sg * f[c]; r[c] = p[c] - ws.qpos[9 + c]; }
        // rot dof i: f . ((R e_i) x r) = (R e_i) . (r x f)
        kg_cross(t, r, f);
        for (int i = 0; i < 3; i++) row[12 + i] += sg * (ws.oR[i] * t[0] + ws.oR[3 + i] * t[1] + ws.oR[6 + i] * t[2]);
        return;
    }
    for (int i = 0; i < 3; i++) row[i] += sg * kg_dot(f, ws.sax[i]);
    if (g == 1) return;
    int k = (g - 2) >> 1;
    float a[3] = {ws.B[k][2], ws.B[k][5], ws.B[k][8]}, r[3], t[3];
    kg_sub(r, p, ws.p1[k]); kg_cross(t, a, r);
    row[3 + 2 * k] += sg * kg_dot(f, t);
    if (g & 1) { kg_sub(r, p, ws.p2[k]); kg_cross(t, a, r); row[4 + 2 * k] += sg * kg_dot(f, t); }
}

template <class WS> KG_DEV void kg_make_rows(const KgModel &m, WS &ws, int kg_lane) {
    KG_PHASE(l) {
        const float K = m.solK, Bd = m.solB;
        if (l < 9) {      // the tendon equalities and the joint limits: a lane per group
            const int g = l;
            if (g < 3) {
                float *row = ws.J + g * KG_JS;
                for (int j = 0; j < KG_NV; j++) row[j] = 0.f;
                row[3 + 2 * g] = 1.f; row[4 + 2 * g] = -2.f;
                float pos = ws.qpos[3 + 2 * g] - 2.f * ws.qpos[4 + 2 * g];
                float vel = ws.qvel[3 + 2 * g] - 2.f * ws.qvel[4 + 2 * g];
                float imp = kg_impedance(m, pos);
                float R = fmaxf((1.f - imp) / imp * m.tendon_invw[g], 1e-15f);
                KG_GP(ws, g, 3) = 1.f; KG_GD(ws, g) = 1.f / R; KG_GAREF(ws, g, 0) = -Bd * vel - K * imp * pos;
            } else if (g < 9) {
                int a = g - 3, j = a < 3 ? a : 3 + 2 * (a - 3);
                float lo = a < 3 ? m.slide_range[2 * a] : m.finger_range[2 * (a - 3)];
                float hi = a < 3 ? m.slide_range[2 * a + 1] : m.finger_range[2 * (a - 3) + 1];
                float q = ws.qpos[j], sgn = 0.f, pos = 0.f;
                if (q - lo < 0.f) { sgn = 1.f; pos = q - lo; } else if (hi - q < 0.f) { sgn = -1.f; pos = hi - q; }
                float *row = ws.J + g * KG_JS;
                for (int jj = 0; jj < KG_NV; jj++) row[jj] = 0.f;
                if (sgn != 0.f) {
                    row[j] = sgn;
                    float imp = kg_impedance(m, pos);
                    float R = fmaxf((1.f - imp) / imp * m.dof_invw[j], 1e-15f);
                    KG_GP(ws, g, 3) = 2.f; KG_GD(ws, g) = 1.f / R; KG_GAREF(ws, g, 0) = -Bd * sgn * ws.qvel[j] - K * imp * pos;
                } else KG_GP(ws, g, 3) = 0.f;
            }
        }
        // contacts: a lane per BASE ROW (normal, tangent 1, tangent 2 of each contact) -- lanes 9..31 take the first 23 rows,
        // lanes 0..8 (busy with the tendon / limit groups above) the next 9, and so on: the Jacobian rows of a contact are
        // built side by side instead of one after the other.  The row's velocity J.v waits in the row's last pad word for the
        // phase below (bdot and the rest of the solver scratch still ARE the contact staging read here).
        const int nb = kg_nbase(ws);
        for (int b = 9 + ((l + 23) & 31); b < nb; b += 32) {
            const int g = 9 + (b - 9) / 3, k = (b - 9) % 3;
            const int s = kg_gslot(ws, g), g1 = ws.st_g1[s], g2 = ws.st_g2[s];
            float n[3], t1[3], t2[3];
            kg_copy(n, ws.st_n[s]);
            kg_make_frame(n, t1, t2);
            const float fx = k == 0 ? n[0] : (k == 1 ? t1[0] : t2[0]), fy = k == 0 ? n[1] : (k == 1 ? t1[1] : t2[1]), fz = k == 0 ? n[2] : (k == 1 ? t1[2] : t2[2]);
            float *row = ws.J + b * KG_JS;
            for (int j = 0; j < KG_NV; j++) row[j] = 0.f;
            const float *p = ws.st_pos[s];
            kg_jac_add(ws, g2, p, fx, fy, fz, 1.f, row); kg_jac_add(ws, g1, p, fx, fy, fz, -1.f, row);
            float v = 0.f, rr[16], qv[16];
            kg_ld16(row, rr); kg_ld16(ws.qvel, qv);
#pragma unroll
            for (int j = 0; j < KG_NV; j++) v += rr[j] * qv[j];
            row[KG_JS - 1] = v;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        const float K = m.solK, Bd = m.solB;
        for (int g = 9 + l; g < ws.ngroup; g += 32) {
            const int s = kg_gslot(ws, g), g1 = ws.st_g1[s], g2 = ws.st_g2[s], b = 9 + 3 * (g - 9);
            const float vn = ws.J[b * KG_JS + KG_JS - 1], v1 = ws.J[(b + 1) * KG_JS + KG_JS - 1], v2 = ws.J[(b + 2) * KG_JS + KG_JS - 1];
            float mu = m.pair_mu[s < 32 ? (s >> 2) : 8 + (s - 32)], dist = ws.st_dist[s], r = dist - m.margin;   // staging slot -> pair lane
            float tran = (g1 == 8 ? m.shape[ws.shape].invw_tran : m.body_invw[g1]) + (g2 == 8 ? m.shape[ws.shape].invw_tran : m.body_invw[g2]);
            float imp = kg_impedance(m, r);
            float R = fmaxf((1.f - imp) / imp * (tran + mu * mu * tran), 1e-15f);
            float Rpy = fmaxf(2.f * mu * mu / m.impratio * R, 1e-15f);
            float Dv = 1.f / Rpy, kr = K * imp * r;
            KG_GMU(ws, g) = mu;
            KG_GD(ws, g) = Dv;
            KG_GAREF(ws, g, 0) = -Bd * (vn + mu * v1) - kr; KG_GAREF(ws, g, 1) = -Bd * (vn - mu * v1) - kr;
            KG_GAREF(ws, g, 2) = -Bd * (vn + mu * v2) - kr; KG_GAREF(ws, g, 3) = -Bd * (vn - mu * v2) - kr;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P7 Newton solver
// Scalar rows of group g at the point described by base-row products bd[]:
//   tendon/limit: jar = bd[g] - aref ;  contact c: jar_k = bd[n] +- mu bd[t] - aref_k
template <class WS> KG_DEV int kg_group_rows(const WS &ws, int g, const float *bd, float *jar) {
    int tp = kg_gtype(ws, g);
    if (tp == 0) return 0;
    if (tp < 3) { jar[0] = bd[g] - KG_GAREF(ws, g, 0); return 1; }
    int b = 9 + 3 * (g - 9);
    float mu = KG_GMU(ws, g), n = bd[b], t1 = mu * bd[b + 1], t2 = mu * bd[b + 2];
    jar[0] = n + t1 - KG_GAREF(ws, g, 0); jar[1] = n - t1 - KG_GAREF(ws, g, 1);
    jar[2] = n + t2 - KG_GAREF(ws, g, 2); jar[3] = n - t2 - KG_GAREF(ws, g, 3);
    return 4;
}
template <class WS> KG_DEV int kg_nbase(const WS &ws) { return 9 + 3 * (ws.ngroup - 9); }

// base-row products  out[b] = J_b . x   (lane per base row)
template <class WS> KG_NOINL void kg_base_dots(WS &ws, const float *x, float *out, int kg_lane) {
    KG_PHASE(l) {
        int nb = kg_nbase(ws);
        float xv[16];
        kg_ld16(x, xv);
        for (int b = l; b < nb; b += 32) {
            float row[16];
            kg_ld16(ws.J + b * KG_JS, row);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < KG_NV; j++) s += row[j] * xv[j];
            out[b] = s;
        }
    }
    KG_SYNC();
}

// the same for two vectors at once (one pass over the rows)
template <class WS> KG_NOINL void kg_base_dots2(WS &ws, const float *x1, float *out1, const float *x2, float *out2, int kg_lane) {
    KG_PHASE(l) {
        int nb = kg_nbase(ws);
        float xv[16], yv[16];
        kg_ld16(x1, xv); kg_ld16(x2, yv);
        for (int b = l; b < nb; b += 32) {
            float row[16];
            kg_ld16(ws.J + b * KG_JS, row);
            float s = 0.f, t = 0.f;
#pragma unroll
            for (int j = 0; j < KG_NV; j++) { s += row[j] * xv[j]; t += row[j] * yv[j]; }
            out1[b] = s; out2[b] = t;
        }
    }
    KG_SYNC();
}

// y <- (M + diag_add I)^-1 x in closed form (y != x: lanes read x while others already write y).  M is block diagonal {hand 9x9, object 6x6}; the hand block is
// [mu I3, C; C^T, blockdiag(A_1,A_2,A_3)] (slides couple to each finger's 2x2 hinge block, fingers do not couple to
// each other), the object block is [m' I3, T; T^T, Io'].  Both are eliminated by 3x3 Schur complements: lanes 0..2
// own a finger, lane 3 the object.  (The reference engine does a generic sparse L^T D L here.)
KG_DEV void kg_inv3_sym(const float *S, float *Si) {   // inverse of a symmetric 3x3 given as xx,yy,zz,xy,xz,yz
    float c0 = S[1] * S[2] - S[5] * S[5], c1 = S[4] * S[5] - S[3] * S[2], c2 = S[3] * S[5] - S[4] * S[1];
    float det = S[0] * c0 + S[3] * c1 + S[4] * c2, id = 1.0f / det;
    Si[0] = c0 * id; Si[3] = c1 * id; Si[4] = c2 * id;
    Si[1] = (S[0] * S[2] - S[4] * S[4]) * id; Si[5] = (S[3] * S[4] - S[0] * S[5]) * id;
    Si[2] = (S[0] * S[1] - S[3] * S[3]) * id;
}
template <class WS> KG_NOINL void kg_minv(const KgModel &m, WS &ws, float diag_add, const float *x, float *y, int kg_lane) {
    KG_PHASE(l) {
        if (l < 3) {
            int h1 = 3 + 2 * l, h2 = 4 + 2 * l;
            float a11 = ws.M[h1 * KG_MS + h1] + diag_add, a12 = ws.M[h1 * KG_MS + h2], a22 = ws.M[h2 * KG_MS + h2] + diag_add;
            float id = 1.0f / (a11 * a22 - a12 * a12);
            float i11 = a22 * id, i12 = -a12 * id, i22 = a11 * id;
            float f1 = x[h1], f2 = x[h2];
            float y1 = i11 * f1 + i12 * f2, y2 = i12 * f1 + i22 * f2;   // A^-1 f_h
            float g[3][2], S[6], r[3];
            for (int i = 0; i < 3; i++) {
                float c1 = ws.M[i * KG_MS + h1], c2 = ws.M[i * KG_MS + h2];
                g[i][0] = c1 * i11 + c2 * i12; g[i][1] = c1 * i12 + c2 * i22;   // G = C A^-1
                r[i] = c1 * y1 + c2 * y2;
            }
            const int ii[6] = {0, 1, 2, 0, 0, 1}, jj[6] = {0, 1, 2, 1, 2, 2};
            for (int e = 0; e < 6; e++) S[e] = g[ii[e]][0] * ws.M[jj[e] * KG_MS + h1] + g[ii[e]][1] * ws.M[jj[e] * KG_MS + h2];
            for (int e = 0; e < 6; e++) ws.mv_S[l][e] = S[e];
            for (int i = 0; i < 3; i++) ws.mv_r[l][i] = r[i];
            ws.mv_y[l][0] = y1; ws.mv_y[l][1] = y2;
            ws.mv_i[l][0] = i11; ws.mv_i[l][1] = i12; ws.mv_i[l][2] = i22;
        } else if (l == 3) {
            // object: x_t = (f_t - T x_r)/m',  (Io' - T^T T / m') x_r = f_r - T^T f_t / m'
            float mp = ws.M[9 * KG_MS + 9] + diag_add, im = 1.0f / mp;
            // the Schur complement Io' - T^T T / m' does not depend on the pose (T = -m R [c]x, R^T R = I): its inverse is a
            // per-shape constant of the model card, one for each of the two diagonal shifts this function is called with
            float T[9], rhs[3];
            for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) T[3 * j + i] = ws.M[(9 + j) * KG_MS + 12 + i];
            float ft[3] = {x[9], x[10], x[11]};
            for (int i = 0; i < 3; i++) rhs[i] = x[12 + i] - (T[i] * ft[0] + T[3 + i] * ft[1] + T[6 + i] * ft[2]) * im;
            const float *Si = m.shape[ws.shape].Sinv[diag_add != 0.f ? 1 : 0];
            float xr[3] = {Si[0] * rhs[0] + Si[3] * rhs[1] + Si[4] * rhs[2], Si[3] * rhs[0] + Si[1] * rhs[1] + Si[5] * rhs[2],
                           Si[4] * rhs[0] + Si[5] * rhs[1] + Si[2] * rhs[2]};
            for (int j = 0; j < 3; j++) y[9 + j] = (ft[j] - (T[3 * j] * xr[0] + T[3 * j + 1] * xr[1] + T[3 * j + 2] * xr[2])) * im;
            for (int i = 0; i < 3; i++) y[12 + i] = xr[i];
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l < 3) {
            // every finger lane redundantly solves the 3x3 slide system, then back-substitutes its own hinges
            float mu = ws.M[0] + diag_add;
            float S[6], rhs[3], Si[6];
            for (int e = 0; e < 6; e++) S[e] = (e < 3 ? mu : 0.f) - ws.mv_S[0][e] - ws.mv_S[1][e] - ws.mv_S[2][e];
            for (int i = 0; i < 3; i++) rhs[i] = x[i] - ws.mv_r[0][i] - ws.mv_r[1][i] - ws.mv_r[2][i];
            kg_inv3_sym(S, Si);
            float xs[3] = {Si[0] * rhs[0] + Si[3] * rhs[1] + Si[4] * rhs[2], Si[3] * rhs[0] + Si[1] * rhs[1] + Si[5] * rhs[2],
                           Si[4] * rhs[0] + Si[5] * rhs[1] + Si[2] * rhs[2]};
            int h1 = 3 + 2 * l, h2 = 4 + 2 * l;
            float t1 = 0.f, t2 = 0.f;   // C^T xs
            for (int i = 0; i < 3; i++) { t1 += ws.M[i * KG_MS + h1] * xs[i]; t2 += ws.M[i * KG_MS + h2] * xs[i]; }
            y[h1] = ws.mv_y[l][0] - (ws.mv_i[l][0] * t1 + ws.mv_i[l][1] * t2);
            y[h2] = ws.mv_y[l][1] - (ws.mv_i[l][1] * t1 + ws.mv_i[l][2] * t2);
            y[l] = l == 0 ? xs[0] : (l == 1 ? xs[1] : xs[2]);      // (every finger lane holds the same xs)
        }
    }
    KG_SYNC();
}

template <class WS> KG_DEV void kg_solve(const KgModel &m, WS &ws, int kg_lane) {
    const float tol = 1e-8f;
    // ---- unconstrained acceleration
    kg_minv(m, ws, 0.f, ws.fs, ws.qas, kg_lane);
    // ---- warm start: the cheaper of qacc_warmstart and qacc_smooth
    kg_base_dots2(ws, ws.qwarm, ws.bdot, ws.qas, ws.bdir, kg_lane);
    // per lane, for the group(s) it owns: the residuals jar[4] of the rows at the current point (rows the group does not have
    // are zero), their slopes jv[4] along the Newton direction (line search), D (0 = no rows), equality flag; lss = the
    // residuals at qacc_smooth until the start point is chosen
    constexpr int NG = (9 + WS::MAXCON + 31) / 32;
    KG_LANEVAR(float, lsq, NG * 9);
    KG_LANEVAR(float, lss, NG * 4);
    KG_LANEVAR(int, lse, NG);
    KG_PARTIAL(pcw); KG_PARTIAL(pcs);
    KG_PHASE(l) {
        float cw = 0.f, cs = 0.f;
        if (l < KG_NV) {
            float s = 0.f, mr[16], qw[16], qs[16];
            kg_ld16(ws.M + l * KG_MS, mr); kg_ld16(ws.qwarm, qw); kg_ld16(ws.qas, qs);
#pragma unroll
            for (int j = 0; j < KG_NV; j++) s += mr[j] * (qw[j] - qs[j]);
            ws.Mdq[l] = s;
            cw = 0.5f * s * (ws.qwarm[l] - ws.qas[l]);
        }
#pragma unroll
        for (int gi = 0; gi < NG; gi++) {
            const int g = l + 32 * gi;
            float jw[4] = {0, 0, 0, 0}, js[4] = {0, 0, 0, 0};
            int n = 0;
            if (g < 9 + WS::MAXCON) { n = kg_group_rows(ws, g, ws.bdot, jw); kg_group_rows(ws, g, ws.bdir, js); }
            const bool eq = n && kg_gtype(ws, g) == 1;
            const float Dg = n ? KG_GD(ws, g) : 0.f;
            for (int k = 0; k < n; k++) if (eq || jw[k] < 0.f) cw += 0.5f * Dg * jw[k] * jw[k];
            for (int k = 0; k < n; k++) if (eq || js[k] < 0.f) cs += 0.5f * Dg * js[k] * js[k];
            float *q = KG_LV(lsq, l) + 9 * gi, *qs = KG_LV(lss, l) + 4 * gi;
#pragma unroll
            for (int k = 0; k < 4; k++) { q[k] = k < n ? jw[k] : 0.f; qs[k] = k < n ? js[k] : 0.f; }
            q[8] = Dg;
            KG_LV(lse, l)[gi] = eq;
        }
        KG_PUT(pcw, l, cw); KG_PUT(pcs, l, cs);
    }
    KG_SYNC();                                                  // (Mdq written above is read below by other lanes? no: own lane; kept for bdot/bdir)
    KG_PHASE(l) {
        float cw = KG_SUM(pcw, l), cs = KG_SUM(pcs, l);
        bool use_smooth = cs < cw;
        if (l < KG_NV) {
            ws.qacc[l] = use_smooth ? ws.qas[l] : ws.qwarm[l];
            if (use_smooth) ws.Mdq[l] = 0.f;
        }
        if (use_smooth) {
            int nb = kg_nbase(ws);
            for (int b = l; b < nb; b += 32) ws.bdot[b] = ws.bdir[b];
#pragma unroll
            for (int k = 0; k < 4 * NG; k++) KG_LV(lsq, l)[9 * (k / 4) + (k & 3)] = KG_LV(lss, l)[k];
        }
    }
    KG_SYNC();
    const float scale = m.shape[ws.shape].scale;
    KG_PROBE(7);
    // the solve's control state lives in registers: every lane of a warp total holds the same value, so every lane takes
    // the same decisions without a broadcast through shared memory
    float prev_cost = 3.0e38f;
    int done = 0, n_iter = 0;
    for (int iter = 0; iter <= KG_NEWTON_MAXIT; iter++) {
        // ---- (a) rows at the current point: cost and the weights w = D * jar of the active rows
        KG_PARTIAL(pc);
        KG_PHASE(l) {
            float c = 0.f;
            if (l < KG_NV) c = 0.5f * ws.Mdq[l] * (ws.qacc[l] - ws.qas[l]);
#pragma unroll
            for (int gi = 0; gi < NG; gi++) {
                const int g = l + 32 * gi;
                float jar[4], w[4] = {0, 0, 0, 0}, a[4] = {0, 0, 0, 0};
                float *q = KG_LV(lsq, l) + 9 * gi;
                const float Dg = q[8];                                  // (D and the equality flag do not change during the solve)
                const bool eq = KG_LV(lse, l)[gi] != 0;
                const int n = Dg != 0.f ? (g < 9 ? 1 : 4) : 0;
                if (iter > 0 && n) {                                    // (iteration 0 starts from the warm-start evaluation above)
                    float jn[4] = {0, 0, 0, 0};
                    kg_group_rows(ws, g, ws.bdot, jn);
#pragma unroll
                    for (int k = 0; k < 4; k++) q[k] = k < n ? jn[k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) jar[k] = q[k];
#pragma unroll
                for (int k = 0; k < 4; k++) if (k < n && (eq || jar[k] < 0.f)) { a[k] = Dg; w[k] = a[k] * jar[k]; c += 0.5f * w[k] * jar[k]; }
                if (n == 1) { ws.bw[g] = w[0]; KG_GAREF(ws, g, 1) = a[0]; }
                else if (n == 4) {
                    int b = 9 + 3 * (g - 9);
                    float mu = KG_GMU(ws, g);
                    ws.bw[b] = w[0] + w[1] + w[2] + w[3]; ws.bw[b + 1] = mu * (w[0] - w[1]); ws.bw[b + 2] = mu * (w[2] - w[3]);
                    // the group's 3x3 block of J^T D J in the (n, t1, t2) base: [A B1 B2; B1 C1 0; B2 0 C2], for phase (b)
                    KG_GH(ws, g, 0) = a[0] + a[1] + a[2] + a[3];
                    KG_GH(ws, g, 1) = mu * (a[0] - a[1]); KG_GH(ws, g, 2) = mu * mu * (a[0] + a[1]);
                    KG_GH(ws, g, 3) = mu * (a[2] - a[3]); KG_GH(ws, g, 4) = mu * mu * (a[2] + a[3]);
                } else if (g < 9) ws.bw[g] = 0.f;
            }
            KG_PUT(pc, l, c);
        }
        KG_SYNC();                                              // (bw and the groups' coefficients go to phase (b))
        {
            float cost = 0.f;
            KG_PHASE(l) { cost = KG_SUM(pc, l); }
            // MuJoCo's test is scale*improvement < 1e-8; in fp32 the cost itself only carries ~2^-23 relative
            // precision, so an "improvement" below that noise floor is treated as converged as well.
            const float improvement = scale * (prev_cost - cost);
            if (iter > 0 && improvement < tol + 2.4e-7f * scale * fabsf(cost)) done = 1;
            if (iter == KG_NEWTON_MAXIT) done = 1;
            prev_cost = cost;
        }
        KG_PROBE(8);
        if (done) break;
        // ---- (b) gradient and Hessian.  Two lanes per row of H: lane r owns the columns 0..7 of row r, lane 16 + r the
        // columns 8..14 (each entry accumulates over the groups in the same order whoever owns it); lane r also forms
        // the gradient entry.  The groups' coefficients come from phase (a).
        KG_PARTIAL(pgs); KG_PARTIAL(pga);
        KG_PHASE(l) {
            float gsq = 0.f, gab = 0.f;
            const int r = l & 15, j0 = (l >> 4) * 8;
            if (r < KG_NV) {
                float gr = ws.Mdq[r];
                gab = fabsf(gr);
                float h[8];
                kg_ld8(ws.M + r * KG_MS + j0, h);
                // scalar rows: tendon k has entries on the two hinges of finger k only, a limit row on its own dof only, so a
                // row of H meets at most one of each (the others would add exact zeros); then every contact in use
                auto scalar_row = [&](int g) {
                    const float *rowp = ws.J + g * KG_JS;
                    float ji = rowp[r];
                    float t = ji * ws.bw[g];
                    gr += t; gab += fabsf(t);
                    const float da = KG_GAREF(ws, g, 1);   // D * active stored by (a)
                    float u = ji * da;
                    if (da != 0.f) {
                        float row[8];
                        kg_ld8(rowp + j0, row);
#pragma unroll
                        for (int j = 0; j < 8; j++) h[j] += u * row[j];
                    }
                };
                if (r >= 3 && r < 9) scalar_row((r - 3) >> 1);
                {
                    const int gl = r < 3 ? 3 + r : ((r < 9 && (r & 1)) ? 6 + ((r - 3) >> 1) : -1);
                    if (gl >= 0 && kg_gtype(ws, gl) == 2) scalar_row(gl);
                }
#pragma unroll 2
                for (int g = 9; g < ws.ngroup; g++) {
                    int b = 9 + 3 * (g - 9);
                    const float *rn = ws.J + b * KG_JS, *r1 = rn + KG_JS, *r2 = rn + 2 * KG_JS;
                    float jn = rn[r], j1 = r1[r], j2 = r2[r];
                    float t0 = jn * ws.bw[b], t1 = j1 * ws.bw[b + 1], t2 = j2 * ws.bw[b + 2];
                    gr += t0 + t1 + t2; gab += fabsf(t0) + fabsf(t1) + fabsf(t2);
                    const float A = KG_GH(ws, g, 0), B1 = KG_GH(ws, g, 1), C1 = KG_GH(ws, g, 2), B2 = KG_GH(ws, g, 3), C2 = KG_GH(ws, g, 4);
                    if (A != 0.f) {
                        float u = A * jn + B1 * j1 + B2 * j2, v1 = B1 * jn + C1 * j1, v2 = B2 * jn + C2 * j2;
                        float vn[8], w1[8], w2[8];
                        kg_ld8(rn + j0, vn); kg_ld8(r1 + j0, w1); kg_ld8(r2 + j0, w2);
#pragma unroll
                        for (int j = 0; j < 8; j++) h[j] += u * vn[j] + v1 * w1[j] + v2 * w2[j];
                    }
                }
                if (j0 == 0) { ws.grad[r] = gr; gsq = gr * gr; }
                else { gab = 0.f; h[7] = 0.f; }           // (column 15 is padding; the gradient is lane r's)
                kg_st8(ws.L + r * KG_MS + j0, h);
            }
            KG_PUT(pgs, l, gsq); KG_PUT(pga, l, gab * gab);
        }
        KG_SYNC();                                              // (grad, L)
        {
            float gn = 0.f, ga = 0.f;
            KG_PHASE(l) {
                gn = KG_SUM(pgs, l); ga = KG_SUM(pga, l);
                if (l < KG_NV) ws.dir[l] = -ws.grad[l];
            }
            // gradient test: MuJoCo's absolute 1e-8, or the fp32 rounding floor of the sum that formed the gradient
            if (scale * sqrtf(gn) < tol || gn < 1.4e-13f * ga) done = 1;
        }
        KG_SYNC();
        KG_PROBE(9);
        if (done) break;
        kg_chol_factor_solve(ws.L, ws.dir, kg_lane);
        KG_PROBE(10);
        // ---- (c) directional products
        // (one phase: base-row products J_b . dir by a lane per base row, M dir by a lane per dof)
        KG_PARTIAL(pp1); KG_PARTIAL(pp2);
        KG_PHASE(l) {
            float xv[16];
            kg_ld16(ws.dir, xv);
            const int nb = kg_nbase(ws);
            for (int b = l; b < nb; b += 32) {
                float row[16];
                kg_ld16(ws.J + b * KG_JS, row);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < KG_NV; j++) s += row[j] * xv[j];
                ws.bdir[b] = s;
            }
            float p1 = 0.f, p2 = 0.f;
            if (l < KG_NV) {
                float mr[16];
                kg_ld16(ws.M + l * KG_MS, mr);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < KG_NV; j++) s += mr[j] * xv[j];
                ws.Md[l] = s;
                const float dl = ws.dir[l];
                p1 = ws.Mdq[l] * dl; p2 = s * dl;
            }
            KG_PUT(pp1, l, p1); KG_PUT(pp2, l, p2);
        }
        KG_SYNC();                                              // (bdir, Md)
        // warp totals are carried in registers: every lane of kg_total's result holds the same value (in the host emulation
        // the phase body runs once per lane and assigns the same value each time)
        float gp1 = 0.f, gp2 = 0.f;                         // Gauss part of phi'(0), phi''
        KG_PHASE(l) { gp1 = KG_SUM(pp1, l); gp2 = KG_SUM(pp2, l); }
        KG_PROBE(11);
        // ---- (d) exact line search: root of the piecewise-linear phi'(alpha).  dir is the exact Newton direction of
        // the current quadratic piece, so the first candidate is alpha = 1; if no row changes activity on [0,1] that is
        // the exact minimiser (phi'(1) = 0 up to rounding) and the solve is finished.
        {
            float alpha = 1.f, lo = 0.f, hi = -1.f, g0abs = 0.f, best_alpha = 1.f;
            int exact = 0;
            // what a lane's rows contribute does not change between the trial steps: the residuals at the current point are
            // phase (a)'s, their slopes along dir (jv) are taken once here; rows a group does not have add exact zeros
            KG_PHASE(l) {
#pragma unroll
                for (int gi = 0; gi < NG; gi++) {
                    const int g = l + 32 * gi;
                    float *q = KG_LV(lsq, l) + 9 * gi;
                    float jv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (q[8] != 0.f) {
                        if (g < 9) jv[0] = ws.bdir[g];
                        else {
                            int b = 9 + 3 * (g - 9);
                            float mu = KG_GMU(ws, g), nn = ws.bdir[b], t1 = mu * ws.bdir[b + 1], t2 = mu * ws.bdir[b + 2];
                            jv[0] = nn + t1; jv[1] = nn - t1; jv[2] = nn + t2; jv[3] = nn - t2;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) q[4 + k] = jv[k];
                }
            }
            for (int ls = 0; ls < KG_LS_MAXIT; ls++) {
                if (kg_lane == 0) KG_COUNT(27, 1);
                KG_PARTIAL(pd1); KG_PARTIAL(pd2); KG_PARTIAL(pd0); KG_PARTIAL_I(pch);
                KG_PHASE(l) {
                    float d1 = 0.f, d2 = 0.f, d0 = 0.f;
                    int nchg = 0;
#pragma unroll
                    for (int gi = 0; gi < NG; gi++) {
                        const float *q = KG_LV(lsq, l) + 9 * gi;
                        const bool eq = KG_LV(lse, l)[gi] != 0;
                        const float Dl = q[8];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float jr = q[k], jvk = q[4 + k];
                            float x = jr + alpha * jvk;
                            float dj = Dl * jvk;
                            bool act0 = eq || jr < 0.f, act = eq || x < 0.f;
                            if (act) { d1 += dj * x; d2 += dj * jvk; }
                            if (act0) d0 += dj * jr;
                            nchg += (act != act0);
                        }
                    }
                    KG_PUT(pd1, l, d1); KG_PUT(pd2, l, d2); KG_PUT(pd0, l, d0); KG_PUT(pch, l, nchg);
                }
                float t_lo = 0.f, t_hi = 0.f, t_d0 = 0.f;
                int t_chg = 0;
                KG_PHASE(l) {
                    t_lo = KG_SUM(pd1, l); t_hi = KG_SUM(pd2, l);
                    if (ls == 0) { t_d0 = KG_SUM(pd0, l); t_chg = KG_SUM_I(pch, l); }
                }
                float gA = gp1 + alpha * gp2 + t_lo, hA = gp2 + t_hi;
                if (ls == 0) {
                    g0abs = fabsf(gp1 + t_d0);                    // |phi'(0)|
                    if (t_chg == 0) { exact = 1; best_alpha = 1.f; break; }
                }
                best_alpha = alpha;
                if (fabsf(gA) <= 1e-6f * g0abs) break;
                if (gA < 0.f) lo = alpha; else hi = alpha;
                float nxt = alpha - gA / hA;
                if (!(nxt > lo) || (hi > 0.f && !(nxt < hi))) nxt = hi > 0.f ? 0.5f * (lo + hi) : 2.f * alpha;
                if (fabsf(nxt - alpha) <= 1e-7f * fmaxf(1.f, fabsf(alpha))) { best_alpha = nxt; break; }
                alpha = nxt;
                best_alpha = alpha;
            }
            const float al = best_alpha;
            KG_PHASE(l) {
                if (l < KG_NV) { ws.qacc[l] += al * ws.dir[l]; ws.Mdq[l] += al * ws.Md[l]; }
                int nb = kg_nbase(ws);
                for (int b = l; b < nb; b += 32) ws.bdot[b] += al * ws.bdir[b];
            }
            n_iter = iter + 1;
            KG_SYNC();
            KG_PROBE(12);
            // a full step on an unchanged active set is the exact minimiser: nothing the next pass over the rows would
            // compute is needed any more (the total force below is M qacc)
            if (exact && iter + 1 < KG_NEWTON_MAXIT) break;
        }
    }
    // ---- total force qfrc_smooth + qfrc_constraint, warm start for the next substep.
    // At the minimiser the gradient M (qacc - qacc_smooth) - qfrc_constraint vanishes, so the total force IS M qacc =
    // qfrc_smooth + Mdq.  That form is what fp32 can deliver: the row form J^T D (J qacc - aref) subtracts two numbers of
    // size |aref| whose difference is 1e-4 of them in a fresh contact (each row loses 3-4 digits, all rows in the same
    // direction), while qacc itself is pinned to 1e-7 by the stiff Hessian.  Only a solve that ran into the iteration cap
    // is not at the minimiser; it keeps the row form, which is what the reference engine integrates in that case.
    KG_PHASE(l) {
        if (l < KG_NV) {
            float f = ws.fs[l];
            if (n_iter < KG_NEWTON_MAXIT) f += ws.Mdq[l];
            else {
                int nb = kg_nbase(ws);
                for (int b = 0; b < nb; b++) {
                    if (b < 9 && kg_gtype(ws, b) == 0) continue;
                    f -= ws.J[b * KG_JS + l] * ws.bw[b];
                }
            }
            ws.ftot[l] = f;
            ws.qwarm[l] = ws.qacc[l];
        }
        if (l == 0) { ws.s_iter = n_iter; ws.stat_newton += n_iter; }
#if !defined(__CUDA_ARCH__) && defined(KG_EMU_HIST)
        if (l == 0) fprintf(stderr, "NEWTON %d %d\n", ws.s_iter, ws.ngroup - 9);
#endif
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P9 Euler
template <class WS> KG_DEV void kg_euler(const KgModel &m, WS &ws, int kg_lane) {
    const float h = m.timestep;
    kg_minv(m, ws, h * m.damping, ws.ftot, ws.Md, kg_lane);     // implicit joint damping: (M + h B) a = f   (Md: free after the solve)
    KG_PHASE(l) {
        if (l < KG_NV) {
            float v = ws.qvel[l] + h * ws.Md[l];
            ws.qvel[l] = v;
            if (l < 12) ws.qpos[l] += h * v;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l == 0) {
            float w[3] = {ws.qvel[12], ws.qvel[13], ws.qvel[14]};
            float q[4] = {ws.qpos[12], ws.qpos[13], ws.qpos[14], ws.qpos[15]};
            float wn = sqrtf(kg_dot(w, w)), ang = h * wn;
            if (ang > 0.f) {
                float s, c;
                sincosf(0.5f * ang, &s, &c);
                float k = s / wn;
                float r[4] = {c, w[0] * k, w[1] * k, w[2] * k};
                float t[4] = {q[0] * r[0] - q[1] * r[1] - q[2] * r[2] - q[3] * r[3],
                              q[0] * r[1] + q[1] * r[0] + q[2] * r[3] - q[3] * r[2],
                              q[0] * r[2] - q[1] * r[3] + q[2] * r[0] + q[3] * r[1],
                              q[0] * r[3] + q[1] * r[2] - q[2] * r[1] + q[3] * r[0]};
                for (int c2 = 0; c2 < 4; c2++) q[c2] = t[c2];
            }
            float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            for (int c2 = 0; c2 < 4; c2++) ws.qpos[12 + c2] = q[c2] * inv;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ P8 rangefinder rays
// mj_ray restated (see oracle/kgo_rays.c for the literal version): own-body geoms excluded, plane hit
// from the front only, meshes culled by the geom_size box, then nearest raw-triangle hit via the BVH.
KG_DEV float kg_ray_box(const float *size, const float *p, const float *v) {
    float best = -1.f;
    for (int ax = 0; ax < 3; ax++) {
        if (fabsf(v[ax]) < 1e-15f) continue;
        int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
        for (int sg = -1; sg <= 1; sg += 2) {
            float x = (sg * size[ax] - p[ax]) / v[ax];
            if (x < 0.f) continue;
            float c1 = p[a1] + x * v[a1], c2 = p[a2] + x * v[a2];
            if (fabsf(c1) <= size[a1] && fabsf(c2) <= size[a2] && (best < 0.f || x < best)) best = x;
        }
    }
    return best;
}
KG_DEV float kg_ray_tri(const float *t, const float *p, const float *v, const float *b0, const float *b1) {
    float pl[3][2];
    for (int i = 0; i < 3; i++) {
        float dif[3] = {t[3 * i] - p[0], t[3 * i + 1] - p[1], t[3 * i + 2] - p[2]};
        pl[i][0] = kg_dot(b0, dif); pl[i][1] = kg_dot(b1, dif);
    }
    for (int c = 0; c < 2; c++) {
        if (pl[0][c] > 0 && pl[1][c] > 0 && pl[2][c] > 0) return -1.f;
        if (pl[0][c] < 0 && pl[1][c] < 0 && pl[2][c] < 0) return -1.f;
    }
    float A0 = pl[0][0] - pl[2][0], A1 = pl[1][0] - pl[2][0], A2 = pl[0][1] - pl[2][1], A3 = pl[1][1] - pl[2][1];
    float bb0 = -pl[2][0], bb1 = -pl[2][1];
    float det = A0 * A3 - A1 * A2;
    if (fabsf(det) < 1e-30f) return -1.f;
    float t0 = (A3 * bb0 - A1 * bb1) / det, t1 = (-A2 * bb0 + A0 * bb1) / det;
    if (t0 < 0 || t1 < 0 || t0 + t1 > 1) return -1.f;
    float e0[3], e1[3], e2[3], nrm[3];
    for (int c = 0; c < 3; c++) { e0[c] = t[c] - t[6 + c]; e1[c] = t[3 + c] - t[6 + c]; e2[c] = p[c] - t[6 + c]; }
    kg_cross(nrm, e0, e1);
    float den = kg_dot(v, nrm);
    if (fabsf(den) < 1e-30f) return -1.f;
    float x = -kg_dot(e2, nrm) / den;
    return x >= 0.f ? x : -1.f;
}
// ray (pnt, vec) in the frame of a mesh geom
KG_DEV void kg_ray_local(const float *gpos, const float *gmat, const float *pnt, const float *vec, float *lp, float *lv) {
    float dp[3];
    kg_sub(dp, pnt, gpos);
    kg_multmv(lp, gmat, dp); kg_multmv(lv, gmat, vec);
}
// nearest raw-triangle hit of the local ray (lp, lv) with a mesh, through its BVH; -1 = none.  The caller has already
// culled the ray against the geom_size box.
KG_DEV float kg_ray_mesh_bvh(const KgModel &m, int mesh, const float *lp, const float *lv) {
    float b0[3], b1[3];
    if (fabsf(lv[0]) < fabsf(lv[1]) && fabsf(lv[0]) < fabsf(lv[2])) { b0[0] = 0; b0[1] = lv[2]; b0[2] = -lv[1]; }
    else if (fabsf(lv[1]) < fabsf(lv[2])) { b0[0] = -lv[2]; b0[1] = 0; b0[2] = lv[0]; }
    else { b0[0] = lv[1]; b0[1] = -lv[0]; b0[2] = 0; }
    kg_normalize(b0);
    kg_cross(b1, b0, lv);
    kg_normalize(b1);
    const KgMesh &mi = m.mesh[mesh];
    const float *tv = m.tri + 9 * (size_t)mi.tri0;
    const float *node = m.bvh_node + 8 * (size_t)mi.bvh0;    // 32-byte nodes: (lo.xyz, first|left), (hi.xyz, count)
    float inv[3];
    for (int c = 0; c < 3; c++) inv[c] = fabsf(lv[c]) > 1e-30f ? 1.0f / lv[c] : (lv[c] >= 0 ? 1e30f : -1e30f);
    float best = -1.f;
    int stack[48], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        int n = stack[--sp];
        auto nlo = KG_LDG4(node + 8 * n); auto nhi = KG_LDG4(node + 8 * n + 4);
        float l3[3] = {nlo.x, nlo.y, nlo.z}, h3[3] = {nhi.x, nhi.y, nhi.z};
        float t0 = 0.f, t1 = best < 0.f ? 1e30f : best;
        bool hit = true;
        for (int c = 0; c < 3; c++) {
            float a = (l3[c] - lp[c]) * inv[c], b = (h3[c] - lp[c]) * inv[c];
            if (a > b) { float s = a; a = b; b = s; }
            t0 = fmaxf(t0, a); t1 = fminf(t1, b);
        }
        if (t0 > t1 * (1.f + 1e-5f) + 1e-7f) hit = false;
        if (!hit) continue;
        int first = (int)KG_F2U(nlo.w), cnt = (int)KG_F2U(nhi.w);
        if (cnt == 0) { if (sp < 46) { stack[sp++] = first; stack[sp++] = first + 1; } }
        else for (int i = 0; i < cnt; i++) {
            float x = kg_ray_tri(tv + 9 * (size_t)(first + i), lp, lv, b0, b1);
            if (x >= 0.f && (best < 0.f || x < best)) best = x;
        }
    }
    return best;
}
template <class WS> KG_NOINL void kg_site_pose(const KgModel &m, const WS &ws, int i, float *p, float *z) {
    int sb = m.site_body[i];
    float R[9], org[3];
    if (sb == 0) { for (int c = 0; c < 9; c++) R[c] = ws.Rh[c]; kg_copy(org, ws.hp); }
    else {
        int k = sb <= 3 ? sb - 1 : sb - 4;
        // body frames: recover R1 / R2 from the geom frames would need the inverse; rebuild from B and q
        float q1 = ws.sens_q[3 + k], q2 = ws.sens_q[6 + k], s, c;
        sincosf(sb <= 3 ? q1 : q1 + q2, &s, &c);
        const float *Bk = ws.B[k];
        for (int r = 0; r < 3; r++) { float b0 = Bk[3 * r], b1 = Bk[3 * r + 1]; R[3 * r] = c * b0 + s * b1; R[3 * r + 1] = -s * b0 + c * b1; R[3 * r + 2] = Bk[3 * r + 2]; }
        kg_copy(org, sb <= 3 ? ws.p1[k] : ws.p2[k]);
    }
    float t[3];
    kg_mulmv(t, R, m.site_pos + 3 * i); kg_add(p, org, t);
    kg_mulmv(z, R, m.site_z + 3 * i);
}
// The 17 rangefinders against the 8 mesh geoms are 136 (site, geom) pairs.  Phase 1: a lane per site (pose, floor hit);
// phase 2: the pairs strided over all 32 lanes, box cull, survivors appended to a hit list; phase 3: the survivors
// strided over all 32 lanes, BVH traversal, nearest hit per site by an integer min on the (non-negative) distances.
// Scratch lives in J (dead between the last solve and the obs staging).
#if defined(__CUDA_ARCH__)
#define KG_ATOMIC_INC(p) atomicAdd((p), 1)
#define KG_ATOMIC_MIN_U(p, v) atomicMin((p), (v))
#else
static inline int KG_ATOMIC_INC(int *p) { return (*p)++; }
static inline void KG_ATOMIC_MIN_U(unsigned *p, unsigned v) { if (v < *p) *p = v; }
#endif
#define KG_RAY_NONE 0x7f800000u     // +inf: no hit yet
template <class WS> KG_DEV void kg_rays(const KgModel &m, WS &ws, int kg_lane) {
    float *site = ws.J;                                                   // [KG_NSITE][6] origin, direction
    int *hits = reinterpret_cast<int *>(ws.J + 6 * KG_NSITE + 2);         // [<= 8 KG_NSITE] surviving pair ids
    unsigned *rmin = reinterpret_cast<unsigned *>(KG_RANGE(ws));
    static_assert((6 * KG_NSITE + 2 + 8 * KG_NSITE) * 4 <= KgWSLight::NBASE * KG_JS * 4, "ray scratch must fit into J");
    KG_PHASE(l) {
        if (l < KG_NSITE) {
            float p[3], v[3];
            kg_site_pose(m, ws, l, p, v);
            for (int c = 0; c < 3; c++) { site[6 * l + c] = p[c]; site[6 * l + 3 + c] = v[c]; }
            float best = -1.f;
            if (v[2] < -1e-15f) {
                float x = -p[2] / v[2];
                if (x >= 0.f && fabsf(p[0] + x * v[0]) <= 2.f && fabsf(p[1] + x * v[1]) <= 2.f) best = x;
            }
            rmin[l] = best < 0.f ? KG_RAY_NONE : KG_F2U(fabsf(best));
        }
        if (l == 0) ws.s_iter = 0;
    }
    KG_SYNC();
    KG_PHASE(l) {
        for (int pr = l; pr < 8 * KG_NSITE; pr += 32) {
            const int i = pr >> 3, g = 1 + (pr & 7);
            const int sb = m.site_body[i];
            const int own = sb == 0 ? 1 : (sb <= 3 ? 2 * sb : 2 * (sb - 3) + 1);
            if (g == own) continue;
            const float *size = g == 8 ? m.shape[ws.shape].size : (g == 1 ? m.hand_size : ((g & 1) ? m.hand_size + 6 : m.hand_size + 3));
            float lp[3], lv[3];
            kg_ray_local(ws.gpos[g], ws.gmat[g], site + 6 * i, site + 6 * i + 3, lp, lv);
            if (kg_ray_box(size, lp, lv) < 0.f) continue;
            hits[KG_ATOMIC_INC(&ws.s_iter)] = pr;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        const int n = ws.s_iter;
        for (int k = l; k < n; k += 32) {
            const int pr = hits[k], i = pr >> 3, g = 1 + (pr & 7);
            float lp[3], lv[3];
            kg_ray_local(ws.gpos[g], ws.gmat[g], site + 6 * i, site + 6 * i + 3, lp, lv);
            const float x = kg_ray_mesh_bvh(m, kg_mesh_of_geom(ws, g), lp, lv);
            if (x >= 0.f) KG_ATOMIC_MIN_U(&rmin[i], KG_F2U(fabsf(x)));
        }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l < KG_NSITE) {
            const unsigned u = rmin[l];
            float r = -1.f;
            if (u != KG_RAY_NONE) {
#if defined(__CUDA_ARCH__)
                r = __uint_as_float(u);
#else
                memcpy(&r, &u, 4);
#endif
            }
            KG_RANGE(ws)[l] = r;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ O1-O9, R1: obs + reward
// Restates _get_obs('local') / _get_reward (ENV:438-534, 631-687) on the kinematics of the LAST forward
// pass (the one-substep staleness of SURVEY 8a O-note is reproduced: this runs before nothing else
// touches ws.gpos / ws.sens_q).
// (u_obj . u_ctr)^20: the 20th power multiplies the relative rounding error of the cosine by 20, which in fp32 (four 2-ulp
// divisions and two square roots) lands at 1-2e-5 -- above the 1e-5 obs tolerance.  Seven of these per RL step and env, so
// the few dozen operations are done in fp64 (inputs are the fp32 kinematics; the result is rounded once).
KG_DEV float kg_dotprod20(const float *hand, const float *p) {
    const double ox = fabs((double)p[0] - (double)hand[0]), oy = fabs((double)p[1] - (double)hand[1]);
    const double cx = fabs((double)hand[0]), cy = fabs((double)hand[1]);
    const double d = (ox * cx + oy * cy) / sqrt((ox * ox + oy * oy) * (cx * cx + cy * cy));
    const double d2 = d * d, d4 = d2 * d2, d8 = d4 * d4, d16 = d8 * d8;
    return (float)(d16 * d4);
}
KG_DEV float kg_tri_area(const float *a, const float *b) { float c[3]; kg_cross(c, a, b); return 0.5f * sqrtf(kg_dot(c, c)); }

template <class WS> KG_DEV void kg_obs(const KgModel &m, WS &ws, int kg_lane, float *reward, int *done) {
    // hand frame (O1): T = (R_palm C)^T, wrist = p_palm + T^T [-0.009, 0.048, 0]
    KG_PHASE(l) {
        const float *Rp = ws.gmat[1];
        float T[9];   // rows of T = columns of Rp*C ; Rp*C columns: (-Rp col1, -Rp col2, Rp col0)
        for (int c = 0; c < 3; c++) { T[c] = -Rp[3 * c + 1]; T[3 + c] = -Rp[3 * c + 2]; T[6 + c] = Rp[3 * c]; }
        float off[3] = {-0.009f, 0.048f, 0.0f}, w[3], wrist[3], tt[3];
        kg_multmv(w, T, off); kg_add(wrist, ws.gpos[1], w);
        kg_mulmv(tt, T, wrist);
        const float *obj = ws.gpos[8];
        float *o = KG_WSOBS(ws);
        const int fgeom[6] = {2, 4, 6, 3, 5, 7};
        if (l < 6) {
            float r[3];
            kg_mulmv(r, T, ws.gpos[fgeom[l]]);
            for (int c = 0; c < 3; c++) o[3 * l + c] = r[c] - tt[c];
            o[75 + l] = kg_dotprod20(ws.hp, ws.gpos[fgeom[l]]);
        } else if (l == 6) {
            float r[3];
            kg_mulmv(r, T, wrist);
            for (int c = 0; c < 3; c++) o[18 + c] = r[c] - tt[c];
            kg_mulmv(r, T, obj);
            for (int c = 0; c < 3; c++) { r[c] -= tt[c]; o[21 + c] = r[c]; }
            // _get_angles (ENV:569-582) computes acos(y / sqrt(x^2 + y^2)) in fp64; acos is ill-conditioned at +-1 (an object
            // straight ahead), so fp32 evaluates the same angle as atan2(|x|, y), which is accurate to an ulp everywhere
            o[49] = atan2f(fabsf(r[0]), r[1]);   // z_angle
            o[48] = atan2f(fabsf(r[2]), r[1]);   // x_angle
            o[81] = kg_dotprod20(ws.hp, obj);
            const float gw[3] = {0, 0, -1};
            kg_mulmv(r, T, gw);
            for (int c = 0; c < 3; c++) o[67 + c] = r[c];
            const KgShape &sh = m.shape[ws.shape];
            o[33] = sh.obs_size[0]; o[34] = sh.obs_size[1]; o[35] = sh.obs_size[2] * 2.f;
        } else if (l == 7) {
            for (int i = 0; i < 9; i++) o[24 + i] = (i < 2 ? -ws.sens_q[i] : ws.sens_q[i]);
        } else if (l >= 8 && l < 20) {
            float p[3], z[3], df[3];
            kg_site_pose(m, ws, m.fod_site[l - 8], p, z);
            kg_sub(df, p, obj);
            o[36 + l - 8] = sqrtf(kg_dot(df, df));
        }
        if (l < KG_NSITE) { float r = KG_RANGE(ws)[l]; o[50 + l] = (r == -1.f) ? 6.f : r; }
    }
    KG_SYNC();
    // experimental_sensor (O8), reward (R1)
    KG_PHASE(l) {
        if (l == 0) {
            float *o = KG_WSOBS(ws);
            const float *fl = o;
            float s1[3], s2[3];
            kg_sub(s1, fl, fl + 6); kg_sub(s2, fl, fl + 3);
            float front_area = kg_tri_area(s1, s2);
            float top1 = kg_tri_area(fl, fl + 9), top2 = kg_tri_area(fl + 9, fl + 12), top3 = kg_tri_area(fl + 3, fl + 12);
            float top4 = kg_tri_area(fl + 6, fl + 15), top5 = kg_tri_area(fl + 9, fl + 15);
            float top_area = fmaxf(top1 + top2 + top3, top1 + top4 + top5);
            const float *Rp = ws.gmat[1];
            float T[9];
            for (int c = 0; c < 3; c++) { T[c] = -Rp[3 * c + 1]; T[3 + c] = -Rp[3 * c + 2]; T[6 + c] = Rp[3 * c]; }
            float off[3] = {-0.009f, 0.048f, 0.0f}, w[3], wrist[3], tt[3];
            kg_multmv(w, T, off); kg_add(wrist, ws.gpos[1], w);
            kg_mulmv(tt, T, wrist);
            float sp[3] = {0, 0, 0}; int ns = 0;
            for (int i = 0; i < 5; i++) {
                float rng = o[50 + i];
                if (rng < 0.06f) {
                    float p[3], z[3], r[3];
                    kg_site_pose(m, ws, i, p, z);
                    kg_mulmv(r, T, p);
                    sp[0] += r[0] - tt[0]; sp[1] += r[1] - tt[1] + rng; sp[2] += r[2] - tt[2];
                    ns++;
                }
            }
            if (ns == 0) { sp[0] = sp[1] = sp[2] = 0.2f; } else { float inv = 1.f / ns; sp[0] *= inv; sp[1] *= inv; sp[2] *= inv; }
            o[70] = sp[0]; o[71] = sp[1]; o[72] = sp[2];
            const float *sz = m.shape[ws.shape].obs_size;
            float g0 = fabsf(o[67]), g1 = fabsf(o[68]), g2 = fabsf(o[69]);
            int gi = 0; if (g1 > g0) gi = 1; if (g2 > (gi == 0 ? g0 : g1)) gi = 2;
            float fp, tp;
            if (gi == 2) { fp = fabsf(sz[0] * sz[2]) / front_area; tp = fabsf(sz[0] * sz[1]) / top_area; }
            else if (gi == 1) { fp = fabsf(sz[0] * sz[2]) / front_area; tp = fabsf(sz[1] * sz[2]) / top_area; }
            else { fp = fabsf(sz[0] * sz[1]) / front_area; tp = fabsf(sz[0] * sz[2]) / top_area; }
            o[73] = fp; o[74] = tp;
            float z = ws.gpos[8][2];
            int dn = (fabsf(z - 0.2f) < 0.005f) || (z >= 0.2f);
            *reward = dn ? 50.f : 0.f;
            *done = dn;
        }
    }
    KG_SYNC();
}

// _get_obs(state_rep="global") (ENV:484-534): the 74-value layout = global finger / wrist / object positions (24) followed
// by the entries the two layouts share.  Runs right after kg_obs (the local vector is still staged in the workspace).
#define KG_NOBS_GLOBAL 74
template <class WS> KG_DEV void kg_obs_global_store(const WS &ws, float *out, int N, int env, int kg_lane) {
    KG_PHASE(l) {
        const float *o = KG_WSOBS(ws);
        const float *Rp = ws.gmat[1];
        // wrist = p_palm + T^T [-0.009, 0.048, 0] with T^T = Rp C, columns (-Rp col1, -Rp col2, Rp col0)
        float wrist[3];
        for (int c = 0; c < 3; c++) wrist[c] = ws.gpos[1][c] + (-Rp[3 * c + 1]) * -0.009f + (-Rp[3 * c + 2]) * 0.048f;
        const int fgeom[6] = {2, 4, 6, 3, 5, 7};
        for (int i = l; i < KG_NOBS_GLOBAL; i += 32) {
            float v;
            if (i < 18) v = ws.gpos[fgeom[i / 3]][i % 3];
            else if (i < 21) v = wrist[i - 18];
            else if (i < 24) v = ws.gpos[8][i - 21];
            else if (i < 67) v = o[i];            // joints, size, finger-object distances, angles, rangefinders
            else v = o[i + 8];                    // finger dot products, object dot product
            out[i * N + env] = v;
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ S1: ctrl from the action (ENV:1501-1534)
// a6 = [slide_x, slide_y, slide_z (wrist), f1, f2, f3]; a 4-d action [wrist, f1, f2, f3] is padded to [0, 0, wrist, f1, f2, f3]
// by the caller exactly as ENV:1502-1503 does.  slide_vector = Tfw[0:3,0:3] a6[0:3], x and y negated (ENV:1518-1522);
// gravity compensation = Tfw[0:3,0:3] (0, 0, 0.733*10/25), x and y negated (ENV:1511-1515).
template <class WS> KG_DEV void kg_ctrl_from_action(WS &ws, const float *a6, int kg_lane) {
    KG_PHASE(l) {
        if (l == 0) {
            const float *Rp = ws.gmat[1];
            // rows of T = Tfw[0:3,0:3] are (-Rp col1, -Rp col2, Rp col0)
            const float T[9] = {-Rp[1], -Rp[4], -Rp[7], -Rp[2], -Rp[5], -Rp[8], Rp[0], Rp[3], Rp[6]};
            const float gcomp = 0.733f * 10.f / 25.f;
            float sv[3];
            for (int r = 0; r < 3; r++) sv[r] = T[3 * r] * a6[0] + T[3 * r + 1] * a6[1] + T[3 * r + 2] * a6[2];
            ws.ctrl[0] = -sv[0]; ws.ctrl[2] = -sv[1]; ws.ctrl[4] = sv[2];
            ws.ctrl[1] = -T[2] * gcomp; ws.ctrl[3] = -T[5] * gcomp; ws.ctrl[5] = T[8] * gcomp;
            ws.ctrl[6] = a6[3]; ws.ctrl[7] = a6[4]; ws.ctrl[8] = a6[5];
        }
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ C1-C3: expert controller (EXP:318-352, 367-671, 745-804)
struct KgExpertState { float init_obj_x, init_dot, prev_fx[3]; int have_prev, num_good, total_steps; };
KG_DEV float kg_pid_touch(float a, float b) { float e = a - b; return e + e / 15.f; }
KG_DEV float kg_pid_velocity(float dot) { float e = 1.f - dot, v = e + e / 15.f, a = (v / 1.25f) * 0.3f; return a < 0.05f ? 0.05f : a; }
KG_DEV void kg_pd_controller(const KgExpertState &e, const float *obs, int lift, float *a) {
    const float cv = 0.5f, minv = 0.5f, flv = 0.5f;
    float dot = obs[81], f1 = 0, f2 = 0, f3 = 0;
    if (fabsf(e.init_obj_x) <= 0.03f) {
        f1 = f2 = f3 = cv;
        if (fabsf(dot - e.init_dot) > 0.01f) { f1 = cv; f2 = f3 = cv / 2; }
        if (lift) { f1 = flv / 2; f2 = f3 = flv; }
    } else if (e.init_obj_x > 0.0f) {
        if (fabsf(dot - e.init_dot) < 0.01f) { f1 = 0; f2 = f3 = kg_pid_touch(dot, obs[79]); }
        else {
            if (fabsf(1.f - dot) > 0.01f) { f1 = minv; f2 = f3 = kg_pid_velocity(dot); }
            else { f1 = kg_pid_touch(dot, obs[78]); f2 = f3 = 0; }
            if (lift) { f1 = flv / 2; f2 = f3 = flv; }
        }
    } else {
        if (fabsf(dot - e.init_dot) < 0.01f) { f1 = kg_pid_touch(dot, obs[78]); f2 = f3 = 0; }
        else {
            if (fabsf(1.f - dot) > 0.01f) { f1 = kg_pid_velocity(dot); f2 = f3 = minv; }
            else { f2 = f3 = kg_pid_touch(dot, obs[79]); f1 = 0; }
            if (lift) { f1 = flv / 2; f2 = f3 = flv; }
        }
    }
    a[1] = f1; a[2] = f2; a[3] = f3;
    for (int i = 1; i < 4; i++) { if (a[i] < 0.5f) a[i] = 0.5f; else if (a[i] > 0.8f) a[i] = 0.8f; }   // check_vel_in_range
}
// mode 0 naive, 1 position-dependent, 2 combined.  Advances the episode bookkeeping of EXP:745-804.
KG_DEV void kg_expert_action(KgExpertState &e, const float *obs, int mode, float *action) {
    int grasp = 0;
    if (e.have_prev) {
        float ch = fabsf(e.prev_fx[0] - obs[9]) / 15.f + fabsf(e.prev_fx[1] - obs[12]) / 15.f + fabsf(e.prev_fx[2] - obs[15]) / 15.f;
        grasp = ch < 0.0002f;
    }
    if (grasp) e.num_good++;
    int lift = (e.total_steps > 10) && (e.num_good >= 1);
    float x = obs[21];
    action[1] = action[2] = action[3] = 0.5f;
    if (mode == 1) kg_pd_controller(e, obs, lift, action);
    else if (mode == 2) {
        if (x < -0.04f || x > 0.04f) kg_pd_controller(e, obs, lift, action);
        else if ((-0.04f <= x && x <= -0.02f) || (0.02f <= x && x <= 0.04f)) {
            float ex[4];
            kg_pd_controller(e, obs, lift, ex);
            action[1] = action[2] = action[3] = ex[2];   // np.interp quirk at EXP:654
        }
    }
    action[0] = lift ? 0.6f : 0.0f;
    if (e.total_steps > 0) { e.prev_fx[0] = obs[9]; e.prev_fx[1] = obs[12]; e.prev_fx[2] = obs[15]; e.have_prev = 1; }
    e.total_steps++;
}

// ------------------------------------------------------------------------------------------------ env-level drivers
// Struct-of-arrays device state: field-major [field][N] so that consecutive envs are consecutive floats.
struct KgEnvArrays {
    float *qpos, *qvel, *qwarm;   // [16][N], [15][N], [15][N]
    float *hand_quat;             // [4][N]
    int *shape, *elapsed;         // [N]
    float *ex_f;                  // [5][N] init_obj_x, init_dot, prev_fx[3]
    int *ex_i;                    // [3][N] have_prev, num_good, total_steps
    float *obs;                   // [82][N] last observation (input of the expert controller)
    float *last_action;           // [4][N] action applied by the last step (given or expert)
    unsigned long long *stats;    // [8] mpr pairs, newton iterations, contact overflows, substeps, contacts (64-bit: they run for a whole training)
    int *load;                    // [N] cost of each env's last RL step (Newton iterations + 2 x MPR pairs): key of the launch schedule
    long long *prof;              // [KG_NPROF] KG_PROF builds only
    float *obs_global;            // [74][N] or NULL: _get_obs(state_rep="global") of the same step (kg_set_obs_global)
    float *ranges;                // [KG_NSITE][N] rangefinder readings of the current RL step, written by the batch-wide ray kernel (GPU path)
    unsigned char *ccache;        // [N][KG_CCACHE_BYTES] collision caches (warm-start vertices, separating axes) kept across RL steps
    int *orient;                  // [N] hand orientation of the current episode as the sampler drew it: 0 normal, 1 rotated, 2 top (env.get_orientation(), ENV:1180-1222); -1 after a reset with explicit poses
    int *grasp_given;             // [N] Grasp_Reward: the grasp-classifier reward was paid in this episode (ENV:103, 655-657; cleared by reset ENV:1372)
    int N, max_steps;
    int act_dim;                  // 4 = [wrist, f1, f2, f3], 6 = [slide_x, slide_y, slide_z, f1, f2, f3] (ENV:1502-1503)
};

template <class WS> KG_DEV void kg_load_env(const KgModel &m, WS &ws, const KgEnvArrays &A, int env, int kg_lane) {
    KG_PHASE(l) {
        const int N = A.N;
        if (l < 16) ws.qpos[l] = A.qpos[l * N + env];
        if (l < 15) { ws.qvel[l] = A.qvel[l * N + env]; ws.qwarm[l] = A.qwarm[l * N + env]; }
        if (l == 31) { ws.stat_mpr = 0; ws.stat_newton = 0; ws.stat_overflow = 0; ws.stat_pad = 0; ws.ntask = 0; ws.qvel[15] = 0.f; ws.qwarm[15] = 0.f; }
#ifdef KG_PROF
        if (l < KG_NPROF) ws.prof[l] = 0;
#if defined(__CUDA_ARCH__)
        if (l == 0) ws.prof_t0 = clock64();
#endif
#endif
    }
    KG_SYNC();
    float hq[4] = {A.hand_quat[env], A.hand_quat[A.N + env], A.hand_quat[2 * A.N + env], A.hand_quat[3 * A.N + env]};
    kg_episode_constants(m, ws, hq, A.shape[env], kg_lane);
}
template <class WS> KG_DEV void kg_store_env(WS &ws, const KgEnvArrays &A, int env, int kg_lane) {
    KG_PHASE(l) {
        const int N = A.N;
        if (l < 16) A.qpos[l * N + env] = ws.qpos[l];
        if (l < 15) { A.qvel[l * N + env] = ws.qvel[l]; A.qwarm[l * N + env] = ws.qwarm[l]; }
    }
    KG_SYNC();
}

// ---- carry records: the part of KgWS that travels through global memory between the kernels of one RL step
KG_DEV void kg_copy16(void *dst, const void *src, int bytes, int kg_lane) {
#if defined(__CUDA_ARCH__)
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (int i = kg_lane; i < bytes / 16; i += 32) d4[i] = s4[i];
#else
    (void)kg_lane;
    memcpy(dst, src, (size_t)bytes);
#endif
}
template <class WS> KG_DEV void kg_carry_store(const WS &ws, unsigned char *carry, int kg_lane) {
    KG_SYNC();
    kg_copy16(carry, &ws, (int)KG_CARRY_A_BYTES, kg_lane);
    kg_copy16(carry + KG_CARRY_A_BYTES, ws.cv1, (int)KG_CCACHE_BYTES, kg_lane);      // (its offset differs between the light and the full workspace)
    kg_copy16(carry + KG_CARRY_A_BYTES + KG_CCACHE_BYTES, reinterpret_cast<const unsigned char *>(&ws) + KG_CARRY_B_OFF, (int)KG_CARRY_B_BYTES, kg_lane);
    KG_SYNC();
}
template <class WS> KG_DEV void kg_carry_load(WS &ws, const unsigned char *carry, int kg_lane) {
#if defined(__CUDA_ARCH__)
    // global -> shared without a register round trip: every 16-byte chunk of both sections is in flight at once (LDGSTS),
    // one wait for the whole record instead of one L2 round trip per loop iteration
    {
        const unsigned dstA = (unsigned)__cvta_generic_to_shared(&ws);
        const unsigned dstB = dstA + (unsigned)KG_CARRY_B_OFF;
        for (int i = kg_lane; i < (int)(KG_CARRY_A_BYTES / 16); i += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstA + 16u * i), "l"(carry + 16 * i) : "memory");
        for (int i = kg_lane; i < (int)(KG_CARRY_B_BYTES / 16); i += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstB + 16u * i), "l"(carry + KG_CARRY_A_BYTES + KG_CCACHE_BYTES + 16 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
#else
    kg_copy16(&ws, carry, (int)KG_CARRY_A_BYTES, kg_lane);
    kg_copy16(reinterpret_cast<unsigned char *>(&ws) + KG_CARRY_B_OFF, carry + KG_CARRY_A_BYTES + KG_CCACHE_BYTES, (int)KG_CARRY_B_BYTES, kg_lane);
#endif
    KG_SYNC();
#ifdef KG_PROF
    KG_PHASE(l) {
        if (l < KG_NPROF) ws.prof[l] = 0;
#if defined(__CUDA_ARCH__)
        if (l == 0) ws.prof_t0 = clock64();
#endif
    }
    KG_SYNC();
#endif
}
// the collision cache of an env (640 bytes, section C of its record or its slot of A.ccache) into the workspace
template <class WS> KG_DEV void kg_ccache_load(WS &ws, const unsigned char *src, int kg_lane) {
    KG_SYNC();
    kg_copy16(ws.cv1, src, (int)KG_CCACHE_BYTES, kg_lane);
    KG_SYNC();
}
template <class WS> KG_DEV void kg_prof_flush(WS &ws, const KgEnvArrays &A, int kg_lane) {
#if defined(KG_PROF) && defined(__CUDA_ARCH__)
    __syncwarp();
    if (A.prof && kg_lane < KG_NPROF) atomicAdd((unsigned long long *)A.prof + kg_lane, (unsigned long long)ws.prof[kg_lane]);
#else
    (void)ws; (void)A; (void)kg_lane;
#endif
}

// The narrowphase tasks of this env (pair lanes in `mask`) go to the global task list: (env << 8) | pair lane.
// q[0] = tasks queued.  One atomic per warp.
KG_DEV void kg_emit_tasks(unsigned mask, int env, int *q, unsigned *tasks, int kg_lane) {
    if (mask == 0) return;
#if defined(__CUDA_ARCH__)
    int base = 0;
    if (kg_lane == 0) base = atomicAdd(q, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((mask >> kg_lane) & 1) tasks[base + __popc(mask & ((1u << kg_lane) - 1))] = ((unsigned)env << 8) | (unsigned)kg_lane;
#else
    (void)kg_lane;
    for (int k = 0; k < 32; k++) if ((mask >> k) & 1) tasks[q[0]++] = ((unsigned)env << 8) | (unsigned)k;
#endif
}

// First kernel of an RL step for one env: [expert action] -> ctrl -> kinematics + lane-parallel collision of substep 0.
// action_in == nullptr selects the on-device expert (mode 0 naive, 1 position-dependent, 2 combined).
// Returns the narrowphase task mask; the caller stores the carry record and emits the tasks.
template <class WS> KG_DEV unsigned kg_step_begin(const KgModel &m, WS &ws, const KgEnvArrays &A, int env, const float *action_in, int expert_mode, int kg_lane) {
    kg_load_env(m, ws, A, env, kg_lane);
    float act[4] = {0, 0, 0, 0}, sxy[2] = {0, 0};      // act = [wrist (slide_z), f1, f2, f3]; sxy = slide_x, slide_y of a 6-d action
    if (action_in) {
        if (A.act_dim == 6) {
            sxy[0] = action_in[env]; sxy[1] = action_in[A.N + env];
            for (int c = 0; c < 4; c++) act[c] = action_in[(c + 2) * A.N + env];
        } else for (int c = 0; c < 4; c++) act[c] = action_in[c * A.N + env];
    } else {
        KG_PHASE(l) { for (int i = l; i < KG_NOBS; i += 32) KG_WSOBS(ws)[i] = A.obs[i * A.N + env]; }
        KG_SYNC();
        KG_PHASE(l) {
            if (l == 0) {
                const int N = A.N;
                KgExpertState e;
                e.init_obj_x = A.ex_f[env]; e.init_dot = A.ex_f[N + env];
                e.prev_fx[0] = A.ex_f[2 * N + env]; e.prev_fx[1] = A.ex_f[3 * N + env]; e.prev_fx[2] = A.ex_f[4 * N + env];
                e.have_prev = A.ex_i[env]; e.num_good = A.ex_i[N + env]; e.total_steps = A.ex_i[2 * N + env];
                float a4[4];
                kg_expert_action(e, KG_WSOBS(ws), expert_mode, a4);
                A.ex_f[2 * N + env] = e.prev_fx[0]; A.ex_f[3 * N + env] = e.prev_fx[1]; A.ex_f[4 * N + env] = e.prev_fx[2];
                A.ex_i[env] = e.have_prev; A.ex_i[N + env] = e.num_good; A.ex_i[2 * N + env] = e.total_steps;
                for (int c = 0; c < 4; c++) ws.red[c] = a4[c];
            }
        }
        KG_SYNC();
        for (int c = 0; c < 4; c++) act[c] = ws.red[c];
        KG_SYNC();
    }
    KG_PHASE(l) { if (l < 4) A.last_action[l * A.N + env] = act[l]; }
    KG_SYNC();
    {
        const float a6[6] = {sxy[0], sxy[1], act[0], act[1], act[2], act[3]};
        kg_ctrl_from_action(ws, a6, kg_lane);
    }
    KG_PROBE(19);
    kg_kinematics(m, ws, kg_lane);
    KG_PROBE(0);
    kg_ccache_load(ws, A.ccache + (size_t)env * KG_CCACHE_BYTES, kg_lane);      // (J, which it aliases, held the expert's obs staging until here)
    unsigned mask = kg_collision_lane(m, ws, kg_lane);
    KG_PROBE(2);
    return mask;
}

// Second half of one mj_step (the narrowphase kernel has filed the hull contacts of this substep into the carry):
// dynamics, constraint rows, Newton solve, Euler.  The workspace holds the carry of this env.
template <class WS> KG_DEV void kg_substep_finish(const KgModel &m, WS &ws, int kg_lane) {
    KG_STAGE(0);
    KG_PROBE(18);
    kg_dynamics(m, ws, kg_lane);
    KG_PROBE(0);
    kg_compaction(m, ws, kg_lane);
    KG_PROBE(3);
    KG_STAGE(2);
    KG_PROBE(4);
    kg_make_rows(m, ws, kg_lane);
    KG_PROBE(5);
    KG_STAGE(3);
    KG_PROBE(6);
    kg_solve(m, ws, kg_lane);
    KG_PROBE(13);
    KG_STAGE(4);
    KG_PROBE(14);
    kg_euler(m, ws, kg_lane);
    KG_PROBE(15);
}
// First half of the next mj_step: kinematics of the new state + lane-parallel collision; returns the task mask.
template <class WS> KG_DEV unsigned kg_substep_start(const KgModel &m, WS &ws, const unsigned char *rec, int kg_lane) {
    kg_kinematics(m, ws, kg_lane);
    KG_PROBE(0);
    KG_STAGE(1);
    KG_PROBE(1);
    kg_ccache_load(ws, rec + KG_CARRY_A_BYTES, kg_lane);      // section C of the env's record, into the storage J no longer needs
    unsigned mask = kg_collision_lane(m, ws, kg_lane);
    KG_PROBE(2);
    return mask;
}

// Last kernel of an RL step for one env (after the last substep's finish): rays -> obs/reward/done (+TimeLimit),
// state and outputs back to the struct-of-arrays.
template <class WS> KG_DEV void kg_step_end(const KgModel &m, WS &ws, const KgEnvArrays &A, int env, int nsub, const unsigned char *rec, const float *ranges,
                        float *obs_out, float *reward_out, int *done_out, float *action_out, int kg_lane) {
    // the collision cache goes back to its per-env home for the next RL step (record -> array; it is not in the workspace)
    kg_copy16(A.ccache + (size_t)env * KG_CCACHE_BYTES, rec + KG_CARRY_A_BYTES, (int)KG_CCACHE_BYTES, kg_lane);
    KG_PROBE(19);
    if (ranges) {      // the batch-wide ray kernel (kg_api.cu) already traced this step's 17 rays from the same kinematics
        KG_PHASE(l) { if (l < KG_NSITE) KG_RANGE(ws)[l] = ranges[l * A.N + env]; }
        KG_SYNC();
    } else kg_rays(m, ws, kg_lane);
    KG_PROBE(16);
    float rew = 0.f; int dn = 0;
    kg_obs(m, ws, kg_lane, &rew, &dn);
    if (A.obs_global) kg_obs_global_store(ws, A.obs_global, A.N, env, kg_lane);
    KG_PROBE(17);
    kg_store_env(ws, A, env, kg_lane);
    KG_PHASE(l) {
        const int N = A.N;
        for (int i = l; i < KG_NOBS; i += 32) { float v = KG_WSOBS(ws)[i]; A.obs[i * N + env] = v; if (obs_out) obs_out[i * N + env] = v; }
        if (l == 0) {
            // reward/done were produced by lane 0 of kg_obs; TimeLimit (gym_kinova_gripper/__init__.py:3-7, MAIN:164)
            int el = A.elapsed[env] + 1;
            A.elapsed[env] = el;
            int d = dn || (el >= A.max_steps);
            if (reward_out) reward_out[env] = rew;
            if (done_out) done_out[env] = d;
            if (A.load) A.load[env] = d ? 0 : ws.stat_newton + 2 * ws.stat_mpr;   // a finished episode restarts contact-free
            if (action_out) for (int c = 0; c < 4; c++) action_out[c * N + env] = A.last_action[c * N + env];
#if defined(__CUDA_ARCH__)
            if (A.stats) {
                atomicAdd(A.stats, (unsigned long long)ws.stat_mpr); atomicAdd(A.stats + 1, (unsigned long long)ws.stat_newton);
                atomicAdd(A.stats + 2, (unsigned long long)ws.stat_overflow); atomicAdd(A.stats + 3, (unsigned long long)nsub);
                atomicAdd(A.stats + 4, (unsigned long long)ws.stat_pad);
            }
#else
            if (A.stats) { A.stats[0] += (unsigned long long)ws.stat_mpr; A.stats[1] += (unsigned long long)ws.stat_newton; A.stats[2] += (unsigned long long)ws.stat_overflow; A.stats[3] += (unsigned long long)nsub; A.stats[4] += (unsigned long long)ws.stat_pad; }
#endif
        }
    }
    KG_SYNC();
}

// reset of one env (ENV:1310-1410 without the file system): hand euler -> quaternion, hand joints 0, object at
// obj_pos with identity orientation, velocities and warm start 0, forward kinematics + rays + first obs,
// centroid fix-up (ENV:1382-1389), expert-controller episode state (EXP:367-379, 724-727).
// handp = the slide joint positions determine_hand_location (ENV:1286-1307) puts into qpos[0:3] (0 for 'normal').
template <class WS> KG_DEV void kg_env_reset(const KgModel &m, WS &ws, const KgEnvArrays &A, int env, int shape, const float *euler, const float *objp,
                         const float *handp, float *obs_out, int kg_lane) {
    float hq[4] = {1, 0, 0, 0};
    for (int i = 0; i < 3; i++) {   // MuJoCo eulerseq "xyz": q = qx qy qz
        float s, c;
        sincosf(0.5f * euler[i], &s, &c);
        float r[4] = {c, 0, 0, 0};
        r[i + 1] = s;
        float t[4] = {hq[0] * r[0] - hq[1] * r[1] - hq[2] * r[2] - hq[3] * r[3],
                      hq[0] * r[1] + hq[1] * r[0] + hq[2] * r[3] - hq[3] * r[2],
                      hq[0] * r[2] - hq[1] * r[3] + hq[2] * r[0] + hq[3] * r[1],
                      hq[0] * r[3] + hq[1] * r[2] - hq[2] * r[1] + hq[3] * r[0]};
        for (int c2 = 0; c2 < 4; c2++) hq[c2] = t[c2];
    }
    {
        float n = 1.0f / sqrtf(hq[0] * hq[0] + hq[1] * hq[1] + hq[2] * hq[2] + hq[3] * hq[3]);
        for (int c = 0; c < 4; c++) hq[c] *= n;
    }
    KG_PHASE(l) {
        if (l < 16) ws.qpos[l] = 0.f;
        if (l < 16) { ws.qvel[l] = 0.f; ws.qwarm[l] = 0.f; }
    }
    KG_SYNC();
    KG_PHASE(l) {
        if (l == 0) {
            ws.qpos[0] = handp[0]; ws.qpos[1] = handp[1]; ws.qpos[2] = handp[2];
            ws.qpos[9] = objp[0]; ws.qpos[10] = objp[1]; ws.qpos[11] = objp[2]; ws.qpos[12] = 1.f;
        }
    }
    KG_SYNC();
    kg_episode_constants(m, ws, hq, shape, kg_lane);
    kg_kinematics(m, ws, kg_lane);
    {
        float d[3] = {objp[0] - ws.gpos[8][0], objp[1] - ws.gpos[8][1], objp[2] - ws.gpos[8][2]};
        if (kg_dot(d, d) > 0.05f * 0.05f) {
            KG_SYNC();
            KG_PHASE(l) { if (l == 0) { ws.qpos[9] = objp[0] + d[0]; ws.qpos[10] = objp[1] + d[1]; ws.qpos[11] = objp[2] + d[2]; } }
            KG_SYNC();
            kg_kinematics(m, ws, kg_lane);
        }
    }
    kg_rays(m, ws, kg_lane);
    float rew; int dn;
    kg_obs(m, ws, kg_lane, &rew, &dn);
    if (A.obs_global) kg_obs_global_store(ws, A.obs_global, A.N, env, kg_lane);
    kg_store_env(ws, A, env, kg_lane);
    KG_PHASE(l) {      // a new episode starts with an empty collision cache: vertex 0 as warm start, no cached axis
        unsigned *cc = reinterpret_cast<unsigned *>(A.ccache + (size_t)env * KG_CCACHE_BYTES);
        for (int i = l; i < (int)(KG_CCACHE_BYTES / 4); i += 32) cc[i] = 0u;
    }
    KG_SYNC();
    KG_PHASE(l) {
        const int N = A.N;
        for (int i = l; i < KG_NOBS; i += 32) { float v = KG_WSOBS(ws)[i]; A.obs[i * N + env] = v; if (obs_out) obs_out[i * N + env] = v; }
        if (l < 4) A.hand_quat[l * N + env] = hq[l];
        if (l == 0) {
            A.shape[env] = shape; A.elapsed[env] = 0; A.grasp_given[env] = 0;
            A.ex_f[env] = KG_WSOBS(ws)[21]; A.ex_f[N + env] = KG_WSOBS(ws)[81];
            A.ex_f[2 * N + env] = 0.f; A.ex_f[3 * N + env] = 0.f; A.ex_f[4 * N + env] = 0.f;
            A.ex_i[env] = 0; A.ex_i[N + env] = 0; A.ex_i[2 * N + env] = 0;
        }
    }
    KG_SYNC();
}
