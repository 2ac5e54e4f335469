// kg_api.cu -- kernels + the C ABI of include/kg.h.
//
// Built two ways from this one source:
//   nvcc -gencode arch=compute_100a,code=sm_100a ...        -> kinovagrasping_b200/libkg_b200.so  (THE PRODUCT)
//   g++ -x c++ -DKG_EMU ...                                 -> tests/emu/libkg_emu.so  (TEST-ONLY host emulation
//        of the warp phases, used to debug the algorithm against the oracle without a GPU; the package
//        never loads it and has no CPU fallback).
#include "../../include/kg.h"
#if !defined(KG_EMU) && !defined(KG_MPR_SMEM_HULLS) && !defined(KG_MPR_L1)
#define KG_MPR_L1      // narrowphase kernel reads the hull tables through the read-only L1 path (no shared-memory staging: +2 % measured)
#endif
#include "kg_core.cuh"
#include "kg_forest.cuh"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#define KG_MAXGROUPS 32
#ifndef KG_EMU
#include <cuda_runtime.h>
#ifndef KG_WARPS
#define KG_WARPS 16         // environments (= warps) per CTA; one CTA per SM: 16 workspaces + the finger hulls fill its shared memory
#endif
#ifndef KG_MINBLOCKS
#define KG_MINBLOCKS 1      // CTAs per SM the register allocation must leave room for (512 threads x 128 registers = the whole file)
#endif
#ifndef KG_LWARPS
#ifdef KG_PROF
#define KG_LWARPS 7         // (the probe arrays of a KG_PROF build enlarge every workspace by 264 bytes)
#else
#define KG_LWARPS 8         // light workspace: warps per CTA ...
#endif
#endif
#ifndef KG_LMINBLOCKS
#define KG_LMINBLOCKS 3     // ... and CTAs per SM: 24 resident warps x 80 registers.  (7 x 4 = 28 warps x 72 registers makes 4096 envs a
                            // single wave and was measured: no faster -- profiles/README.md r2: a launch lasts as long as its slowest
                            // warp, and at that occupancy the SM is already at 62 % of its issue slots)
#endif
#define KG_SPLIT (KG_MAXCON_LIGHT < KG_MAXCON)
#define KG_LSMEM (KG_LWARPS * sizeof(KgWSLight))
#if KG_MPR_TEAMS > 1 && !defined(KG_MPR_L1)
#error "narrowphase teams need the hull tables on the read-only L1 path (KG_MPR_L1)"
#endif
__constant__ KgModel c_model;
// Programmatic dependent launch (kg_set_dependent_launch): the kernels of a group's chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization.  Every chain kernel first waits for its predecessor grid to have
// completed and flushed (griddepcontrol.wait -- a no-op when the launch carried no such edge) and then lets ITS dependent be
// launched right away (griddepcontrol.launch_dependents): the next kernel's launch and CTA dispatch happen while this one
// runs, its CTAs sit at their own wait until this grid is done.  Execution order and results are unchanged.
#ifdef KG_PDL_TRIGGER_EARLY
#define KG_PDL_PROLOGUE() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
#else      // no explicit trigger: the dependent is launched when every CTA of this grid has exited
#define KG_PDL_PROLOGUE() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif
#endif

struct KgGraphKey { int cfg, full, qs0, mode; const void *action; void *obs, *reward, *done, *act_out; };
struct kg_handle_s {
    int n = 0, device = 0, frame_skip = 15, env_offset = 0;
    void *stream = nullptr;
    std::string err;
    KgModel model;            // host copy; table pointers are device pointers (or host pointers under KG_EMU)
    KgEnvArrays A;
    int *ep_count = nullptr;  // [n] episodes started per env (RNG counter)
    int *perm = nullptr;      // [n] launch schedule of the per-env kernels: per env group, warp slot -> env (group-relative), heaviest envs first
    int schedule = 1;         // 0 = slot i runs env i; 1 = slots sorted by the load of each env's last step
    unsigned char *carry = nullptr;   // [n][KG_CARRY_BYTES] per-env carry records between the kernels of one RL step
    unsigned *tasks = nullptr;        // [n * 22] narrowphase task list of the current substep: (env << 8) | pair lane
    int *q = nullptr;                 // two {queued, next} pairs: the per-env kernels fill one while the other is being reset
    int qsel = 0;
    int ngroups = 1;                  // env groups pipelined on separate streams inside one kg_step (KG_MAXGROUPS)
    int *subseq = nullptr;            // [n] substeps finished in the current RL step: the heavy lane and the light warps CLAIM an env's
                                      // substep s with atomicCAS(s -> s + 1) on it, so exactly one of them runs it
    int model_ref = 0;                // this handle holds a reference on its device's shared model tables
    int *heavy_count = nullptr;       // device: envs whose last substep of the current RL step exceeded the light workspace
    int *heavy_host = nullptr;        // pinned host copy of the previous step's count (read without synchronising)
    int heavy_ctas = 1;               // heavy-lane CTAs at the front of every split-mode substep grid
    int hand_location = 0;            // kg_set_hand_location
    int *cycle = nullptr; int cycle_n = 0;      // kg_set_shape_cycle: Latin-square shape schedule over (env, episode)
    int full_mode = 0, mode_auto = 1; // 1 = every env through the full workspace; chosen per step from heavy_host unless pinned
    std::vector<void *> allocs;
    // scratch for the *_host entry points
    float *d_action = nullptr, *d_obs = nullptr, *d_reward = nullptr, *d_euler = nullptr, *d_objp = nullptr, *d_handp = nullptr, *d_q = nullptr;
    int *d_done = nullptr, *d_mask = nullptr, *d_shape = nullptr;
    float *d_debug = nullptr, *d_ctrl = nullptr;
    float *d_obs_global = nullptr;    // [74][n] internal buffer kg_obs_global_buffer hands out
    KgForest forest = {};             // kg_set_grasp_forest: the grasp classifier's trees (device tables owned by this handle)
    int grasp_on = 0;                 // kg_set_grasp_reward: with_grasp_reward (ENV:104, 791)
    int *d_leaf = nullptr;            // [n_trees][n] leaf reached by every (tree, env) of the current step
    float *d_grasp = nullptr;         // [n] grasp_reward of the last step (info["grasp_reward"], ENV:684)
#ifndef KG_EMU
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t gstream[KG_MAXGROUPS] = {}, rstream[KG_MAXGROUPS] = {};
    cudaEvent_t gfork = nullptr, gjoin[KG_MAXGROUPS] = {}, gray0[KG_MAXGROUPS] = {}, gray1[KG_MAXGROUPS] = {};
#endif
    float last_ms = 0.f;
    int cfg_version = 0;              // bumped by every setter that changes a kernel argument (graph cache key)
#if !defined(KG_EMU)
    int use_graph = 1;                // kg_set_graph
    int pdl = 0;                      // kg_set_dependent_launch
    cudaStream_t capstream = nullptr; // origin stream of graph captures (the caller's stream may be the legacy stream, which cannot capture)
    struct GraphEntry { KgGraphKey key; cudaGraphExec_t exec; unsigned long long used; };
    std::vector<GraphEntry> graphs;
    unsigned long long graph_clock = 0;
    int timing = 0;                   // kg_set_timing: an event after every launch of a step, groups not pipelined
    std::vector<cudaEvent_t> tev;
    std::vector<int> tkind;
#endif
    int fhull_bytes = 0;       // shared-memory copy of the finger hulls in every narrowphase CTA
    int mpr_grid = 1;          // persistent narrowphase grid: resident CTAs per SM x SMs
#ifdef KG_EMU
    KgHostTables tables;
#endif
};

// ------------------------------------------------------------------------------------------------ backend helpers
#ifdef KG_EMU
static int be_alloc(kg_handle h, void **p, size_t bytes) { *p = calloc(1, bytes ? bytes : 1); h->allocs.push_back(*p); return *p ? 0 : -1; }
static int be_h2d(kg_handle, void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
static int be_d2h(kg_handle, void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
static int be_d2d(kg_handle, void *d, const void *s, size_t b) { memcpy(d, s, b); return 0; }
static int be_sync(kg_handle) { return 0; }
#define KG_MODEL(h) ((h)->model)
#else
#define CUDA_TRY(h, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { (h)->err = std::string(#x) + ": " + cudaGetErrorString(e_); return KG_ECUDA; } } while (0)
static int be_alloc(kg_handle h, void **p, size_t bytes) {
    CUDA_TRY(h, cudaMalloc(p, bytes ? bytes : 4));
    CUDA_TRY(h, cudaMemset(*p, 0, bytes ? bytes : 4));
    h->allocs.push_back(*p);
    return 0;
}
static int be_h2d(kg_handle h, void *d, const void *s, size_t b) { CUDA_TRY(h, cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, (cudaStream_t)h->stream)); return 0; }
static int be_d2h(kg_handle h, void *d, const void *s, size_t b) { CUDA_TRY(h, cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, (cudaStream_t)h->stream)); return 0; }
static int be_d2d(kg_handle h, void *d, const void *s, size_t b) { CUDA_TRY(h, cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, (cudaStream_t)h->stream)); return 0; }
static int be_sync(kg_handle h) { CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)h->stream)); return 0; }
#endif
#define TRY(x) do { int r_ = (x); if (r_ != 0) return r_ < 0 ? r_ : KG_ECUDA; } while (0)

// ------------------------------------------------------------------------------------------------ device-side sampling
KG_DEV unsigned long long kg_mix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct KgSample { int shape, orient; float euler[3], objp[3]; };
// select_orienation (ENV:1180-1222) + sample_initial_object_hand_pos (ENV:1008-1054): the reference drops
// file line 0 and draws randint(0, len(data)), i.e. file lines 1..rows-1.
KG_DEV KgSample kg_sample_pose(const KgModel &m, unsigned long long seed, int env, int episode, int shape, int orient, int variant, int test_table) {
    unsigned long long r = kg_mix(seed ^ kg_mix((unsigned long long)env * 0x100000001B3ull + (unsigned long long)episode));
    unsigned long long r2 = kg_mix(r);
    int rows = test_table ? m.pose_rows_test : m.pose_rows_train;
    const float *tab = test_table ? m.pose_test : m.pose_train;
    int o = orient;
    if (orient == 3) {
        float u = (float)((r2 >> 40) * (1.0 / 16777216.0));
        o = u < 0.333f ? 0 : (u > 0.667f ? 2 : 1);
    }
    int row = 1 + (int)(r % (unsigned long long)(rows - 1));
    KgSample s;
    s.shape = shape; s.orient = o;
    const float *p = tab + (((size_t)shape * 3 + (variant == 1 ? 0 : o)) * rows + row) * 9;
    s.objp[0] = p[0]; s.objp[1] = p[1]; s.objp[2] = p[2];
    const float *e = variant == 1 ? tab + (((size_t)shape * 3 + 2) * rows + row) * 9 : p;
    s.euler[0] = e[6]; s.euler[1] = e[7]; s.euler[2] = e[8];
    return s;
}

// determine_hand_location (ENV:1286-1307): the slide-joint start positions of the 'rotated' / 'top' orientations,
// [-x, -y, z] of Tfw[0:3,0:3] v with v = (0.051, -0.075, 0.06) / (-0.005, -0.155, Z + 0.06), Z = 0.14 (self.obj_size stays
// 'm', ENV:62).  The reference evaluates it BEFORE write_xml, i.e. with the Tfw its env object still holds from the
// PREVIOUS episode (zeros for a freshly constructed env): hq_prev is that episode's hand quaternion.
KG_DEV void kg_hand_location(const KgModel &m, const float *hq_prev, int has_prev, int orient, float *hand) {
    hand[0] = hand[1] = hand[2] = 0.f;
    if (orient == 0 || !has_prev) return;
    float Rh[9], Rp[9];
    kg_quat2mat(Rh, hq_prev);
    kg_mulmm(Rp, Rh, m.palm_gR);
    const float vx = orient == 2 ? -0.005f : 0.051f, vy = orient == 2 ? -0.155f : -0.075f, vz = orient == 2 ? 0.14f + 0.06f : 0.06f;
    // rows of Tfw[0:3,0:3] = (-Rp col1, -Rp col2, Rp col0)
    const float l0 = -(Rp[1] * vx + Rp[4] * vy + Rp[7] * vz), l1 = -(Rp[2] * vx + Rp[5] * vy + Rp[8] * vz), l2 = Rp[0] * vx + Rp[3] * vy + Rp[6] * vz;
    hand[0] = -l0; hand[1] = -l1; hand[2] = l2;
}

// one mj_forward with a dump of the intermediates (test hook behind kg_forward_debug), in the same three pieces as
// a production substep: begin (kinematics + lane collision) -> narrowphase tasks -> end (rows, solve)
KG_DEV unsigned kg_debug_begin(const KgModel &m, KgWS &ws, const KgEnvArrays &A, int env, const float *ctrl, int kg_lane) {
    kg_load_env(m, ws, A, env, kg_lane);
    KG_PHASE(l) { if (l < 9) ws.ctrl[l] = ctrl[l]; }
    KG_SYNC();
    kg_kinematics(m, ws, kg_lane);
    kg_ccache_load(ws, A.ccache + (size_t)env * KG_CCACHE_BYTES, kg_lane);
    return kg_collision_lane(m, ws, kg_lane);
}
KG_DEV void kg_debug_end(const KgModel &m, KgWS &ws, float *out, int kg_lane) {
    kg_dynamics(m, ws, kg_lane);
    kg_compaction(m, ws, kg_lane);
    kg_make_rows(m, ws, kg_lane);
    KG_PHASE(l) {      // the contact staging shares storage with the solver scratch: dump it before the solve
        if (l == 0) out[300] = (float)ws.ncon_total;
        if (l < 23) {                        // the dump has room for 23 contacts (kg.h)
            float *o = out + 301 + 8 * l;
            if (9 + l < ws.ngroup) {
                int s = kg_gslot(ws, 9 + l);
                o[0] = ws.st_dist[s]; o[1] = ws.st_pos[s][0]; o[2] = ws.st_pos[s][1]; o[3] = ws.st_pos[s][2];
                o[4] = ws.st_n[s][0]; o[5] = ws.st_n[s][1]; o[6] = ws.st_n[s][2]; o[7] = KG_GMU(ws, 9 + l) + 10.f * ws.st_g1[s] + 100.f * ws.st_g2[s];
            } else for (int c = 0; c < 8; c++) o[c] = 0.f;
        }
    }
    KG_SYNC();
    KG_PHASE(l) {      // the rows as the solver is about to see them (kg.h: layout of the dump)
        if (l == 0) out[593] = (float)ws.ngroup;
        {
            float *o = out + 594 + 8 * l;
            const int tp = kg_gtype(ws, l);
            o[0] = (float)tp; o[1] = tp ? KG_GD(ws, l) : 0.f; o[2] = tp == 3 ? KG_GMU(ws, l) : 0.f;
            for (int k = 0; k < 4; k++) o[3 + k] = (tp == 3 || (tp && k == 0)) ? KG_GAREF(ws, l, k) : 0.f;
            o[7] = (float)(l < 9 ? l : 9 + 3 * (l - 9));
        }
        const int nb = kg_nbase(ws) < 78 ? kg_nbase(ws) : 78;
        for (int i = l; i < 78 * 15; i += 32) out[850 + i] = i / 15 < nb ? ws.J[(i / 15) * KG_JS + (i % 15)] : 0.f;
    }
    KG_SYNC();
    kg_solve(m, ws, kg_lane);
    KG_PHASE(l) {
        for (int i = l; i < 225; i += 32) out[i] = ws.M[(i / 15) * KG_MS + (i % 15)];
        if (l < 15) {
            out[225 + l] = ws.bias[l]; out[240 + l] = ws.fs[l]; out[255 + l] = ws.qas[l];
            out[270 + l] = ws.qacc[l]; out[285 + l] = ws.ftot[l];
        }
        for (int i = l; i < 27; i += 32) out[485 + i] = ws.gpos[i / 3][i % 3];
        for (int i = l; i < 81; i += 32) out[512 + i] = ws.gmat[i / 9][i % 9];
    }
    KG_SYNC();
}

// ------------------------------------------------------------------------------------------------ kernels / emu loops
#ifndef KG_EMU
// Launch schedule.  Environments that share a CTA share its instruction stream (and, in builds with stage rendezvous,
// its pace); the cost of an env's step is dominated by its contact state (MPR pairs, Newton iterations) and changes
// slowly from one RL step to the next, so the warp slots are filled in order of decreasing cost of the previous step:
// the heaviest warps start first and the light ones fill the tail of the grid.  One CTA counting sort (256 keys).
#define KG_SCHED_KEYS 256
__global__ void __launch_bounds__(1024) kg_schedule_kernel(const int *load, int *perm, int n) {
    __shared__ int hist[KG_SCHED_KEYS], base[KG_SCHED_KEYS];
    for (int i = threadIdx.x; i < KG_SCHED_KEYS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[min(max(load[i], 0), KG_SCHED_KEYS - 1)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {      // exclusive prefix over the keys in DESCENDING order, 8 keys per lane
        int s[8], tot = 0;
        for (int k = 0; k < 8; k++) { s[k] = tot; tot += hist[KG_SCHED_KEYS - 1 - (8 * threadIdx.x + k)]; }
        int inc = tot;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += v; }
        for (int k = 0; k < 8; k++) base[KG_SCHED_KEYS - 1 - (8 * threadIdx.x + k)] = inc - tot + s[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[atomicAdd(&base[min(max(load[i], 0), KG_SCHED_KEYS - 1)], 1)] = i;
}

// ---- per-env kernels: one warp per environment, one workspace per warp in shared memory.
// SPLIT mode (the default): every kernel runs the LIGHT workspace (KG_MAXCON_LIGHT contact groups, KG_LWARPS warps per
// CTA, KG_LMINBLOCKS CTAs per SM -> 24 resident warps per SM instead of 16).  A substep whose contact count exceeds the
// light workspace's rows is skipped by its light warp and run, in the SAME launch, by a HEAVY LANE: the first
// `nheavy_ctas` CTAs of the grid hold full workspaces (as many as fit their shared memory), scan the contact counts of
// all envs and run the substeps the light warps skip.  Both sides apply the same test to the env's carry record, but that
// record is REWRITTEN by whoever runs the env (and a heavy CTA may finish an env before its light CTA is even scheduled),
// so the test alone cannot decide: a side that wants to run substep s of an env first claims it with
// atomicCAS(subseq[env], s, s + 1).  The claim precedes every store to the record, so a side that saw a rewritten (or
// torn) record finds the claim taken and skips; every env is run exactly once per launch.
// FULL mode: every env through the full workspace (16 warps per CTA, one CTA per SM) -- chosen by the host when the
// last RL step it has a count for had more than a third of the envs above the light limit (contact-heavy shapes).
// A launch covers the env GROUP [base, base + cnt) (launch_step pipelines the groups of a batch on separate streams);
// perm, when given, holds per group a permutation of 0..cnt-1 at perm[base..].
#define KG_ENV_OF_WARP(WS, WARPS, BLOCK) \
    extern __shared__ __align__(16) unsigned char smem[]; \
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; \
    const int slot = (BLOCK) * (WARPS) + warp; \
    if (slot >= cnt) return; \
    const int env = base + (perm ? perm[base + slot] : slot); \
    WS &ws = *(reinterpret_cast<WS *>(smem) + warp)

// first kernel of an RL step: action / on-device expert, kinematics and lane-parallel collision of substep 0
__global__ void __launch_bounds__(KG_LWARPS * 32, KG_LMINBLOCKS) kg_begin_kernel(KgEnvArrays A, int base, int cnt, const int *perm, const float *action, int mode,
                                                                                 unsigned char *carry, int *q, unsigned *tasks, int *subseq) {
    KG_PDL_PROLOGUE();
    KG_ENV_OF_WARP(KgWSLight, KG_LWARPS, blockIdx.x);
    if (lane == 0) subseq[env] = 0;
    const unsigned mask = kg_step_begin(c_model, ws, A, env, action, mode, lane);
    kg_carry_store(ws, carry + (size_t)env * KG_CARRY_BYTES, lane);
    kg_emit_tasks(mask, env, q, tasks, lane);
    kg_prof_flush(ws, A, lane);
}
// one per substep: finish substep s (dynamics, rows, Newton solve, Euler) and, unless it is the last one, start
// substep s+1 (kinematics, lane-parallel collision, narrowphase tasks); the last one ends the RL step (rays, obs, reward)
template <class WS, bool LAST>
__device__ __forceinline__ void kg_sub_body(WS &ws, const KgEnvArrays &A, int env, unsigned char *rec, int *q, unsigned *tasks,
                                            int nsub, float *obs, float *reward, int *done, float *act_out, int lane) {
    kg_substep_finish(c_model, ws, lane);
    if (LAST) kg_step_end(c_model, ws, A, env, nsub, rec, A.ranges, obs, reward, done, act_out, lane);
    else {
        const unsigned mask = kg_substep_start(c_model, ws, rec, lane);
        kg_carry_store(ws, rec, lane);
        kg_emit_tasks(mask, env, q, tasks, lane);
    }
    kg_prof_flush(ws, A, lane);
}
// contacts of the substep a carry record is about to finish = the per-pair counts the collision stage and the
// narrowphase tasks filed (one word per lane, summed over the warp)
__device__ __forceinline__ int kg_warp_sum(int v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// claim substep `sidx` of an env for this warp (see the SPLIT mode note above); warp-uniform result
__device__ __forceinline__ bool kg_claim(int *seq, int sidx, int lane) {
    int ok = 0;
    if (lane == 0) ok = atomicCAS(seq, sidx, sidx + 1) == sidx;
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
#define KG_HEAVY_WARPS ((int)(KG_LSMEM / sizeof(KgWSNoSync)))                          // heavy-lane workspaces that fit a light CTA's shared memory
#define KG_FHEAVY_WARPS ((int)(KG_WARPS * sizeof(KgWS) / sizeof(KgWSNoSync)))         // ... a full-mode CTA's
// The heavy lane: the first `nheavy_ctas` CTAs of a per-env substep grid hold KgWSNoSync workspaces (rows for every contact
// the model can have), scan the contact counts of all envs of the group and run the substeps whose count exceeds `limit`,
// the capacity of the workspace the other CTAs of this launch use.  Persistent warps, no stage rendezvous.
template <bool LAST>
__device__ __forceinline__ void kg_heavy_lane(unsigned char *smem, int nwarps, int limit, const KgEnvArrays &A, int base, int cnt, unsigned char *carry, int *q, unsigned *tasks,
                                              int nheavy_ctas, int *heavy_count, int *subseq, int sidx, int nsub, float *obs, float *reward, int *done, float *act_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nwarps) return;
    KgWSNoSync &ws = *(reinterpret_cast<KgWSNoSync *>(smem) + warp);
    const int W = nheavy_ctas * nwarps;
    int nh = 0;
    const int end = base + cnt;
    for (int e0 = base + blockIdx.x * nwarps + warp; e0 < end; e0 += 4 * W) {
        int c[4];      // four independent count loads in flight
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int env = e0 + k * W;
            c[k] = env < end ? reinterpret_cast<const int *>(carry + (size_t)env * KG_CARRY_BYTES + KG_COFF(cnt))[lane] : 0;
        }
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            if (kg_warp_sum(c[k]) <= limit) continue;
            const int env = e0 + k * W;
            if (!kg_claim(subseq + env, sidx, lane)) continue;
            unsigned char *rec = carry + (size_t)env * KG_CARRY_BYTES;
            kg_carry_load(ws, rec, lane);
            kg_sub_body<KgWSNoSync, LAST>(ws, A, env, rec, q, tasks, nsub, obs, reward, done, act_out, lane);
            __syncwarp();
            nh++;
        }
    }
    if (LAST && nh && lane == 0 && limit == KG_MAXCON_LIGHT) atomicAdd(heavy_count, nh);      // what the host picks the mode of later steps from
}
template <bool LAST>
__global__ void __launch_bounds__(KG_LWARPS * 32, KG_LMINBLOCKS) kg_sub_kernel(KgEnvArrays A, int base, int cnt, const int *perm, unsigned char *carry, int *q, unsigned *tasks,
                                                                               int nheavy_ctas, int *heavy_count, int *subseq, int sidx, int nsub, float *obs, float *reward, int *done, float *act_out) {
    KG_PDL_PROLOGUE();
    if ((int)blockIdx.x < nheavy_ctas) {
        extern __shared__ __align__(16) unsigned char smem_h[];
        kg_heavy_lane<LAST>(smem_h, KG_HEAVY_WARPS, KG_MAXCON_LIGHT, A, base, cnt, carry, q, tasks, nheavy_ctas, heavy_count, subseq, sidx, nsub, obs, reward, done, act_out);
        return;
    }
    KG_ENV_OF_WARP(KgWSLight, KG_LWARPS, blockIdx.x - nheavy_ctas);
#ifdef KG_TRACE
    const long long tr0 = clock64();
#endif
    unsigned char *rec = carry + (size_t)env * KG_CARRY_BYTES;
    kg_carry_load(ws, rec, lane);
    if (kg_warp_sum(ws.cnt[lane]) > KG_MAXCON_LIGHT) return;      // the heavy lane runs this one
    if (!kg_claim(subseq + env, sidx, lane)) return;              // ... or already has (the record was rewritten)
    kg_sub_body<KgWSLight, LAST>(ws, A, env, rec, q, tasks, nsub, obs, reward, done, act_out, lane);
#ifdef KG_TRACE
    // per (substep, env): cycles of this warp's substep, its Newton iterations and contact count, its SM and start time
    if (lane == 0 && A.prof) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *tr = A.prof + KG_NPROF + 2 * ((size_t)sidx * A.N + env);
        tr[0] = ((clock64() - tr0) << 20) | ((long long)(ws.s_iter & 0xff) << 12) | ((long long)(ws.ncon_total & 0xff) << 4);
        tr[1] = ((long long)smid << 48) | (tr0 & 0xffffffffffffll);
    }
#endif
}
// FULL mode: every env through the full workspace (KG_MAXCON contacts); the rare substep with even more contacts goes to the
// heavy lane of the same launch, exactly as in split mode.
template <bool LAST>
__global__ void __launch_bounds__(KG_WARPS * 32, KG_MINBLOCKS) kg_sub_full_kernel(KgEnvArrays A, int base, int cnt, const int *perm, unsigned char *carry, int *q, unsigned *tasks,
                                                                                  int nheavy_ctas, int *heavy_count, int *subseq, int sidx, int nsub, float *obs, float *reward, int *done, float *act_out) {
    KG_PDL_PROLOGUE();
    if ((int)blockIdx.x < nheavy_ctas) {
        extern __shared__ __align__(16) unsigned char smem_h[];
        kg_heavy_lane<LAST>(smem_h, KG_FHEAVY_WARPS, KG_MAXCON, A, base, cnt, carry, q, tasks, nheavy_ctas, heavy_count, subseq, sidx, nsub, obs, reward, done, act_out);
        return;
    }
    KG_ENV_OF_WARP(KgWS, KG_WARPS, blockIdx.x - nheavy_ctas);
    unsigned char *rec = carry + (size_t)env * KG_CARRY_BYTES;
    kg_carry_load(ws, rec, lane);
    const int total = kg_warp_sum(ws.cnt[lane]);
    if (LAST && (slot & 15) == 0 && total > KG_MAXCON_LIGHT && lane == 0) atomicAdd(heavy_count, 16);     // 1-in-16 sample of the split-mode count
    if (total > KG_MAXCON) return;                                // the heavy lane runs this one
    if (!kg_claim(subseq + env, sidx, lane)) return;
    kg_sub_body<KgWS, LAST>(ws, A, env, rec, q, tasks, nsub, obs, reward, done, act_out, lane);
}
// ---- narrowphase kernel: persistent warps pull (env, pair) tasks from the global list; every task is one full MPR run
// by all 32 lanes.  The finger hulls (the meshes every scan touches) are staged once per CTA, the object hull per warp.
#ifndef KG_MPR_WARPS
#define KG_MPR_WARPS 8
#endif
#ifndef KG_MPR_MINBLOCKS
#define KG_MPR_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(KG_MPR_WARPS * 32, KG_MPR_MINBLOCKS) kg_mpr_kernel(unsigned char *carry, int *q, int *q_other, const unsigned *tasks, long long *prof) {
    KG_PDL_PROLOGUE();
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) { q_other[0] = 0; q_other[1] = 0; }   // the list the next per-env kernel fills
    const int ntask = q[0];
    if (ntask == 0) return;
    KgMprWS &mw = *(reinterpret_cast<KgMprWS *>(smem) + KG_MPR_TEAMS * warp + lane / KG_MPR_W);      // one workspace per team
#ifndef KG_MPR_L1
    float4 *fh = reinterpret_cast<float4 *>(smem + KG_MPR_TEAMS * KG_MPR_WARPS * sizeof(KgMprWS));
    {
        const int n4 = c_model.hnpad[1] + c_model.hnpad[2];
        const float4 *src = reinterpret_cast<const float4 *>(c_model.hull_pad) + c_model.hpad0[1];
        for (int i = threadIdx.x; i < n4; i += blockDim.x) fh[i] = __ldg(src + i);
    }
    if (lane == 0) { mw.fhull = reinterpret_cast<const float *>(fh); mw.hull_shape = -1; }
#endif
#ifdef KG_PROF
    if (lane < KG_NPROF) mw.prof[lane] = 0;
#endif
#ifndef KG_MPR_L1
    __syncthreads();
#else
    __syncwarp();
#endif
#ifdef KG_PROF
    long long t0 = clock64(), nt = 0;
#endif
    const unsigned team = KG_MPR_MASK(lane);
#if KG_MPR_TEAMS == 1
    // dynamic task list: no look-ahead on purpose -- a warp that reserved its next task while still inside a long MPR
    // would hold that task back from the idle warps (measured: 10 % slower at 4096 envs)
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(q + 1, 1);
        t = __shfl_sync(team, t, 0);
        if (t >= ntask) break;
        const unsigned task = tasks[t];
        kg_mpr_task(c_model, mw, carry + (size_t)(task >> 8) * KG_CARRY_BYTES, (int)(task & 0xff), lane);
#ifdef KG_PROF
        nt++;
#endif
    }
#else
    // KG_MPR_TEAMS teams per warp in LOCKSTEP: every trip of this loop each team either picks up a new task or advances its
    // MPR by one iteration, then the warp meets again -- so teams that are in the same state of the algorithm execute its
    // (scalar, redundant) portal algebra as one instruction stream.
    KgMprIt it;
    unsigned char *rec = nullptr;
    int lp = 0, g1 = 0, g2 = 0, mesh1 = 0, mesh2 = 0, c1 = 0, c2 = 0;
    bool active = false, exhausted = false;
    for (;;) {
        if (!active && !exhausted) {
            int t = 0;
            if ((lane & (KG_MPR_W - 1)) == 0) t = atomicAdd(q + 1, 1);
            t = __shfl_sync(team, t, lane & ~(KG_MPR_W - 1));
            if (t >= ntask) exhausted = true;
            else {
                const unsigned task = tasks[t];
                rec = carry + (size_t)(task >> 8) * KG_CARRY_BYTES; lp = (int)(task & 0xff);
                kg_pair_of_lane(lp, g1, g2);
                kg_mpr_stage(mw, rec, g1, g2, lane);
                mesh1 = kg_mesh_of_geom_s(mw.shape, g1); mesh2 = kg_mesh_of_geom_s(mw.shape, g2);
                c1 = 0; c2 = 0;
                kg_mpr_init(mw, it);
                active = true;
#ifdef KG_PROF
                nt++;
#endif
            }
        }
        if (__all_sync(0xffffffffu, !active)) break;
        if (active) {
            int sepf = 0;
            float depth = 0.f, nrm[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, gap = 0.f;
            const int r = kg_mpr_iter(c_model, mw, it, mesh1, mesh2, 0.5f * c_model.margin, c1, c2, &depth, nrm, pos, &sepf, &gap, lane);
            if (r) {
                kg_mpr_file(c_model, mw, rec, lp, g1, g2, r == 1, depth, nrm, pos, sepf, gap, c1, c2, lane);
                active = false;
            }
        }
        __syncwarp();
    }
#endif
#ifdef KG_PROF
    __syncwarp();
    if (prof) {
        if (lane >= 20 && lane < 28) atomicAdd((unsigned long long *)prof + lane, (unsigned long long)mw.prof[lane]);
        if (lane == 0) { atomicAdd((unsigned long long *)prof + 28, (unsigned long long)(clock64() - t0)); atomicAdd((unsigned long long *)prof + 29, (unsigned long long)nt); }
    }
#endif
}

// ---- rangefinder kernel: the 17 rays of an RL step for a whole env group, one LANE per (site, geom) pair -- 5 warps per env,
// a warp = 4 sites x the 8 mesh geoms.  The rays see the kinematics of the last substep's INPUT state (mj_step traces them in
// its forward pass, before integrating: SURVEY 8a O-note), i.e. what the env's carry record holds once the second-to-last
// per-env kernel has stored it, so this kernel runs on a side stream NEXT TO the last narrowphase + per-env kernel instead of
// inside the latter: a BVH traversal is a chain of dependent loads, and here ~0.5 M independent ones hide each other's latency
// (in the per-env kernel 32 lanes of one warp did, and the step's last launch took 2.5x a normal substep).  Same arithmetic
// per pair as kg_rays, and the nearest hit is a min, so the readings are bit-identical.
#define KG_RAY_WARPS 5
__global__ void __launch_bounds__(KG_RAY_WARPS * 32) kg_rays_kernel(const unsigned char *carry, int base, int N, float *ranges) {
    __shared__ __align__(16) unsigned char rec[KG_CARRY_A_BYTES];
    const int env = base + blockIdx.x;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(carry + (size_t)env * KG_CARRY_BYTES);
        uint4 *dst = reinterpret_cast<uint4 *>(rec);
        for (int i = threadIdx.x; i < (int)(KG_CARRY_A_BYTES / 16); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const KgWSLight &ws = *reinterpret_cast<const KgWSLight *>(rec);      // only carry section A of it exists here (state, constants, kinematics)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = 4 * warp + (lane >> 3), g = 1 + (lane & 7);
    unsigned best = KG_RAY_NONE;
    if (i < KG_NSITE) {
        float p[3], v[3];
        kg_site_pose(c_model, ws, i, p, v);
        if ((lane & 7) == 0 && v[2] < -1e-15f) {      // the floor (mj_ray's plane case), once per site
            const float x = -p[2] / v[2];
            if (x >= 0.f && fabsf(p[0] + x * v[0]) <= 2.f && fabsf(p[1] + x * v[1]) <= 2.f) best = __float_as_uint(fabsf(x));
        }
        const int sb = c_model.site_body[i];
        const int own = sb == 0 ? 1 : (sb <= 3 ? 2 * sb : 2 * (sb - 3) + 1);
        if (g != own) {
            const float *size = g == 8 ? c_model.shape[ws.shape].size : (g == 1 ? c_model.hand_size : ((g & 1) ? c_model.hand_size + 6 : c_model.hand_size + 3));
            float lp[3], lv[3];
            kg_ray_local(ws.gpos[g], ws.gmat[g], p, v, lp, lv);
            if (!(kg_ray_box(size, lp, lv) < 0.f)) {
                const float x = kg_ray_mesh_bvh(c_model, kg_mesh_of_geom(ws, g), lp, lv);
                if (x >= 0.f) best = min(best, __float_as_uint(fabsf(x)));
            }
        }
    }
    // nearest hit of a site = min over its 8 lanes (distances are non-negative floats: their bit patterns order like integers)
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 4));
    if (i < KG_NSITE && (lane & 7) == 0) ranges[i * N + env] = best == KG_RAY_NONE ? -1.f : __uint_as_float(best);
}

__global__ void __launch_bounds__(KG_LWARPS * 32) kg_reset_kernel(KgEnvArrays A, const int *mask, const int *shape, const float *euler, const float *objp, const float *handp,
                                                                 int sampled, int orient, int variant, int test_table, int hand_location, unsigned long long seed, int *ep_count, int env_offset,
                                                                 const int *cycle, int cycle_n, float *obs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * KG_LWARPS + warp;
    if (env >= A.N) return;
    if (mask && mask[env] == 0) return;
    KgWSLight &ws = *(reinterpret_cast<KgWSLight *>(smem) + warp);
    // shape of the new episode: given, or the Latin-square schedule (row = env, column = episode: every env walks through the
    // keys, every episode index sees them uniformly -- Generate_Latin_Square's property, ENV:895-964), or unchanged
    int sh = shape ? shape[env] : (sampled && cycle_n > 0 ? cycle[(unsigned)(env + env_offset + ep_count[env]) % (unsigned)cycle_n] : A.shape[env]);
    float e[3], p[3], hp[3] = {0.f, 0.f, 0.f};
    int drawn = -1;
    if (sampled) {
        KgSample s = kg_sample_pose(c_model, seed, env + env_offset, ep_count[env], sh, orient, variant, test_table);
        for (int c = 0; c < 3; c++) { e[c] = s.euler[c]; p[c] = s.objp[c]; }
        drawn = s.orient;
        if (hand_location) {
            const float hq[4] = {A.hand_quat[env], A.hand_quat[A.N + env], A.hand_quat[2 * A.N + env], A.hand_quat[3 * A.N + env]};
            kg_hand_location(c_model, hq, ep_count[env] > 0, s.orient, hp);
        }
    } else {
        for (int c = 0; c < 3; c++) { e[c] = euler[c * A.N + env]; p[c] = objp[c * A.N + env]; if (handp) hp[c] = handp[c * A.N + env]; }
    }
    __syncwarp();
    if (lane == 0) { ep_count[env] += 1; A.orient[env] = drawn; }
    kg_env_reset(c_model, ws, A, env, sh, e, p, hp, obs, lane);
}
__global__ void kg_debug_begin_kernel(KgEnvArrays A, int env, const float *ctrl, unsigned char *carry, int *q, unsigned *tasks) {
    extern __shared__ __align__(16) unsigned char smem[];
    KgWS &ws = *reinterpret_cast<KgWS *>(smem);
    const int lane = threadIdx.x & 31;
    const unsigned mask = kg_debug_begin(c_model, ws, A, env, ctrl, lane);
    kg_carry_store(ws, carry + (size_t)env * KG_CARRY_BYTES, lane);
    kg_emit_tasks(mask, env, q, tasks, lane);
}
__global__ void kg_debug_end_kernel(int env, const unsigned char *carry, float *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    KgWS &ws = *reinterpret_cast<KgWS *>(smem);
    const int lane = threadIdx.x & 31;
    kg_carry_load(ws, carry + (size_t)env * KG_CARRY_BYTES, lane);
    kg_debug_end(c_model, ws, out, lane);
}
#else
static KgWS g_ws;
static KgMprWS g_mw;
#endif

static size_t ws_bytes() { return sizeof(KgWS); }
#ifndef KG_EMU
static_assert(KG_WARPS * sizeof(KgWS) <= 227 * 1024, "KG_WARPS workspaces must fit one SM's shared memory");
static_assert(KG_LSMEM >= 2 * sizeof(KgWSNoSync), "a light CTA's shared memory must hold at least two heavy-lane workspaces");
static_assert(KG_LMINBLOCKS * (KG_LWARPS * sizeof(KgWSLight) + 1024) <= 228 * 1024, "KG_LMINBLOCKS light CTAs must fit one SM's shared memory");
#ifdef KG_MPR_L1
#define KG_MPR_SMEM(h) (KG_MPR_TEAMS * KG_MPR_WARPS * sizeof(KgMprWS))
#else
#define KG_MPR_SMEM(h) (KG_MPR_TEAMS * KG_MPR_WARPS * sizeof(KgMprWS) + (size_t)(h)->fhull_bytes)
#endif
#endif

#ifdef KG_EMU
// host emulation of the narrowphase kernel: the tasks of list `sel` in order, then the list resets
static void emu_run_tasks(kg_handle h, int sel) {
    int *q = h->q + 2 * sel;
    for (int t = 0; t < q[0]; t++) {
        const unsigned task = h->tasks[t];
        kg_mpr_task(h->model, g_mw, h->carry + (size_t)(task >> 8) * KG_CARRY_BYTES, (int)(task & 0xff), 0);
    }
    q[0] = 0; q[1] = 0;
}
#endif

#if !defined(KG_EMU)
#define KG_TMARK(kind) do { if (!h->timing) break; if (h->tev.size() <= tix) { cudaEvent_t e_; cudaEventCreate(&e_); h->tev.push_back(e_); h->tkind.push_back(0); } h->tkind[tix] = (kind); cudaEventRecord(h->tev[tix++], st); } while (0)
#else
#define KG_TMARK(kind) ((void)0)
#endif
#ifndef KG_EMU
// kernel launch with or without the programmatic-dependent-launch attribute (see KG_PDL_PROLOGUE)
template <class... KArgs, class... Args>
static void kg_launch(bool pdl, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// Everything one RL step puts on a stream: the copy of the last step's heavy count to the host, then per env group
//   [schedule] begin | nsub x ( narrowphase tasks | per-env substep kernel )     (2 nsub + 2 launches of our kernels)
// forked over the group streams and joined back into `st` (which may be a capturing stream).
static int enqueue_step(kg_handle h, cudaStream_t st, const float *action, int mode, float *obs, float *reward, int *done, float *act_out, bool full, int qs0) {
    const int nsub = h->frame_skip;
    const size_t smem = KG_LSMEM, hsmem = KG_WARPS * ws_bytes();
    cudaMemcpyAsync(h->heavy_host, h->heavy_count, sizeof(int), cudaMemcpyDeviceToHost, st);      // last step's count -> pinned host word
    cudaMemsetAsync(h->heavy_count, 0, sizeof(int), st);
    size_t tix = 0;
    const int G = h->timing ? 1 : h->ngroups;
    // programmatic edges only between kernels that follow each other directly on a group's stream: not into the last
    // narrowphase / per-env kernel of a step (the ray fork's event record sits in front of the one, the wait for the rays in
    // front of the other), not in timing mode (an event after every launch)
    const bool pdl = h->pdl && !h->timing;
    // the ray kernel of group g, forked off the group's stream once its last carry records are written (timing mode: inline)
#define KG_RAYS(g) do { const int rb_ = gbase[g], rc_ = gbase[(g) + 1] - rb_; \
        if (h->timing) { kg_rays_kernel<<<rc_, KG_RAY_WARPS * 32, 0, gs[g]>>>(h->carry, rb_, h->n, h->A.ranges); KG_TMARK(6); } \
        else { cudaEventRecord(h->gray0[g], gs[g]); cudaStreamWaitEvent(h->rstream[g], h->gray0[g], 0); \
               kg_rays_kernel<<<rc_, KG_RAY_WARPS * 32, 0, h->rstream[g]>>>(h->carry, rb_, h->n, h->A.ranges); cudaEventRecord(h->gray1[g], h->rstream[g]); } } while (0)
    // The env groups of the batch are independent pipelines: group g's kernels go to its own stream, so the narrowphase
    // kernel of one group (latency-bound, few registers) overlaps the per-env kernel of another (and the launch gaps and
    // the tail of every kernel are covered by the other groups' work).  Fork from / join into the caller's stream.
    cudaStream_t gs[KG_MAXGROUPS];
    int gbase[KG_MAXGROUPS + 1];
    for (int g = 0; g <= G; g++) gbase[g] = (int)((long long)h->n * g / G);
    if (G > 1) {
        cudaEventRecord(h->gfork, st);
        for (int g = 0; g < G; g++) { gs[g] = h->gstream[g]; cudaStreamWaitEvent(gs[g], h->gfork, 0); }
    } else gs[0] = st;
    for (int g = 0; g < G; g++) {
        const int base = gbase[g], cnt = gbase[g + 1] - base;
        if (cnt <= 0) continue;
        const int *perm = h->schedule ? h->perm : nullptr;
        const int blocks = (cnt + KG_LWARPS - 1) / KG_LWARPS;
        KG_TMARK(0);
        if (h->schedule) kg_schedule_kernel<<<1, 1024, 0, gs[g]>>>(h->A.load + base, h->perm + base, cnt);
        KG_TMARK(1);
        kg_launch(pdl && h->schedule, kg_begin_kernel, blocks, KG_LWARPS * 32, smem, gs[g], h->A, base, cnt, perm, action, mode, h->carry, h->q + 4 * g + 2 * qs0, h->tasks + (size_t)base * 22, h->subseq);
        KG_TMARK(2);
        if (nsub == 1) KG_RAYS(g);
    }
    for (int s = 0; s < nsub; s++) {
        const int qa = (qs0 + s) & 1, qb = qa ^ 1;
        const bool last = s == nsub - 1;
        for (int g = 0; g < G; g++) {
            const int base = gbase[g], cnt = gbase[g + 1] - base;
            if (cnt <= 0) continue;
            const int *perm = h->schedule ? h->perm : nullptr;
            int *q = h->q + 4 * g;
            unsigned *tasks = h->tasks + (size_t)base * 22;
            // persistent narrowphase grid of this group: room for ~2 tasks per env at one task per team, at most the whole GPU
            int mgrid = (2 * cnt + KG_MPR_WARPS * KG_MPR_TEAMS - 1) / (KG_MPR_WARPS * KG_MPR_TEAMS);
            mgrid = mgrid < 8 ? 8 : (mgrid > h->mpr_grid ? h->mpr_grid : mgrid);
            kg_launch(pdl && s != nsub - 1, kg_mpr_kernel, mgrid, KG_MPR_WARPS * 32, KG_MPR_SMEM(h), gs[g], h->carry, q + 2 * qa, q + 2 * qb, tasks, h->A.prof);
            KG_TMARK(3);
            if (last && !h->timing) cudaStreamWaitEvent(gs[g], h->gray1[g], 0);      // this group's rays are in
            if (full) {
                const int fblocks = (cnt + KG_WARPS - 1) / KG_WARPS;
                const int hb = (cnt + 64 * KG_FHEAVY_WARPS - 1) / (64 * KG_FHEAVY_WARPS);      // more than KG_MAXCON contacts is rare: a thin lane
                const int nhc = hb < 1 ? 1 : (hb < h->heavy_ctas ? hb : h->heavy_ctas);
                if (last) kg_launch(false, kg_sub_full_kernel<true>, fblocks + nhc, KG_WARPS * 32, hsmem, gs[g], h->A, base, cnt, perm, h->carry, q + 2 * qb, tasks, nhc, h->heavy_count, h->subseq, s, nsub, obs, reward, done, act_out);
                else kg_launch(pdl, kg_sub_full_kernel<false>, fblocks + nhc, KG_WARPS * 32, hsmem, gs[g], h->A, base, cnt, perm, h->carry, q + 2 * qb, tasks, nhc, h->heavy_count, h->subseq, s, nsub, obs, reward, done, act_out);
            } else {
                const int blocks = (cnt + KG_LWARPS - 1) / KG_LWARPS;
                // heavy-lane CTAs at the front of the grid: one workspace per 8 envs of the group (they loop over whatever is heavy)
                const int hb = (cnt + 8 * KG_HEAVY_WARPS - 1) / (8 * KG_HEAVY_WARPS);
                const int nhc = hb < 1 ? 1 : (hb < h->heavy_ctas ? hb : h->heavy_ctas);
                if (last) kg_launch(false, kg_sub_kernel<true>, blocks + nhc, KG_LWARPS * 32, smem, gs[g], h->A, base, cnt, perm, h->carry, q + 2 * qb, tasks, nhc, h->heavy_count, h->subseq, s, nsub, obs, reward, done, act_out);
                else kg_launch(pdl, kg_sub_kernel<false>, blocks + nhc, KG_LWARPS * 32, smem, gs[g], h->A, base, cnt, perm, h->carry, q + 2 * qb, tasks, nhc, h->heavy_count, h->subseq, s, nsub, obs, reward, done, act_out);
            }
            KG_TMARK(last ? 5 : 4);
            if (s == nsub - 2) KG_RAYS(g);
        }
    }
    if (G > 1) for (int g = 0; g < G; g++) { cudaEventRecord(h->gjoin[g], gs[g]); cudaStreamWaitEvent(st, h->gjoin[g], 0); }
#undef KG_RAYS
    if (h->grasp_on) {      // grasp-classifier reward on the obs the last per-env kernels just wrote (kg_forest.cuh)
        const int nt = h->forest.n_trees * h->n;
        kg_forest_leaf_kernel<<<(nt + 255) / 256, 256, 0, st>>>(h->forest, h->A.obs, h->n, h->d_leaf);
        kg_grasp_decide_kernel<<<(h->n + 255) / 256, 256, 0, st>>>(h->forest, h->d_leaf, h->n, h->A.grasp_given, reward, h->d_grasp);
        KG_TMARK(7);
    }
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}
#endif
// One RL step = a wavefront of kernels on the handle's stream:
//   [schedule] begin | nsub x ( narrowphase tasks | per-env substep kernel )     (2 nsub + 1 launches of our kernels)
static int launch_step(kg_handle h, const float *action, int mode, float *obs, float *reward, int *done, float *act_out) {
    const int nsub = h->frame_skip;
#ifdef KG_EMU
    for (int env = 0; env < h->n; env++) {
        unsigned mask = kg_step_begin(h->model, g_ws, h->A, env, action, mode, 0);
        kg_carry_store(g_ws, h->carry + (size_t)env * KG_CARRY_BYTES, 0);
        kg_emit_tasks(mask, env, h->q + 2 * h->qsel, h->tasks, 0);
    }
    for (int s = 0; s < nsub; s++) {
        emu_run_tasks(h, h->qsel);
        h->qsel ^= 1;
        for (int env = 0; env < h->n; env++) {
            unsigned char *rec = h->carry + (size_t)env * KG_CARRY_BYTES;
            kg_carry_load(g_ws, rec, 0);
            kg_substep_finish(h->model, g_ws, 0);
            if (s == nsub - 1) kg_step_end(h->model, g_ws, h->A, env, nsub, rec, nullptr, obs, reward, done, act_out, 0);
            else {
                unsigned mask = kg_substep_start(h->model, g_ws, rec, 0);
                kg_carry_store(g_ws, rec, 0);
                kg_emit_tasks(mask, env, h->q + 2 * h->qsel, h->tasks, 0);
            }
        }
    }
    if (h->grasp_on) {
        for (int t = 0; t < h->forest.n_trees; t++) for (int env = 0; env < h->n; env++) h->d_leaf[(size_t)t * h->n + env] = kg_forest_leaf(h->forest, t, h->A.obs, h->n, env);
        for (int env = 0; env < h->n; env++) kg_grasp_decide(h->forest, h->d_leaf, h->n, env, h->A.grasp_given, reward, h->d_grasp);
    }
    return 0;
#else
    cudaStream_t st = (cudaStream_t)h->stream;
    // mode of this step, from the newest count of above-the-light-limit envs the host has (never waits for the device)
    if (KG_SPLIT && h->heavy_host) {
        const float frac = (float)*(volatile int *)h->heavy_host / (float)h->n;
        if (h->mode_auto) { if (!h->full_mode && frac > 0.33f) h->full_mode = 1; else if (h->full_mode && frac < 0.20f) h->full_mode = 0; }
    }
    const bool full = !KG_SPLIT || h->full_mode;
    const int qs0 = h->qsel;
    h->qsel = (qs0 + nsub) & 1;
    cudaEventRecord(h->ev0, st);
    if (h->use_graph && !h->timing) {
        // The chain of a step (2 + 2 nsub launches per env group, forked over the group streams) is captured ONCE per
        // distinct argument set into a CUDA graph and replayed: one launch call per RL step on the host instead of ~65, and
        // the dependent kernels of a group chain start without a host round trip between them.
        KgGraphKey key = {h->cfg_version, full ? 1 : 0, qs0, mode, (const void *)action, (void *)obs, (void *)reward, (void *)done, (void *)act_out};
        cudaGraphExec_t exec = nullptr;
        for (auto &g : h->graphs) if (memcmp(&g.key, &key, sizeof key) == 0) { exec = g.exec; g.used = ++h->graph_clock; break; }
        if (!exec) {
            cudaGraph_t graph = nullptr;
            CUDA_TRY(h, cudaStreamBeginCapture(h->capstream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_step(h, h->capstream, action, mode, obs, reward, done, act_out, full, qs0);
            cudaError_t ce = cudaStreamEndCapture(h->capstream, &graph);
            if (rc != 0 || ce != cudaSuccess || !graph) { h->err = std::string("graph capture: ") + cudaGetErrorString(ce); return KG_ECUDA; }
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return KG_ECUDA; }
            if (h->graphs.size() >= 8) {            // evict the least recently used instance
                size_t lru = 0;
                for (size_t i = 1; i < h->graphs.size(); i++) if (h->graphs[i].used < h->graphs[lru].used) lru = i;
                cudaGraphExecDestroy(h->graphs[lru].exec);
                h->graphs.erase(h->graphs.begin() + lru);
            }
            h->graphs.push_back({key, exec, ++h->graph_clock});
        }
        CUDA_TRY(h, cudaGraphLaunch(exec, st));
    } else {
        TRY(enqueue_step(h, st, action, mode, obs, reward, done, act_out, full, qs0));
    }
    cudaEventRecord(h->ev1, st);
    CUDA_TRY(h, cudaGetLastError());
    return 0;
#endif
}
static int launch_reset(kg_handle h, const int *mask, const int *shape, const float *euler, const float *objp, const float *handp, int sampled, int orient,
                        int variant, int test_table, unsigned long long seed, float *obs) {
#ifdef KG_EMU
    for (int env = 0; env < h->n; env++) {
        if (mask && mask[env] == 0) continue;
        int sh = shape ? shape[env] : (sampled && h->cycle_n > 0 ? h->cycle[(unsigned)(env + h->env_offset + h->ep_count[env]) % (unsigned)h->cycle_n] : h->A.shape[env]);
        float e[3], p[3], hp[3] = {0.f, 0.f, 0.f};
        int drawn = -1;
        if (sampled) {
            KgSample s = kg_sample_pose(h->model, seed, env + h->env_offset, h->ep_count[env], sh, orient, variant, test_table);
            for (int c = 0; c < 3; c++) { e[c] = s.euler[c]; p[c] = s.objp[c]; }
            drawn = s.orient;
            if (h->hand_location) {
                const int N = h->n;
                const float hq[4] = {h->A.hand_quat[env], h->A.hand_quat[N + env], h->A.hand_quat[2 * N + env], h->A.hand_quat[3 * N + env]};
                kg_hand_location(h->model, hq, h->ep_count[env] > 0, s.orient, hp);
            }
        } else for (int c = 0; c < 3; c++) { e[c] = euler[c * h->n + env]; p[c] = objp[c * h->n + env]; if (handp) hp[c] = handp[c * h->n + env]; }
        h->ep_count[env] += 1; h->A.orient[env] = drawn;
        kg_env_reset(h->model, g_ws, h->A, env, sh, e, p, hp, obs, 0);
    }
    return 0;
#else
    cudaStream_t st = (cudaStream_t)h->stream;
    int blocks = (h->n + KG_LWARPS - 1) / KG_LWARPS;
    kg_reset_kernel<<<blocks, KG_LWARPS * 32, KG_LSMEM, st>>>(h->A, mask, shape, euler, objp, handp, sampled, orient, variant, test_table, h->hand_location, seed, h->ep_count, h->env_offset, h->cycle, h->cycle_n, obs);
    CUDA_TRY(h, cudaGetLastError());
    return 0;
#endif
}

// The expert controller alone (EXP:610 get_action): one env's action from the observation the last step / reset left in
// A.obs, the controller's episode state advanced exactly as kg_step_begin does for action == NULL.  out: [act_dim][N].
KG_DEV void kg_expert_only(const KgEnvArrays &A, int env, int mode, float *out) {
    const int N = A.N;
    float obs[KG_NOBS];
    for (int i = 0; i < KG_NOBS; i++) obs[i] = A.obs[(size_t)i * N + env];
    KgExpertState e;
    e.init_obj_x = A.ex_f[env]; e.init_dot = A.ex_f[N + env];
    e.prev_fx[0] = A.ex_f[2 * N + env]; e.prev_fx[1] = A.ex_f[3 * N + env]; e.prev_fx[2] = A.ex_f[4 * N + env];
    e.have_prev = A.ex_i[env]; e.num_good = A.ex_i[N + env]; e.total_steps = A.ex_i[2 * N + env];
    float a4[4];
    kg_expert_action(e, obs, mode, a4);
    A.ex_f[2 * N + env] = e.prev_fx[0]; A.ex_f[3 * N + env] = e.prev_fx[1]; A.ex_f[4 * N + env] = e.prev_fx[2];
    A.ex_i[env] = e.have_prev; A.ex_i[N + env] = e.num_good; A.ex_i[2 * N + env] = e.total_steps;
    const int off = A.act_dim == 6 ? 2 : 0;
    if (off) { out[env] = 0.f; out[N + env] = 0.f; }
    for (int c = 0; c < 4; c++) out[(size_t)(c + off) * N + env] = a4[c];
}
#ifndef KG_EMU
__global__ void __launch_bounds__(128) kg_expert_kernel(KgEnvArrays A, int mode, float *out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < A.N) kg_expert_only(A, env, mode, out);
}
#endif
#ifndef KG_EMU
// FP32 FMA microbenchmark (the FP32 roofline denominator is measured on the box, not taken from a data sheet):
// 8 independent FMA chains per thread, 1024 threads per CTA, 2 CTAs per SM.
__global__ void __launch_bounds__(1024, 2) kg_fma_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;      // never true: keeps the chains alive
}
#endif

#ifndef KG_EMU
// The kernels read the model through the __constant__ symbol c_model, of which there is ONE copy per device, so the
// tables it points to are per DEVICE, not per handle: the first handle on a device uploads them, later handles share
// them (same model card only), the last kg_destroy frees them.  (The reference's own flow needs this: eval_policy,
// MAIN:146, builds a temporary env while the training env is alive and drops it again.)
struct KgDevModel { int refs = 0; unsigned long long hash = 0; KgModel model; std::vector<void *> allocs; };
static std::mutex g_model_mu;
static std::map<int, KgDevModel> g_models;
static unsigned long long kg_fnv1a(const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    unsigned long long x = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { x ^= b[i]; x *= 1099511628211ull; }
    return x;
}
static void model_free_locked(KgDevModel &dm) {
    for (void *p : dm.allocs) cudaFree(p);
    dm.allocs.clear(); dm.refs = 0;
}
static int model_acquire(kg_handle h, int device, const void *blob, size_t nbytes) {
    std::lock_guard<std::mutex> lk(g_model_mu);
    KgDevModel &dm = g_models[device];
    const unsigned long long hash = kg_fnv1a(blob, nbytes);
    if (dm.refs > 0) {
        if (dm.hash != hash) { h->err = "model card: another model card is live on this device (one model per device at a time)"; return KG_EMODEL; }
        h->model = dm.model; dm.refs++; h->model_ref = 1;
        return KG_OK;
    }
    KgHostTables t;
    KgModel m;
    std::string e = kg_load_model(blob, nbytes, &m, &t);
    if (!e.empty()) { h->err = "model card: " + e; return KG_EMODEL; }
    cudaError_t ce = cudaSuccess;
#define UP(field, vec) do { void *p_ = nullptr; size_t b_ = (vec).size() * sizeof((vec)[0]); \
        if (ce == cudaSuccess) ce = cudaMalloc(&p_, b_ ? b_ : 4); \
        if (ce == cudaSuccess) { dm.allocs.push_back(p_); ce = cudaMemcpy(p_, (vec).data(), b_, cudaMemcpyHostToDevice); } \
        m.field = (decltype(m.field))p_; } while (0)
    UP(hull_vert, t.hull_vert); UP(adj_start, t.adj_start); UP(adj, t.adj); UP(adj8, t.adj8); UP(nbr8, t.nbr8); UP(hull_pad, t.hull_pad); UP(hull_oct, t.hull_oct); UP(tri, t.tri);
    UP(bvh_node, t.bvh_node); UP(pose_train, t.pose_train); UP(pose_test, t.pose_test);
#undef UP
    if (ce == cudaSuccess) ce = cudaMemcpyToSymbol(c_model, &m, sizeof(KgModel));
    if (ce != cudaSuccess) { h->err = std::string("model upload: ") + cudaGetErrorString(ce); model_free_locked(dm); return KG_ECUDA; }
    dm.model = m; dm.hash = hash; dm.refs = 1;
    h->model = m; h->model_ref = 1;
    return KG_OK;
}
static void model_release(kg_handle h) {
    if (!h->model_ref) return;
    std::lock_guard<std::mutex> lk(g_model_mu);
    auto it = g_models.find(h->device);
    if (it != g_models.end() && --it->second.refs <= 0) { model_free_locked(it->second); g_models.erase(it); }
    h->model_ref = 0;
}
#endif

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int kg_create(int n_envs, int device, const void *blob, size_t nbytes, kg_handle *out) {
    if (!out || n_envs <= 0 || !blob) return KG_EINVAL;
    kg_handle h = new kg_handle_s();
    *out = h;
    h->n = n_envs; h->device = device;
#ifdef KG_EMU
    KgHostTables &t = h->tables;
    std::string e = kg_load_model(blob, nbytes, &h->model, &t);
    if (!e.empty()) { h->err = "model card: " + e; return KG_EMODEL; }
    KgModel &m = h->model;
    m.hull_vert = t.hull_vert.data(); m.adj_start = t.adj_start.data(); m.adj = t.adj.data(); m.adj8 = t.adj8.data(); m.nbr8 = t.nbr8.data(); m.hull_pad = t.hull_pad.data(); m.hull_oct = t.hull_oct.data(); m.tri = t.tri.data();
    m.bvh_node = t.bvh_node.data();
    m.pose_train = t.pose_train.data(); m.pose_test = t.pose_test.data();
    g_mw.fhull = m.hull_pad + 4 * (size_t)m.hpad0[1];     // host emulation: no staging copy
    g_mw.hull_shape = -1;
#else
    CUDA_TRY(h, cudaSetDevice(device));
    { int r_ = model_acquire(h, device, blob, nbytes); if (r_ != KG_OK) return r_; }
    KgModel &m = h->model;
#endif
    h->fhull_bytes = 16 * (m.hnpad[1] + m.hnpad[2]);
    if (m.hpad0[2] != m.hpad0[1] + m.hnpad[1]) { h->err = "model card: finger hulls must be contiguous"; return KG_EMODEL; }
#ifndef KG_MPR_L1      // (only a build that stages the object's hull in the narrowphase workspace has a size limit)
    for (int k = 3; k < KG_NMESH; k++) if (m.hnpad[k] > KG_MAXOBJVERT) { h->err = "model card: object hull too large"; return KG_EMODEL; }
#endif
#ifndef KG_EMU
    CUDA_TRY(h, cudaFuncSetAttribute(kg_begin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KG_LSMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_sub_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KG_LSMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_sub_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KG_LSMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KG_LSMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_sub_full_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KG_WARPS * ws_bytes())));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_sub_full_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KG_WARPS * ws_bytes())));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_mpr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KG_MPR_SMEM(h)));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_debug_begin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_bytes()));
    CUDA_TRY(h, cudaFuncSetAttribute(kg_debug_end_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_bytes()));
    {
        int per_sm = 0, sms = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kg_mpr_kernel, KG_MPR_WARPS * 32, KG_MPR_SMEM(h)));
        CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        h->mpr_grid = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
        const int hb = (n_envs + KG_HEAVY_WARPS - 1) / KG_HEAVY_WARPS;
        h->heavy_ctas = hb < sms ? hb : (sms > 0 ? sms : 1);
        CUDA_TRY(h, cudaHostAlloc((void **)&h->heavy_host, sizeof(int), cudaHostAllocDefault));
        *h->heavy_host = 0;
    }
    CUDA_TRY(h, cudaEventCreate(&h->ev0)); CUDA_TRY(h, cudaEventCreate(&h->ev1));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->capstream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->gfork, cudaEventDisableTiming));
    for (int g = 0; g < KG_MAXGROUPS; g++) {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->gstream[g], cudaStreamNonBlocking));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->rstream[g], cudaStreamNonBlocking));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->gjoin[g], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->gray0[g], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->gray1[g], cudaEventDisableTiming));
    }
    // pipelined env groups (kg_set_pipeline overrides): measured optimum 4 at 4096 envs, 2 once every group's launch is several waves
    h->ngroups = n_envs >= 32768 ? 2 : (n_envs >= 2048 ? 4 : (n_envs >= 512 ? 2 : 1));
#endif
    const size_t N = (size_t)n_envs;
    KgEnvArrays &A = h->A;
    A.N = n_envs; A.max_steps = 30; A.act_dim = 4;
    TRY(be_alloc(h, (void **)&A.qpos, 16 * N * 4)); TRY(be_alloc(h, (void **)&A.qvel, 15 * N * 4)); TRY(be_alloc(h, (void **)&A.qwarm, 15 * N * 4));
    TRY(be_alloc(h, (void **)&A.hand_quat, 4 * N * 4)); TRY(be_alloc(h, (void **)&A.shape, N * 4)); TRY(be_alloc(h, (void **)&A.elapsed, N * 4));
    TRY(be_alloc(h, (void **)&A.orient, N * 4)); TRY(be_alloc(h, (void **)&A.grasp_given, N * 4)); TRY(be_alloc(h, (void **)&h->d_grasp, N * 4));
    TRY(be_alloc(h, (void **)&A.ex_f, 5 * N * 4)); TRY(be_alloc(h, (void **)&A.ex_i, 3 * N * 4)); TRY(be_alloc(h, (void **)&A.obs, KG_NOBS * N * 4)); TRY(be_alloc(h, (void **)&A.last_action, 4 * N * 4));
    TRY(be_alloc(h, (void **)&A.stats, 8 * 8));
#ifdef KG_TRACE
    TRY(be_alloc(h, (void **)&A.prof, (KG_NPROF + 2 * 16 * N) * 8));
#else
    TRY(be_alloc(h, (void **)&A.prof, KG_NPROF * 8));
#endif
    TRY(be_alloc(h, (void **)&h->ep_count, N * 4));
    TRY(be_alloc(h, (void **)&A.ccache, N * KG_CCACHE_BYTES));
#ifndef KG_EMU
    TRY(be_alloc(h, (void **)&A.ranges, KG_NSITE * N * 4));
#endif
    TRY(be_alloc(h, (void **)&h->d_obs_global, KG_NOBS_GLOBAL * N * 4));
    TRY(be_alloc(h, (void **)&A.load, N * 4)); TRY(be_alloc(h, (void **)&h->perm, N * 4));
    TRY(be_alloc(h, (void **)&h->carry, N * KG_CARRY_BYTES)); TRY(be_alloc(h, (void **)&h->tasks, N * 22 * 4)); TRY(be_alloc(h, (void **)&h->q, 4 * 4 * KG_MAXGROUPS));
    TRY(be_alloc(h, (void **)&h->heavy_count, 4)); TRY(be_alloc(h, (void **)&h->subseq, N * 4));
    TRY(be_alloc(h, (void **)&h->d_action, 6 * N * 4)); TRY(be_alloc(h, (void **)&h->d_handp, 3 * N * 4)); TRY(be_alloc(h, (void **)&h->d_obs, KG_NOBS * N * 4)); TRY(be_alloc(h, (void **)&h->d_reward, N * 4));
    TRY(be_alloc(h, (void **)&h->d_done, N * 4)); TRY(be_alloc(h, (void **)&h->d_mask, N * 4)); TRY(be_alloc(h, (void **)&h->d_shape, N * 4));
    TRY(be_alloc(h, (void **)&h->d_euler, 3 * N * 4)); TRY(be_alloc(h, (void **)&h->d_objp, 3 * N * 4)); TRY(be_alloc(h, (void **)&h->d_q, 31 * N * 4));
    TRY(be_alloc(h, (void **)&h->d_debug, KG_DEBUG_FLOATS * 4)); TRY(be_alloc(h, (void **)&h->d_ctrl, 16 * 4));
    // a valid default state: every env reset to shape 0 at a neutral pose
    {
        std::vector<float> hq(4 * N, 0.f), qp(16 * N, 0.f);
        for (size_t i = 0; i < N; i++) { hq[i] = 1.f; qp[12 * N + i] = 1.f; qp[11 * N + i] = 0.05f; }
        TRY(be_h2d(h, A.hand_quat, hq.data(), hq.size() * 4)); TRY(be_h2d(h, A.qpos, qp.data(), qp.size() * 4));
        TRY(be_sync(h));
    }
    return KG_OK;
}

int kg_destroy(kg_handle h) {
    if (!h) return KG_EINVAL;
#ifdef KG_EMU
    for (void *p : h->allocs) free(p);
#else
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();      // nothing of this handle may still be running when its buffers go
    for (void *p : h->allocs) cudaFree(p);
    model_release(h);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->heavy_host) cudaFreeHost(h->heavy_host);
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.exec);
    if (h->capstream) cudaStreamDestroy(h->capstream);
    if (h->gfork) cudaEventDestroy(h->gfork);
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    for (int g = 0; g < KG_MAXGROUPS; g++) {
        if (h->gstream[g]) cudaStreamDestroy(h->gstream[g]);
        if (h->rstream[g]) cudaStreamDestroy(h->rstream[g]);
        if (h->gjoin[g]) cudaEventDestroy(h->gjoin[g]);
        if (h->gray0[g]) cudaEventDestroy(h->gray0[g]);
        if (h->gray1[g]) cudaEventDestroy(h->gray1[g]);
    }
#endif
    delete h;
    return KG_OK;
}
const char *kg_last_error(kg_handle h) { return h ? h->err.c_str() : "null handle"; }
int kg_num_envs(kg_handle h) { return h ? h->n : KG_EINVAL; }
int kg_set_stream(kg_handle h, void *stream) { if (!h) return KG_EINVAL; h->stream = stream; return KG_OK; }
int kg_sync(kg_handle h) { if (!h) return KG_EINVAL; return be_sync(h); }
int kg_set_max_episode_steps(kg_handle h, int steps) { if (!h || steps <= 0) return KG_EINVAL; h->A.max_steps = steps; h->cfg_version++; return KG_OK; }
int kg_set_obs_global(kg_handle h, float *obs74) { if (!h) return KG_EINVAL; h->A.obs_global = obs74; h->cfg_version++; return KG_OK; }
int kg_obs_global_buffer(kg_handle h, float **buf) { if (!h || !buf) return KG_EINVAL; *buf = h->d_obs_global; return KG_OK; }
int kg_get_obs_global_host(kg_handle h, float *obs74_host) {
    if (!h || !obs74_host || !h->A.obs_global) return KG_EINVAL;
    TRY(be_d2h(h, obs74_host, h->A.obs_global, (size_t)KG_NOBS_GLOBAL * h->n * 4)); TRY(be_sync(h));
    return KG_OK;
}
int kg_set_pipeline(kg_handle h, int groups) {
    if (!h || groups < 1 || groups > KG_MAXGROUPS) return KG_EINVAL;
    h->ngroups = groups; h->cfg_version++;
    return KG_OK;
}
int kg_set_mode(kg_handle h, int mode) {
    if (!h || mode < 0 || mode > 2) return KG_EINVAL;
    h->mode_auto = mode == 0; h->full_mode = mode == 2;
    return KG_OK;
}
int kg_set_frame_skip(kg_handle h, int fs) { if (!h || fs <= 0) return KG_EINVAL; h->frame_skip = fs; h->cfg_version++; return KG_OK; }
int kg_set_schedule(kg_handle h, int mode) { if (!h || mode < 0 || mode > 1) return KG_EINVAL; h->schedule = mode; h->cfg_version++; return KG_OK; }
int kg_set_env_offset(kg_handle h, int offset) { if (!h || offset < 0) return KG_EINVAL; h->env_offset = offset; return KG_OK; }
int kg_get_last_action_host(kg_handle h, float *act) {
    if (!h || !act) return KG_EINVAL;
    TRY(be_d2h(h, act, h->A.last_action, 4 * (size_t)h->n * 4));
    return be_sync(h);
}

int kg_reset_pose(kg_handle h, const int *mask, const int *shape, const float *euler, const float *objp, const float *handp, float *obs) {
    if (!h || !shape || !euler || !objp) { if (h) h->err = "kg_reset: shape/euler/objpos are required"; return KG_EINVAL; }
    return launch_reset(h, mask, shape, euler, objp, handp, 0, 0, 0, 0, 0ull, obs);
}
int kg_reset(kg_handle h, const int *mask, const int *shape, const float *euler, const float *objp, float *obs) {
    return kg_reset_pose(h, mask, shape, euler, objp, nullptr, obs);
}
int kg_set_shape_cycle(kg_handle h, const int *keys_host, int nkeys) {
    if (!h || nkeys < 0 || nkeys > 64 || (nkeys > 0 && !keys_host)) return KG_EINVAL;
    for (int i = 0; i < nkeys; i++) if (keys_host[i] < 0 || keys_host[i] >= KG_NSHAPE) { h->err = "kg_set_shape_cycle: shape id out of range"; return KG_EINVAL; }
    if (!h->cycle) TRY(be_alloc(h, (void **)&h->cycle, 64 * 4));
    if (nkeys > 0) { TRY(be_h2d(h, h->cycle, keys_host, (size_t)nkeys * 4)); TRY(be_sync(h)); }
    h->cycle_n = nkeys;
    return KG_OK;
}
int kg_set_hand_location(kg_handle h, int mode) { if (!h || mode < 0 || mode > 1) return KG_EINVAL; h->hand_location = mode; return KG_OK; }
int kg_set_action_dim(kg_handle h, int dim) {
    if (!h || (dim != 4 && dim != 6)) { if (h) h->err = "kg_set_action_dim: 4 or 6"; return KG_EINVAL; }
    h->A.act_dim = dim; h->cfg_version++;
    return KG_OK;
}
int kg_reset_sampled(kg_handle h, const int *mask, const int *shape, int orient, int variant, int test_table, uint64_t seed, float *obs) {
    if (!h || orient < 0 || orient > 3 || variant < 0 || variant > 1) { if (h) h->err = "kg_reset_sampled: bad orient/variant"; return KG_EINVAL; }
    return launch_reset(h, mask, shape, nullptr, nullptr, nullptr, 1, orient, variant, test_table ? 1 : 0, (unsigned long long)seed, obs);
}
// Grasp-classifier reward.  blob = assets/gc_forest.bin (tools/extract_forest.py: the trees of the RandomForestClassifier the
// reference unpickles at ENV:248).  The tables are copied; the blob may be freed after the call.
int kg_set_grasp_forest(kg_handle h, const void *blob_, size_t nbytes) {
    if (!h || !blob_) return KG_EINVAL;
    using namespace kgdetail;
    const unsigned char *blob = (const unsigned char *)blob_;
    if (nbytes < 16 || std::memcmp(blob, "KGMC0001", 8) != 0) { h->err = "grasp forest: bad magic"; return KG_EMODEL; }
    auto sec = [&](const char *n, uint32_t dtype, size_t esz) -> const Sec * {
        const Sec *s = find(blob, n);
        return (s && s->dtype == dtype && s->offset + (size_t)s->count * esz <= nbytes) ? s : nullptr;
    };
    const Sec *meta = sec("forest_meta", 1, 4);
    if (!meta || meta->count != 4) { h->err = "grasp forest: no forest_meta"; return KG_EMODEL; }
    const int *mv = (const int *)(blob + meta->offset);
    KgForest F = {};
    F.n_trees = mv[0]; F.n_features = mv[1]; F.n_classes = mv[2]; F.n_nodes = mv[3];
    const Sec *root = sec("tree_root", 1, 4), *fe = sec("node_feature", 1, 4), *le = sec("node_left", 1, 4), *ri = sec("node_right", 1, 4);
    const Sec *th = sec("node_thresh", 0, 8), *pr = sec("node_proba", 0, 8), *cl = sec("classes", 0, 8);
    if (F.n_trees < 1 || F.n_classes < 1 || F.n_nodes < 1 || F.n_features < 1 || F.n_features > 75 || !root || !fe || !le || !ri || !th || !pr || !cl ||
        root->count != (uint32_t)F.n_trees + 1 || fe->count != (uint32_t)F.n_nodes || le->count != fe->count || ri->count != fe->count || th->count != fe->count ||
        pr->count != (size_t)F.n_nodes * F.n_classes || cl->count != (uint32_t)F.n_classes) { h->err = "grasp forest: inconsistent sections (features must be obs[0:75])"; return KG_EMODEL; }
    // every walk must end: children inside their tree and behind their parent (sklearn numbers nodes in construction order), features in range
    const int *rt = (const int *)(blob + root->offset), *fv = (const int *)(blob + fe->offset), *lv = (const int *)(blob + le->offset), *rv = (const int *)(blob + ri->offset);
    if (rt[0] != 0 || rt[F.n_trees] != F.n_nodes) { h->err = "grasp forest: bad tree_root"; return KG_EMODEL; }
    for (int t = 0; t < F.n_trees; t++) {
        const int b = rt[t], n = rt[t + 1] - b;
        if (n < 1) { h->err = "grasp forest: empty tree"; return KG_EMODEL; }
        for (int i = 0; i < n; i++) {
            if (lv[b + i] < 0) continue;
            if (lv[b + i] <= i || lv[b + i] >= n || rv[b + i] <= i || rv[b + i] >= n || fv[b + i] < 0 || fv[b + i] >= F.n_features) { h->err = "grasp forest: bad node"; return KG_EMODEL; }
        }
    }
    void *p[7] = {};
    const Sec *all[7] = {root, fe, le, ri, th, pr, cl};
    for (int k = 0; k < 7; k++) {
        const size_t bytes = (size_t)all[k]->count * (all[k]->dtype == 0 ? 8 : 4);
        TRY(be_alloc(h, &p[k], bytes));
        TRY(be_h2d(h, p[k], blob + all[k]->offset, bytes));
    }
    TRY(be_alloc(h, (void **)&h->d_leaf, (size_t)F.n_trees * h->n * 4));
    TRY(be_sync(h));
    F.root = (const int *)p[0]; F.feature = (const int *)p[1]; F.left = (const int *)p[2]; F.right = (const int *)p[3];
    F.thresh = (const double *)p[4]; F.proba = (const double *)p[5]; F.classes = (const double *)p[6];
    h->forest = F;      // (a replaced forest's tables stay allocated until kg_destroy)
    h->cfg_version++;
    return KG_OK;
}
// set_with_grasp_reward (ENV:791; reset(with_grasp=...) ENV:1327).  The per-episode Grasp_Reward flag is cleared by every reset.
int kg_set_grasp_reward(kg_handle h, int on) {
    if (!h) return KG_EINVAL;
    if (on && !h->forest.n_trees) { h->err = "kg_set_grasp_reward: no forest loaded (kg_set_grasp_forest)"; return KG_EINVAL; }
    h->grasp_on = on ? 1 : 0; h->cfg_version++;
    return KG_OK;
}
// env.get_orientation() per env (ENV:1180-1222): what the last sampled reset drew, 0 normal / 1 rotated / 2 top; -1 = reset with explicit poses
int kg_get_orientation_host(kg_handle h, int *orient_host) {
    if (!h || !orient_host) return KG_EINVAL;
    TRY(be_d2h(h, orient_host, h->A.orient, (size_t)h->n * 4));
    return be_sync(h);
}
// info["grasp_reward"] of the last kg_step (ENV:684) and the Grasp_Reward flags (ENV:103); either pointer may be NULL
int kg_get_grasp_reward_host(kg_handle h, float *grasp_host, int *given_host) {
    if (!h) return KG_EINVAL;
    if (grasp_host) TRY(be_d2h(h, grasp_host, h->d_grasp, (size_t)h->n * 4));
    if (given_host) TRY(be_d2h(h, given_host, h->A.grasp_given, (size_t)h->n * 4));
    return be_sync(h);
}
// Grasp_net.predict on caller-supplied feature vectors (the same kernels the step uses): x_host float[n_features][m], m <= n
int kg_grasp_predict_host(kg_handle h, const float *x_host, int m, float *label_host) {
    if (!h || !x_host || !label_host || m < 1 || m > h->n) { if (h) h->err = "kg_grasp_predict_host: need 1 <= m <= n_envs"; return KG_EINVAL; }
    if (!h->forest.n_trees) { h->err = "kg_grasp_predict_host: no forest loaded (kg_set_grasp_forest)"; return KG_EINVAL; }
    TRY(be_h2d(h, h->d_obs, x_host, (size_t)h->forest.n_features * m * 4));
#ifdef KG_EMU
    for (int t = 0; t < h->forest.n_trees; t++) for (int i = 0; i < m; i++) h->d_leaf[(size_t)t * m + i] = kg_forest_leaf(h->forest, t, h->d_obs, m, i);
    for (int i = 0; i < m; i++) h->d_reward[i] = (float)kg_forest_label(h->forest, h->d_leaf, m, i);
#else
    cudaStream_t st = (cudaStream_t)h->stream;
    kg_forest_leaf_kernel<<<(h->forest.n_trees * m + 255) / 256, 256, 0, st>>>(h->forest, h->d_obs, m, h->d_leaf);
    kg_forest_label_kernel<<<(m + 255) / 256, 256, 0, st>>>(h->forest, h->d_leaf, m, h->d_reward);
    CUDA_TRY(h, cudaGetLastError());
#endif
    TRY(be_d2h(h, label_host, h->d_reward, (size_t)m * 4));
    return be_sync(h);
}
int kg_expert_action(kg_handle h, int mode, float *action) {
    if (!h || !action || mode < 0 || mode > 2) { if (h) h->err = "kg_expert_action: action buffer and expert_mode 0, 1 or 2"; return KG_EINVAL; }
#ifdef KG_EMU
    for (int env = 0; env < h->n; env++) kg_expert_only(h->A, env, mode, action);
#else
    kg_expert_kernel<<<(h->n + 127) / 128, 128, 0, (cudaStream_t)h->stream>>>(h->A, mode, action);
    CUDA_TRY(h, cudaGetLastError());
#endif
    return KG_OK;
}
int kg_expert_action_host(kg_handle h, int mode, float *action_host) {
    if (!h || !action_host) return KG_EINVAL;
    TRY(kg_expert_action(h, mode, h->d_action));
    TRY(be_d2h(h, action_host, h->d_action, (size_t)h->A.act_dim * h->n * 4));
    return be_sync(h);
}
int kg_step(kg_handle h, const float *action, int mode, float *obs, float *reward, int *done, float *act_out) {
    if (!h || mode < 0 || mode > 2) { if (h) h->err = "kg_step: expert_mode must be 0, 1 or 2"; return KG_EINVAL; }
    return launch_step(h, action, mode, obs, reward, done, act_out);
}
int kg_step_host(kg_handle h, const float *action, int mode, float *obs, float *reward, int *done) {
    if (!h) return KG_EINVAL;
    const size_t N = h->n;
    if (action) TRY(be_h2d(h, h->d_action, action, (size_t)h->A.act_dim * N * 4));
    TRY(kg_step(h, action ? h->d_action : nullptr, mode, obs ? h->d_obs : nullptr, reward ? h->d_reward : nullptr, done ? h->d_done : nullptr, nullptr));
    if (obs) TRY(be_d2h(h, obs, h->d_obs, KG_NOBS * N * 4));
    if (reward) TRY(be_d2h(h, reward, h->d_reward, N * 4));
    if (done) TRY(be_d2h(h, done, h->d_done, N * 4));
    return be_sync(h);
}
int kg_reset_pose_host(kg_handle h, const int *mask, const int *shape, const float *euler, const float *objp, const float *handp, float *obs) {
    if (!h || !shape || !euler || !objp) { if (h) h->err = "kg_reset_host: shape/euler/objpos are required"; return KG_EINVAL; }
    const size_t N = h->n;
    if (mask) TRY(be_h2d(h, h->d_mask, mask, N * 4));
    TRY(be_h2d(h, h->d_shape, shape, N * 4)); TRY(be_h2d(h, h->d_euler, euler, 3 * N * 4)); TRY(be_h2d(h, h->d_objp, objp, 3 * N * 4));
    if (handp) TRY(be_h2d(h, h->d_handp, handp, 3 * N * 4));
    TRY(kg_reset_pose(h, mask ? h->d_mask : nullptr, h->d_shape, h->d_euler, h->d_objp, handp ? h->d_handp : nullptr, obs ? h->d_obs : nullptr));
    if (obs) TRY(be_d2h(h, obs, h->d_obs, KG_NOBS * N * 4));
    return be_sync(h);
}
int kg_reset_host(kg_handle h, const int *mask, const int *shape, const float *euler, const float *objp, float *obs) {
    return kg_reset_pose_host(h, mask, shape, euler, objp, nullptr, obs);
}
int kg_get_state(kg_handle h, float *qpos, float *qvel) {
    if (!h) return KG_EINVAL;
    const size_t N = h->n;
    if (qpos) TRY(be_d2d(h, qpos, h->A.qpos, 16 * N * 4));
    if (qvel) TRY(be_d2d(h, qvel, h->A.qvel, 15 * N * 4));
    return KG_OK;
}
// A state written from outside invalidates the collision caches: their separation budgets (kg_core.cuh, slack) are bounds
// on the motion since the last collision stage and know nothing about a teleport.
static int invalidate_ccache(kg_handle h) {
#ifdef KG_EMU
    memset(h->A.ccache, 0, (size_t)h->n * KG_CCACHE_BYTES);
#else
    CUDA_TRY(h, cudaMemsetAsync(h->A.ccache, 0, (size_t)h->n * KG_CCACHE_BYTES, (cudaStream_t)h->stream));
#endif
    return 0;
}
int kg_set_state(kg_handle h, const float *qpos, const float *qvel) {
    if (!h) return KG_EINVAL;
    const size_t N = h->n;
    if (qpos) TRY(be_d2d(h, h->A.qpos, qpos, 16 * N * 4));
    if (qvel) TRY(be_d2d(h, h->A.qvel, qvel, 15 * N * 4));
    TRY(invalidate_ccache(h));
    return KG_OK;
}
int kg_get_state_host(kg_handle h, float *qpos, float *qvel) {
    if (!h) return KG_EINVAL;
    const size_t N = h->n;
    if (qpos) TRY(be_d2h(h, qpos, h->A.qpos, 16 * N * 4));
    if (qvel) TRY(be_d2h(h, qvel, h->A.qvel, 15 * N * 4));
    return be_sync(h);
}
int kg_set_state_host(kg_handle h, const float *qpos, const float *qvel) {
    if (!h) return KG_EINVAL;
    const size_t N = h->n;
    if (qpos) TRY(be_h2d(h, h->A.qpos, qpos, 16 * N * 4));
    if (qvel) TRY(be_h2d(h, h->A.qvel, qvel, 15 * N * 4));
    TRY(invalidate_ccache(h));
    return be_sync(h);
}
int kg_get_stats(kg_handle h, long long out[8]) {
    if (!h || !out) return KG_EINVAL;
    long long s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, z[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // 64-bit device counters: 32768 envs x 15 substeps wrap 2^31 in ~4.4k steps
    TRY(be_d2h(h, s, h->A.stats, sizeof s)); TRY(be_sync(h));
    TRY(be_h2d(h, h->A.stats, z, sizeof z)); TRY(be_sync(h));
    for (int i = 0; i < 8; i++) out[i] = s[i];
    return KG_OK;
}
int kg_get_prof(kg_handle h, long long out[KG_NPROF]) {
    if (!h || !out) return KG_EINVAL;
    long long z[KG_NPROF] = {0};
    TRY(be_d2h(h, out, h->A.prof, sizeof z)); TRY(be_sync(h));
    TRY(be_h2d(h, h->A.prof, z, sizeof z)); TRY(be_sync(h));
    return KG_OK;
}
int kg_get_trace(kg_handle h, long long *out) {      /* KG_TRACE builds: [16][n][2] words of the last kg_step (tools/trace_probe.py) */
    if (!h || !out) return KG_EINVAL;
#ifdef KG_TRACE
    TRY(be_d2h(h, out, h->A.prof + KG_NPROF, (size_t)2 * 16 * h->n * 8)); TRY(be_sync(h));
    return KG_OK;
#else
    return KG_EINVAL;
#endif
}
int kg_forward_debug(kg_handle h, int env, const float *ctrl, float *out) {
    if (!h || env < 0 || env >= h->n || !ctrl || !out) return KG_EINVAL;
    TRY(be_h2d(h, h->d_ctrl, ctrl, 9 * 4));
    unsigned char *rec = h->carry + (size_t)env * KG_CARRY_BYTES;
#ifdef KG_EMU
    {
        unsigned mask = kg_debug_begin(h->model, g_ws, h->A, env, h->d_ctrl, 0);
        kg_carry_store(g_ws, rec, 0);
        kg_emit_tasks(mask, env, h->q + 2 * h->qsel, h->tasks, 0);
        emu_run_tasks(h, h->qsel);
        h->qsel ^= 1;
        kg_carry_load(g_ws, rec, 0);
        kg_debug_end(h->model, g_ws, h->d_debug, 0);
    }
#else
    {
        cudaStream_t st = (cudaStream_t)h->stream;
        (void)rec;
        kg_debug_begin_kernel<<<1, 32, ws_bytes(), st>>>(h->A, env, h->d_ctrl, h->carry, h->q + 2 * h->qsel, h->tasks);
        kg_mpr_kernel<<<h->mpr_grid, KG_MPR_WARPS * 32, KG_MPR_SMEM(h), st>>>(h->carry, h->q + 2 * h->qsel, h->q + 2 * (h->qsel ^ 1), h->tasks, nullptr);
        h->qsel ^= 1;
        kg_debug_end_kernel<<<1, 32, ws_bytes(), st>>>(env, h->carry, h->d_debug);
        CUDA_TRY(h, cudaGetLastError());
    }
#endif
    TRY(be_d2h(h, out, h->d_debug, KG_DEBUG_FLOATS * 4));
    return be_sync(h);
}
#ifdef KG_EMU
// emulation-only: 0 = lanes of a phase run 0..31, 1 = 31..0 (tests/test_emu_race.py)
int kg_emu_set_lane_order(int reverse) { kg_emu_lane_order = reverse ? 1 : 0; return KG_OK; }
#endif
int kg_measure_fp32_peak(kg_handle h, float *tflops) {
    if (!h || !tflops) return KG_EINVAL;
    *tflops = 0.f;
#ifndef KG_EMU
    cudaStream_t st = (cudaStream_t)h->stream;
    int sms = 0;
    CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int iters = 4096, grid = 2 * sms;
    cudaEvent_t e0, e1;
    CUDA_TRY(h, cudaEventCreate(&e0)); CUDA_TRY(h, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {            // first repetition warms up
        cudaEventRecord(e0, st);
        kg_fma_peak_kernel<<<grid, 1024, 0, st>>>(h->d_debug, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, st);
        CUDA_TRY(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CUDA_TRY(h, cudaGetLastError());
    const double flop = 2.0 * 8.0 * 16.0 * (double)iters * 1024.0 * (double)grid;
    *tflops = (float)(flop / (best * 1e-3) / 1e12);
#endif
    return KG_OK;
}
int kg_set_dependent_launch(kg_handle h, int on) {
    if (!h) return KG_EINVAL;
#ifndef KG_EMU
    h->pdl = on ? 1 : 0; h->cfg_version++;
#endif
    return KG_OK;
}
int kg_set_graph(kg_handle h, int on) {
    if (!h) return KG_EINVAL;
#ifndef KG_EMU
    h->use_graph = on ? 1 : 0;
#endif
    return KG_OK;
}
int kg_set_timing(kg_handle h, int on) {
    if (!h) return KG_EINVAL;
#ifndef KG_EMU
    h->timing = on ? 1 : 0;
#endif
    return KG_OK;
}
int kg_timing_breakdown(kg_handle h, float out[8]) {
    if (!h || !out) return KG_EINVAL;
    for (int i = 0; i < 8; i++) out[i] = 0.f;
#ifndef KG_EMU
    if (!h->timing || h->tev.empty()) return KG_OK;
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    for (size_t i = 1; i < h->tev.size(); i++) { float ms = 0.f; cudaEventElapsedTime(&ms, h->tev[i - 1], h->tev[i]); out[h->tkind[i]] += ms; }
#endif
    return KG_OK;
}
int kg_last_step_ms(kg_handle h, float *ms) {
    if (!h || !ms) return KG_EINVAL;
#ifdef KG_EMU
    *ms = 0.f;
#else
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    CUDA_TRY(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
#endif
    return KG_OK;
}

}  // extern "C"
