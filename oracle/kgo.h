/* kgo.h -- fp64 CPU ORACLE for the KinovaGrasping env-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kinovagrasping_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.  PARITY UNPINNED against MuJoCo 1.50: the engine the reference calls
 * (mujoco-py 1.50.1.0 -> libmujoco150.so, kinova_gripper_env.py:21,102,1535) is a closed
 * binary that is not in /root/reference and not installable here, and the reference ships
 * no recorded trajectories.  This file restates the published MuJoCo algorithm
 * (SURVEY.md Appendix A).  What IS pinned: the obs/reward/controller algebra, against the
 * reference's own Python executed on this oracle's kinematics (tests/golden/).
 */
#ifndef KGO_H
#define KGO_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGO_NQ 16
#define KGO_NV 15
#define KGO_NU 9
#define KGO_NBODY 9   /* world, link_7, f1, t1, f2, t2, f3, t3, object */
#define KGO_NGEOM 9   /* ground, palm, f1_prox, f1_dist, f2_prox, f2_dist, f3_prox, f3_dist, object */
#define KGO_NSITE 17
#define KGO_NSENSOR 26
#define KGO_MAXCON 64
#define KGO_MAXEFC (3 + 6 + 4 * KGO_MAXCON)
#define KGO_NSHAPE 24
#define KGO_NOBS 82

typedef struct KgoModel KgoModel;

typedef struct {
    int geom1, geom2;
    double dist, pos[3], frame[9], friction, includemargin;
} KgoContact;

/* everything mj_forward produces that a parity test may want to look at */
typedef struct {
    double xpos[KGO_NBODY * 3], xmat[KGO_NBODY * 9], xipos[KGO_NBODY * 3], ximat[KGO_NBODY * 9];
    double geom_xpos[KGO_NGEOM * 3], geom_xmat[KGO_NGEOM * 9];
    double site_xpos[KGO_NSITE * 3], site_zaxis[KGO_NSITE * 3];
    double M[KGO_NV * KGO_NV];
    double qfrc_bias[KGO_NV], qfrc_passive[KGO_NV], qfrc_actuator[KGO_NV], qfrc_smooth[KGO_NV];
    double qacc_smooth[KGO_NV], qacc[KGO_NV], qfrc_constraint[KGO_NV];
    int ncon;
    KgoContact con[KGO_MAXCON];
    int nefc;
    double efc_J[KGO_MAXEFC * KGO_NV], efc_pos[KGO_MAXEFC], efc_margin[KGO_MAXEFC];
    double efc_D[KGO_MAXEFC], efc_R[KGO_MAXEFC], efc_aref[KGO_MAXEFC], efc_force[KGO_MAXEFC];
    int efc_type[KGO_MAXEFC]; /* 0 equality, 1 limit, 2 pyramidal contact */
    double sensordata[KGO_NSENSOR];
    int solver_iter, n_mpr, n_broad;
    double solver_cost;
} KgoData;

typedef struct {
    double qpos[KGO_NQ], qvel[KGO_NV], qacc_warmstart[KGO_NV];
    double hand_quat[4]; /* orientation of link_7 (what write_xml edits, ENV:851-881) */
    int shape;
} KgoState;

/* flags for kgo_forward */
#define KGO_RAYS 1

KgoModel *kgo_model_load(const void *blob, size_t nbytes);
void kgo_model_free(KgoModel *m);
int kgo_sizeof_data(void);
int kgo_sizeof_state(void);

void kgo_state_init(const KgoModel *m, KgoState *s, int shape, const double hand_euler[3],
                    const double obj_pos[3]);
/* mj_forward */
void kgo_forward(const KgoModel *m, const KgoState *s, const double ctrl[KGO_NU], int flags, KgoData *d);
/* test hook: mj_forward with the contact list given by the caller, ncon x {geom1, geom2, dist, pos[3], normal[3]} */
void kgo_forward_contacts(const KgoModel *m, const KgoState *s, const double ctrl[KGO_NU], int flags, int ncon,
                          const double *con, KgoData *d);
/* mj_step = forward + Euler; d holds the forward results (pre-integration, the staleness of SURVEY 8a O-note) */
void kgo_step(const KgoModel *m, KgoState *s, const double ctrl[KGO_NU], int flags, KgoData *d);
void kgo_step_contacts(const KgoModel *m, KgoState *s, const double ctrl[KGO_NU], int flags, int ncon, const double *con,
                       KgoData *d);

/* ---- env layer (ENV:1495-1552, 438-534, 631-687) */
typedef struct {
    KgoState st;
    KgoData d;          /* last forward */
    int elapsed;        /* TimeLimit counter (gym_kinova_gripper/__init__.py:3-7) */
    int max_steps;      /* 30 (MAIN:164,384; EXP:729) */
    /* expert-controller episode state (EXP:367-379, 724-766) */
    double init_obj_x, init_dot, prev_fx[3];
    int have_prev, have_prev2, num_good_grasps, total_steps;
    double obs0_fx[3];
} KgoEnv;

int kgo_sizeof_env(void);
void kgo_env_reset(const KgoModel *m, KgoEnv *e, int shape, const double hand_euler[3],
                   const double obj_pos[3], double obs[KGO_NOBS]);
/* action: 4 values [wrist, f1, f2, f3].  returns done (incl. TimeLimit). info3 = finger,grasp,lift */
int kgo_env_step(const KgoModel *m, KgoEnv *e, const double action[4], double obs[KGO_NOBS],
                 double *reward, double info3[3]);
/* 6-d actions [slide_x, slide_y, slide_z, f1, f2, f3] (ENV:1502-1503, 1518-1522) and reset with slide start positions */
int kgo_env_step6(const KgoModel *m, KgoEnv *e, const double action[6], double obs[KGO_NOBS],
                  double *reward, double info3[3]);
void kgo_env_reset_pose(const KgoModel *m, KgoEnv *e, int shape, const double hand_euler[3],
                        const double obj_pos[3], const double *hand_pos, double obs[KGO_NOBS]);
void kgo_ctrl_from_action6(const KgoData *d, const double action[6], double ctrl[KGO_NU]);
void kgo_hand_location(const KgoData *d_prev, int orient, double hand[3]);
void kgo_set_rays_every_substep(int on);
void kgo_get_obs(const KgoModel *m, const KgoState *s, const KgoData *d, int global_rep, double *obs /*82 or 74*/);
void kgo_ctrl_from_action(const KgoData *d, const double action[4], double ctrl[KGO_NU]);
/* expert controller (EXP:610-671): mode 0 naive, 1 position-dependent, 2 combined.
 * Advances the controller's episode state exactly like the loop at EXP:745-804. */
void kgo_expert_action(KgoEnv *e, const double obs[KGO_NOBS], int mode, double action[4], int *lift_check);

/* stand-alone pieces for unit tests */
int kgo_mpr(const KgoModel *m, int mesh1, const double pos1[3], const double mat1[9], int mesh2,
            const double pos2[3], const double mat2[9], double margin, double *depth, double dir[3],
            double pos[3]);
double kgo_ray_mesh(const KgoModel *m, int mesh, const double gpos[3], const double gmat[9],
                    const double size[3], const double pnt[3], const double vec[3], int use_bvh);
const double *kgo_model_section(const KgoModel *m, const char *name, int *count);
int kgo_mesh_of_geom(const KgoModel *m, int geom, int shape);

#ifdef __cplusplus
}
#endif
#endif
