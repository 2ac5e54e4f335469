/* kg.h -- C ABI of the B200-native batched KinovaGripper env step (libkg_b200.so).
 *
 * The reference has no FFI for this path: its boundary is the gym API of
 * gym_kinova_gripper.envs.kinova_gripper_env.KinovaGripper_Env (reset/step) sitting on the
 * mujoco-py object surface.  Each entry point below names the reference interface it
 * replaces.  All arrays are struct-of-arrays, field-major [field][n_envs], fp32 / int32.
 * "dev" pointers are CUDA device pointers, "host" pointers are ordinary host memory.
 * Return value: 0 on success, a negative KG_E* code otherwise; kg_last_error() gives the text.
 * A handle is bound to one device and one stream at a time (thread-compatible, not thread-safe).
 */
#ifndef KG_H
#define KG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_OK 0
#define KG_EINVAL -1
#define KG_ECUDA -2
#define KG_EMODEL -3

#define KG_NQ 16
#define KG_NVEL 15
#define KG_OBS 82
#define KG_ACT 4

typedef struct kg_handle_s *kg_handle;

/* Replaces KinovaGripper_Env.__init__ + load_model_from_path/MjSim for n_envs environments
 * (kinova_gripper_env.py:51-102, 878-879, 1002-1003).  model_blob = assets/model_card.bin.
 * kg_create makes `device` the calling thread's current CUDA device; every later call on the handle expects it to be
 * current still (one process per GPU, or cudaSetDevice before switching between handles of different devices). */
int kg_create(int n_envs, int device, const void *model_blob, size_t nbytes, kg_handle *out);
int kg_destroy(kg_handle h);
const char *kg_last_error(kg_handle h);
int kg_num_envs(kg_handle h);
/* cudaStream_t as void*; NULL = the legacy default stream. */
int kg_set_stream(kg_handle h, void *stream);
int kg_sync(kg_handle h);

/* Replaces KinovaGripper_Env.reset's physical part (kinova_gripper_env.py:1310-1410: write_xml with the
 * hand euler, _set_state, sim.forward(), first _get_obs, centroid fix-up) with explicit poses.
 * mask_dev: int[n] (non-zero = reset this env) or NULL = all.  shape_dev int[n] (0..13, model-card order),
 * euler_dev float[3][n] hand euler (MuJoCo eulerseq xyz), objpos_dev float[3][n]. obs_dev float[82][n] or NULL. */
int kg_reset(kg_handle h, const int *mask_dev, const int *shape_dev, const float *euler_dev,
             const float *objpos_dev, float *obs_dev);

/* kg_reset with the slide-joint start positions determine_hand_location (kinova_gripper_env.py:1286-1307) computes for
 * the 'rotated' / 'top' orientations and _set_state (692-703) writes into qpos[0:3]: handpos_dev float[3][n] or NULL (zeros,
 * what 'normal' and every freshly constructed env get). */
int kg_reset_pose(kg_handle h, const int *mask_dev, const int *shape_dev, const float *euler_dev,
                  const float *objpos_dev, const float *handpos_dev, float *obs_dev);
int kg_reset_pose_host(kg_handle h, const int *mask_host, const int *shape_host, const float *euler_host,
                       const float *objpos_host, const float *handpos_host, float *obs_host);
/* Shape schedule of kg_reset_sampled when shape_dev is NULL (replaces the Latin-square list the reference keeps in
 * objects.csv, kinova_gripper_env.py:895-999): episode e of global env i gets keys[(i + e) mod nkeys], shape ids in model-card
 * order; every env cycles through the keys and every episode index sees them uniformly.  nkeys = 0: envs keep their shape. */
int kg_set_shape_cycle(kg_handle h, const int *keys_host, int nkeys);
/* env.get_orientation() per env (kinova_gripper_env.py:1180-1222 select_orienation): the hand orientation the last
 * kg_reset_sampled drew for each env -- 0 normal, 1 rotated, 2 top; -1 after a reset with explicit poses.  int[n]. */
int kg_get_orientation_host(kg_handle h, int *orient_host);
/* determine_hand_location for kg_reset_sampled: 0 (default) = slides start at 0 -- what main_DDPGfD.py gets, because it
 * builds a new env per episode (MAIN:381) and a new env's Tfw is all zeros (ENV:110); 1 = what expert_data.py gets, which
 * reuses one env (EXP:696): the offsets are computed from the Tfw of the env's PREVIOUS episode (none for its first). */
int kg_set_hand_location(kg_handle h, int mode);
/* Length of the action vectors kg_step / kg_step_host take: 4 (default) = [wrist, f1, f2, f3], padded to
 * [0, 0, wrist, f1, f2, f3] as kinova_gripper_env.py:1502-1503 does; 6 = [slide_x, slide_y, slide_z, f1, f2, f3]
 * (action[0:3] is the slide vector of kinova_gripper_env.py:1518).  action_dev is float[dim][n]; action_out stays 4-d. */
int kg_set_action_dim(kg_handle h, int dim);

/* Replaces select_orienation + determine_obj_hand_coords + sample_initial_object_hand_pos
 * (kinova_gripper_env.py:1180-1283, 1008-1054): rows are drawn on the device from the with_noise tables.
 * shape_dev: int[n] or NULL (keep each env's current shape).  orient: 0 normal, 1 rotated, 2 top,
 * 3 random (U(0,1) thresholds 0.333/0.667).  variant 0 = code-faithful (euler of the drawn row, 5-char
 * truncated as write_xml does); variant 1 = SURVEY 8d "variant B": object position from the 'normal' table,
 * euler of the same row of the 'top' table.  test_table: 0 train_coords, 1 test_coords. */
int kg_reset_sampled(kg_handle h, const int *mask_dev, const int *shape_dev, int orient, int variant,
                     int test_table, uint64_t seed, float *obs_dev);

/* Replaces KinovaGripper_Env.step (kinova_gripper_env.py:1495-1552: ctrl mapping, frame_skip x sim.step(),
 * _get_obs, _get_reward) and gym's TimeLimit.  action_dev float[4][n] = [wrist, f1, f2, f3], or NULL to
 * let the on-device expert (expert_data.py:610-671; mode 0 naive, 1 position-dependent, 2 combined) act.
 * Outputs (any may be NULL): obs float[82][n], reward float[n] (= lift_reward; + grasp_reward after kg_set_grasp_reward(1)),
 * done int[n] (lift success OR elapsed >= max_episode_steps), action_out float[4][n]. */
int kg_step(kg_handle h, const float *action_dev, int expert_mode, float *obs_dev, float *reward_dev,
            int *done_dev, float *action_out_dev);

/* The same call with HOST buffers (the end-to-end path a gym-style caller uses): H2D of the actions,
 * the launch, D2H of obs/reward/done, and a stream synchronise are all inside this call. */
int kg_step_host(kg_handle h, const float *action_host, int expert_mode, float *obs_host, float *reward_host,
                 int *done_host);
/* The expert controller alone -- controller.get_action(obs) of expert_data.py:610 (PID modes EXP:367-671) as its own call,
 * for callers that keep the reference's split "action = controller(obs); env.step(action)": computes every env's action
 * from the observation the last kg_step / kg_reset* produced and advances the controller's per-episode state exactly as
 * kg_step(action = NULL, expert_mode) would (so do not mix the two within an episode step).  action: float[4][n]
 * ([wrist, f1, f2, f3]; float[6][n] with rows 0-1 zero after kg_set_action_dim(6)), ready to be passed to kg_step. */
int kg_expert_action(kg_handle h, int expert_mode, float *action_dev);
int kg_expert_action_host(kg_handle h, int expert_mode, float *action_host);
int kg_reset_host(kg_handle h, const int *mask_host, const int *shape_host, const float *euler_host,
                  const float *objpos_host, float *obs_host);

/* sim.data.qpos / qvel access (kinova_gripper_env.py:346-353, 692-703).  float[16][n] / float[15][n]. */
int kg_get_state(kg_handle h, float *qpos_dev, float *qvel_dev);
int kg_set_state(kg_handle h, const float *qpos_dev, const float *qvel_dev);
int kg_get_state_host(kg_handle h, float *qpos_host, float *qvel_host);
int kg_set_state_host(kg_handle h, const float *qpos_host, const float *qvel_host);
int kg_set_max_episode_steps(kg_handle h, int steps);      /* env._max_episode_steps (MAIN:164,384) */
int kg_set_frame_skip(kg_handle h, int frame_skip);        /* KinovaGripper_Env(frame_skip=15) */
/* Multi-GPU sharding: this handle owns global envs [offset, offset+n).  Only the pose sampler's counter-based RNG
 * looks at it, so that a run's draws do not depend on how many GPUs the envs are spread over. */
int kg_set_env_offset(kg_handle h, int offset);
/* _get_obs(state_rep="global") (kinova_gripper_env.py:484-534): while obs74 is non-NULL every kg_step / kg_reset* also
 * writes the 74-value global-frame observation of the envs it touches to obs74 (float[74][n]; a device pointer, a host
 * pointer in the test-only emulation).  The 82-value local layout is what step() returns in the reference (ENV:1549). */
int kg_set_obs_global(kg_handle h, float *obs74);
int kg_obs_global_buffer(kg_handle h, float **buf);           /* the handle's own float[74][n] buffer, for kg_set_obs_global */
int kg_get_obs_global_host(kg_handle h, float *obs74_host);   /* copy of the buffer currently set; KG_EINVAL if none */
/* Launch schedule of kg_step (no effect on results): 1 (default) = warp slots are filled in order of decreasing cost of
 * each env's previous step, so that the environments sharing a CTA's stage barriers carry similar contact loads;
 * 0 = env i runs in slot i. */
int kg_set_schedule(kg_handle h, int mode);
/* Workspace mode of kg_step (no effect on results beyond fp32 rounding): 0 (default) = automatic -- split mode (light
 * per-env workspaces, more resident warps; substeps with more contacts than a light workspace holds run in the heavy
 * lane of the same launch) unless the last RL step the host has a count for had more than a third of the envs above the
 * light limit, then full mode; 1 = always split; 2 = always full (every env through the full workspace). */
int kg_set_mode(kg_handle h, int mode);
/* Pipelining of kg_step (no effect on results): the batch is cut into `groups` (1..32) contiguous env groups whose kernel
 * chains run on separate internal streams, forked from and joined into the handle's stream, so one group's narrowphase
 * kernel overlaps another group's per-env kernel.  Default: 2 from 32768 envs, 4 from 2048 envs, 2 from 512 envs, else 1. */
int kg_set_pipeline(kg_handle h, int groups);
/* Launch path of kg_step (no effect on results): 1 (default) = the kernel chain of a step is captured once per distinct
 * argument set into a CUDA graph and replayed with one launch call; 0 = every kernel launched individually. */
int kg_set_graph(kg_handle h, int on);
/* Programmatic dependent launch inside a group's kernel chain (no effect on results): 1 = the narrowphase / per-env kernels
 * that directly follow another kernel of the chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so
 * their launch and CTA dispatch overlap the predecessor (they wait for its completion on the device, griddepcontrol.wait). */
int kg_set_dependent_launch(kg_handle h, int on);
/* Diagnostics (bench.py's roofline): with timing on, kg_step records a CUDA event after every kernel launch (and does
 * not pipeline groups); kg_timing_breakdown then returns the milliseconds of the last step by launch kind:
 * [1] schedule [2] begin [3] narrowphase (all launches) [4] substep kernels except the last [5] last substep kernel
 * [6] rangefinder kernel (in timing mode it runs inline; normally on a side stream next to the last narrowphase + substep). */
/* FP32 FMA microbenchmark on the handle's device (8 independent chains per thread, 2 x 1024 threads per SM): the measured
 * denominator of bench.py's FP32 roofline fraction, in TFLOP/s (2 flop per FMA).  0 in the test-only emulation. */
int kg_measure_fp32_peak(kg_handle h, float *tflops);
int kg_set_timing(kg_handle h, int on);
int kg_timing_breakdown(kg_handle h, float ms_by_kind[8]);
/* Grasp-classifier reward (kinova_gripper_env.py:248 `self.Grasp_net = pickle.load(.../gc_model.pkl)`, 643-659
 * `_get_reward(with_grasp_reward)`, 791 set_with_grasp_reward, 1327 reset(with_grasp=...), 1372 Grasp_Reward=False).
 * kg_set_grasp_forest: the RandomForestClassifier's trees as the data blob tools/extract_forest.py writes
 * (kinovagrasping_b200/assets/gc_forest.bin; host memory, copied).  kg_set_grasp_reward(1): every kg_step then labels
 * obs[0:75] of each env with the forest and adds 5.0 to the reward the first time per episode the label is >= 0.3
 * (label = scikit-learn's RandomForestClassifier.predict, bit-exact); every kg_reset* clears the per-episode flag of the
 * envs it resets.  kg_get_grasp_reward_host: info["grasp_reward"] of the last step (float[n]) and the Grasp_Reward flags
 * (int[n]); either may be NULL. */
int kg_set_grasp_forest(kg_handle h, const void *forest_blob, size_t nbytes);
int kg_set_grasp_reward(kg_handle h, int on);
int kg_get_grasp_reward_host(kg_handle h, float *grasp_host, int *given_host);
/* Grasp_net.predict (ENV:651) on caller-supplied samples with the kernels the step uses: x_host = float[n_features][m]
 * (feature-major, m <= n_envs), label_host = float[m] class labels.  Needs kg_set_grasp_forest. */
int kg_grasp_predict_host(kg_handle h, const float *x_host, int m, float *label_host);
/* The action the last kg_step applied (the caller's, or the on-device expert's: get_action, expert_data.py:610). float[4][n]. */
int kg_get_last_action_host(kg_handle h, float *action_host);

/* counters since the last call: [0] hull pairs sent to MPR, [1] Newton iterations, [2] substeps that
 * overflowed the per-env contact capacity, [3] substeps executed, [4] contacts (summed over substeps), [5..7] reserved. */
int kg_get_stats(kg_handle h, long long stats_host[8]);

/* Profiling builds (-DKG_PROF) only; zeros otherwise.  32 slots since the last call, summed over envs.
 * SM cycles: [0] kinematics+dynamics [1] barrier wait before collision [2] collision (broadphase, plane, separating-axis
 * re-test, MPR) [3] contact compaction [4] wait before rows [5] constraint rows [6] wait before solve [7] solver
 * set-up (M^-1, warm start) [8] cost phase [9] gradient+Hessian [10] Cholesky factor+solve [11] directional products
 * [12] line search [13] solver epilogue [14] wait before Euler [15] Euler [16] rays [17] obs [18] wait before kinematics
 * [19] store/load.  Counters: [20] support calls [21] hill-climb moves [22] MPR calls [23] MPR loop iterations
 * [24] separating-axis re-tests [25] of which still separated [26] plane supports [27] line-search evaluations. */
#define KG_NPROF 32
int kg_get_prof(kg_handle h, long long slots_host[KG_NPROF]);

/* Test hook: one mj_forward of env `env` with the given ctrl[9]; dumps intermediates to host memory.
 * out layout (floats): M[225] bias[15] fsmooth[15] qacc_smooth[15] qacc[15] ftotal[15] ncon[1]
 * con[23][8]{dist,pos3,n3,mu} geom_xpos[27] geom_xmat[81]  -> 225+75+1+184+108 = 593 floats, then the constraint rows
 * as the solver holds them (kg_make_rows): [593] ngroup, [594 + 8 g]{type (0 inactive, 1 tendon equality, 2 joint limit,
 * 3 pyramidal contact), D, mu, aref[4], first base row} for the groups g < 32, [850 + 15 b] the base Jacobian rows b < 78
 * (a scalar group has one, a contact three: normal, tangent 1, tangent 2; its four rows are n +- mu t1, n +- mu t2).
 * Groups beyond the dump's 32 (more than 23 contacts) are not dumped. */
#define KG_DEBUG_FLOATS 2020
int kg_forward_debug(kg_handle h, int env, const float *ctrl_host, float *out_host);
/* Test hook: device time of the last kg_step launch in milliseconds (CUDA events on the handle's stream). */
int kg_last_step_ms(kg_handle h, float *ms);

#ifdef __cplusplus
}
#endif
#endif
