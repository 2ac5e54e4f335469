#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE's own Python (unmodified, imported from
/root/reference) on a fake MjSim that serves kinematic arrays.  Runs only in the authoring
container; the fixtures are committed so the GPU box never needs /root/reference.

Pins (SURVEY.md 8c "What still pins results without MuJoCo"):
  obs_golden.npz     _get_obs('local'|'global'), _get_reward, _get_trans_mat_wrist_pose,
                     experimental_sensor, _get_obj_size  (kinova_gripper_env.py:274-343,438-746)
  expert_golden.npz  check_grasp / get_action / ExpertPIDController driven exactly like the loop at
                     expert_data.py:745-804, for the three pid modes
  learner_golden.npz DDPGfD Actor/Critic forward on a shipped checkpoint + one train_batch step
The kinematic inputs come from the fp64 oracle's rollouts (any physically sensible arrays would do:
the point is that identical inputs must give identical outputs).
"""
import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np

REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
REF = "/root/reference/gym-kinova-gripper"
OUT = os.path.join(REPO, "tests", "golden")


def install_stubs():
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, n): return _Any()

    mod("mujoco_py", MjViewer=_Any, load_model_from_path=_Any, MjSim=_Any, MjRenderContextOffscreen=_Any)
    gym = mod("gym", Env=object, make=_Any)
    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high = np.asarray(low), np.asarray(high)
    gym.spaces = mod("gym.spaces", Box=Box)
    gym.wrappers = mod("gym.wrappers")
    gym.utils = mod("gym.utils", EzPickle=object)
    gym.utils.seeding = mod("gym.utils.seeding", np_random=_Any)
    mod("glfw")
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mod("classifier_network", LinearNetwork=_Any, ReducedLinearNetwork=_Any)
    mod("PIL", Image=_Any(), ImageFont=_Any(), ImageDraw=_Any())
    mod("PIL.Image"); mod("PIL.ImageFont"); mod("PIL.ImageDraw")
    mod("tensorboardX", SummaryWriter=_Any)
    for n in ("heatmap_plot", "boxplot_plot", "heatmap_coords", "trajectory_plot"):
        mod(n, generate_heatmaps=_Any, generate_reward_boxplots=_Any, add_heatmap_coords=_Any,
            filter_heatmap_coords=_Any, plot_trajectory=_Any, coords_dict_to_array=_Any, save_coordinates=_Any)


GEOMS = ["ground", "palm", "f1_prox", "f1_dist", "f2_prox", "f2_dist", "f3_prox", "f3_dist", "object"]
SITES = ["palm", "palm_1", "palm_2", "palm_3", "palm_4", "f1_prox", "f1_prox_1", "f1_dist", "f1_dist_1",
         "f2_prox", "f2_prox_1", "f2_dist", "f2_dist_1", "f3_prox", "f3_prox_1", "f3_dist", "f3_dist_1"]


class FakeData:
    def __init__(self, geom_xpos, geom_xmat, site_xpos, hand_xpos, sensordata):
        self.geom_xpos, self.geom_xmat, self.site_xpos_, self.hand = geom_xpos, geom_xmat, site_xpos, hand_xpos
        self.sensordata = sensordata
        self.ncon = 0
    def get_geom_xpos(self, n): return self.geom_xpos[GEOMS.index(n)]
    def get_geom_xmat(self, n): return self.geom_xmat[GEOMS.index(n)].reshape(3, 3)
    def get_site_xpos(self, n): return self.site_xpos_[SITES.index(n)]
    def get_body_xpos(self, n):
        assert n == "j2s7s300_link_7"
        return self.hand


class FakeModel:
    def __init__(self, geom_size): self.geom_size = geom_size


class FakeSim:
    def __init__(self, data, model): self.data, self.model = data, model


def make_env(KinovaGripper_Env, shape):
    env = object.__new__(KinovaGripper_Env)
    env.pid = False
    env.state_rep = "local"
    env.filename = f"/kinova_description/j2s7s300_end_effector_v1_{shape}.xml"
    env.obj_size = "m"
    env.Grasp_Reward = False
    env.with_grasp_reward = False
    env.Tfw = np.zeros([4, 4])
    env.wrist_pose = np.zeros(3)
    return env


def main():
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "gym_kinova_gripper", "envs"))
    warnings.simplefilter("ignore")
    import importlib.util
    spec = importlib.util.spec_from_file_location("kinova_gripper_env", os.path.join(REF, "gym_kinova_gripper/envs/kinova_gripper_env.py"))
    kge = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kge)
    # numpy<1.25 semantics for `if xs ==[]` (ENV:328): a shape-mismatched == returned the scalar False
    # (with a DeprecationWarning); numpy 2.x raises.  Emulate the legacy behaviour through the module's
    # own `np` global instead of editing the reference source.
    class _LegacyArray(np.ndarray):
        def __eq__(self, other):
            try:
                r = np.ndarray.__eq__(self, other)
            except ValueError:
                return False
            # legacy: bool(empty array) was False (DeprecationWarning), numpy 2.x raises
            return False if isinstance(r, np.ndarray) and r.size == 0 else r
        __hash__ = None
    class _NpCompat:
        def __getattr__(self, n): return getattr(np, n)
        @staticmethod
        def append(a, v, axis=None): return np.append(a, v, axis).view(_LegacyArray)
    kge.np = _NpCompat()
    from oracle import kgo
    m = kgo.Model()
    rng = np.random.default_rng(20261017)
    pose = np.frombuffer(open(kgo.BLOB, "rb").read(), dtype=np.uint8)  # only to prove the blob exists
    del pose

    def pose_row(shape, orient, r):
        from kinovagrasping_b200.modelcard import ModelCard
        mc = pose_row.mc = getattr(pose_row, "mc", None) or ModelCard()
        return mc.pose("train")[shape, orient, r]

    # ------------------------------------------------------------------ obs / reward fixtures
    recs = {k: [] for k in ("shape", "geom_xpos", "geom_xmat", "site_xpos", "hand_xpos", "sensordata", "geom_size",
                            "obs_local", "obs_global", "reward", "done", "lift_reward", "Tfw", "wrist")}
    exp_recs = {k: [] for k in ("mode", "obs", "action", "lift", "len")}

    def record(shape, D):
        gx = D.arr("geom_xpos", (9, 3)); gm = D.arr("geom_xmat", (9, 9)); sx = D.arr("site_xpos", (17, 3))
        hx = D.arr("xpos", (9, 3))[1]; sd = D.arr("sensordata")
        gs = np.zeros((9, 3)); gs[0] = [2, 2, 1]
        sizes = m.section("mesh_geom").reshape(-1, 15)
        for g, mesh in zip(range(1, 8), (0, 1, 2, 1, 2, 1, 2)):
            gs[g] = sizes[mesh, 8:11]
        gs[8] = sizes[3 + shape, 8:11]
        env = make_env(kge.KinovaGripper_Env, kgo.SHAPES[shape])
        env._sim = FakeSim(FakeData(gx, gm, sx, hx, sd), FakeModel(gs))
        with contextlib.redirect_stdout(io.StringIO()):
            ol = np.array(env._get_obs("local"), dtype=np.float64)
            og = np.array(env._get_obs("global"), dtype=np.float64)
            rew, info, done = env._get_reward(False)
        assert ol.shape == (82,) and og.shape == (74,)
        for k, v in (("shape", shape), ("geom_xpos", gx), ("geom_xmat", gm), ("site_xpos", sx), ("hand_xpos", hx),
                     ("sensordata", sd), ("geom_size", gs), ("obs_local", ol), ("obs_global", og), ("reward", rew),
                     ("done", done), ("lift_reward", info["lift_reward"]), ("Tfw", env.Tfw.copy()), ("wrist", env.wrist_pose.copy())):
            recs[k].append(v)
        return ol

    # reference expert loop (EXP:745-804) re-driven on recorded obs sequences
    import expert_data as ED
    class _AS:  # env.action_space stand-in
        low = np.array([-0.8] * 4); high = np.array([0.8] * 4)
    class _E: action_space = _AS()

    def ref_expert_actions(obs_seq, pid_mode):
        controller = ED.ExpertPIDController(obs_seq[0])
        prev_obs = None; num_good = 0; total_steps = 0
        acts, lifts = [], []
        for obs in obs_seq:
            f_old = None if prev_obs is None else prev_obs[9:17]
            ret = ED.check_grasp(f_old, obs[9:17])
            if bool(ret[0]): num_good += 1
            lift_check = total_steps > 10 and num_good >= 1
            with contextlib.redirect_stdout(io.StringIO()):
                a = ED.get_action(obs, lift_check, controller, _E(), pid_mode)
            acts.append(np.array(a, dtype=np.float64)); lifts.append(lift_check)
            if total_steps > 0: prev_obs = obs
            total_steps += 1
        return np.array(acts), np.array(lifts)

    episodes = []
    for shape in range(14):
        for orient, variant in ((0, "A"), (2, "B"), (1, "R")):
            r = int(rng.integers(0, 4499))
            row = pose_row(shape, orient, r + 1)
            euler = row[6:9].astype(np.float64); objp = row[0:3].astype(np.float64)
            episodes.append((shape, euler, objp, int(rng.integers(0, 3))))
    # plus a few side grasps with off-centre objects so that the PD controller branches fire
    for x in (-0.07, -0.05, -0.035, -0.025, 0.0, 0.025, 0.035, 0.05, 0.07):
        for shape in (0, 1, 3):
            episodes.append((shape, np.array([-1.57, 0.0, -1.57]), np.array([x, 0.01, 0.05]), 1 if abs(x) > 0.03 else 2))
    for (shape, euler, objp, mode) in episodes:
        e, obs = m.env_reset(shape, euler, objp)
        record(shape, e.d)
        seq = [obs.copy()]
        for t in range(30):
            act, lift = m.expert_action(e, obs, mode)
            obs, r, done, info = m.env_step(e, act)
            ol = record(shape, e.d)
            assert np.allclose(ol, obs, atol=1e-9), np.abs(ol - obs).max()
            seq.append(obs.copy())
            if done: break
        seq = np.array(seq)
        for md, name in ((0, "naive"), (1, "position-dependent"), (2, "combined")):
            acts, lifts = ref_expert_actions(list(seq), name)
            pad = np.zeros((31, 82)); pad[:len(seq)] = seq
            pa = np.zeros((31, 4)); pa[:len(seq)] = acts
            pl = np.zeros(31, dtype=bool); pl[:len(seq)] = lifts
            exp_recs["mode"].append(md); exp_recs["obs"].append(pad); exp_recs["action"].append(pa)
            exp_recs["lift"].append(pl); exp_recs["len"].append(len(seq))
    # synthetic perturbations of recorded kinematics so that the branch BOUNDARIES of experimental_sensor (ENV:290-343) are
    # pinned, not only their interiors: palm ranges exactly at / just around 0.06 (strict `<`), every subset of the five
    # palm rays below the threshold, misses (-1 -> 6), and palm frames turned so that each of the three
    # argmax(|gravity_local|) cases -- incl. near-ties between two axes -- is hit
    base = [i for i in range(len(recs["shape"]))][::97]
    def turned(ax, ang):
        c, s_ = np.cos(ang), np.sin(ang)
        R = np.eye(3); a, b = [(1, 2), (0, 2), (0, 1)][ax]
        R[a, a] = c; R[a, b] = -s_; R[b, a] = s_; R[b, b] = c
        return R
    nsyn = 0
    for k, i in enumerate(base):
        shape = int(recs["shape"][i])
        gx, gm, sx, hx, sd = (np.array(recs[n][i]).copy() for n in ("geom_xpos", "geom_xmat", "site_xpos", "hand_xpos", "sensordata"))
        variants = []
        for mask in range(32):                                   # which palm rays are "close"
            v = sd.copy()
            for j in range(5):
                v[9 + j] = (0.01 + 0.009 * j) if (mask >> j) & 1 else (0.06 if j % 2 else 0.0601 + 0.01 * j)
            variants.append((gm, v))
        for val in (0.0599999, 0.06, 0.0600001, -1.0):
            v = sd.copy(); v[9:14] = val; v[14:26] = -1.0 if val < 0 else v[14:26]
            variants.append((gm, v))
        for ax in range(3):
            for ang in (np.pi / 4 - 1e-3, np.pi / 4 + 1e-3, np.pi / 2, 1.1):
                g2 = gm.copy(); g2[1] = (turned(ax, ang) @ gm[1].reshape(3, 3)).reshape(9)
                variants.append((g2, sd))
        for g2, v in variants[k % 3::3]:                         # a third of them per base record keeps the fixture small
            class _D:
                pass
            D = _D()
            D.arr = lambda name, shp=None, _g=g2, _v=v: {"geom_xpos": gx, "geom_xmat": _g, "site_xpos": sx, "sensordata": _v,
                                                          "xpos": np.vstack([np.zeros(3), hx, np.zeros((7, 3))])}[name]
            record(shape, D)
            nsyn += 1
    print("synthetic perturbation records", nsyn)
    np.savez_compressed(os.path.join(OUT, "obs_golden.npz"), **{k: np.array(v) for k, v in recs.items()})
    np.savez_compressed(os.path.join(OUT, "expert_golden.npz"), **{k: np.array(v) for k, v in exp_recs.items()})
    print("obs records", len(recs["shape"]), "expert sequences", len(exp_recs["mode"]))
    ol = np.array(recs["obs_local"])
    print("lift rewards seen", sorted(set(recs["lift_reward"])), "wrist-local max", np.abs(ol[:, 18:21]).max())
    a = np.array(exp_recs["action"])
    print("distinct expert finger values", np.unique(np.round(a[..., 1:], 4))[:20])


if __name__ == "__main__":
    main()
