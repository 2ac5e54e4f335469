"""Pins the oracle's obs / reward / expert-controller algebra against fixtures produced by the
reference's own Python (tools/gen_golden.py; kinova_gripper_env.py:274-343,438-746, expert_data.py:318-671)."""
import ctypes as C
import os

import numpy as np


def _fill(kgo, g, i):
    D = kgo.Data()
    for name, src in (("geom_xpos", g["geom_xpos"][i]), ("geom_xmat", g["geom_xmat"][i]), ("site_xpos", g["site_xpos"][i]),
                      ("sensordata", g["sensordata"][i])):
        a = np.ascontiguousarray(src, dtype=np.float64).ravel()
        C.memmove(getattr(D, name), a.ctypes.data, a.nbytes)
    xp = np.zeros(27); xp[3:6] = g["hand_xpos"][i]
    C.memmove(D.xpos, xp.ctypes.data, xp.nbytes)
    return D


def test_obs_reward_match_reference_python(oracle, golden_dir):
    from oracle import kgo
    g = dict(np.load(os.path.join(golden_dir, "obs_golden.npz")))
    n = len(g["shape"])
    assert n > 1000
    worst_l = worst_g = 0.0
    for i in range(n):
        D = _fill(kgo, g, i)
        s = kgo.State(); s.shape = int(g["shape"][i])
        ol = oracle.get_obs(s, D, False)
        og = oracle.get_obs(s, D, True)
        worst_l = max(worst_l, np.abs(ol - g["obs_local"][i]).max())
        worst_g = max(worst_g, np.abs(og - g["obs_global"][i]).max())
        z = og[23]
        done = (abs(z - 0.2) < 0.005) or (z >= 0.2)
        assert done == bool(g["done"][i])
        assert (50.0 if done else 0.0) == g["reward"][i] == g["lift_reward"][i]
    assert worst_l < 1e-9 and worst_g < 1e-9, (worst_l, worst_g)


def test_obs_structure_known_answers(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "obs_golden.npz")))
    ol = g["obs_local"]
    assert ol.shape[1] == 82 and g["obs_global"].shape[1] == 74
    assert np.abs(ol[:, 18:21]).max() < 1e-12          # wrist in its own frame (ENV:285-286,515-517)
    assert set(np.unique(g["reward"])) <= {0.0, 50.0}
    assert np.all((ol[:, 50:67] > 0) & (ol[:, 50:67] <= 6))


def test_expert_controller_matches_reference(oracle, golden_dir):
    from oracle import kgo
    g = dict(np.load(os.path.join(golden_dir, "expert_golden.npz")))
    nseq = len(g["mode"])
    branches = set()
    for k in range(nseq):
        e = kgo.Env()
        obs0 = g["obs"][k][0]
        e.init_obj_x = obs0[21]; e.init_dot = obs0[81]
        for t in range(int(g["len"][k])):
            act, lift = oracle.expert_action(e, g["obs"][k][t], int(g["mode"][k]))
            assert np.array_equal(act, g["action"][k][t]), (k, t, act, g["action"][k][t])
            assert lift == bool(g["lift"][k][t])
            branches.add((int(g["mode"][k]), lift, tuple(np.round(act[1:], 2) == 0.5)))
    assert len(branches) >= 6
