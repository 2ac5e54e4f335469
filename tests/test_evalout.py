"""Batched eval outputs (MAIN:130-272 eval_policy's return value, MAIN:516-528 + plotting_code/heatmap_coords.py
files) on the host emulation; the heatmap files are compared with what the reference's OWN filter_heatmap_coords writes
from the same dictionary (build container only)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import _emu_lib
from kinovagrasping_b200 import env as kenv
from kinovagrasping_b200 import shard
from kinovagrasping_b200.evalout import eval_policy, save_eval_outputs

REF = "/root/reference/gym-kinova-gripper"


def _policy(obs):
    a = torch.zeros((obs.shape[0], 4))
    a[:, 1:] = 0.6
    return a


def test_eval_policy_dictionary_and_files(tmp_path):
    n = 24
    env = shard.make_sharded_env(n, 0, 1, "cpu", lib=_emu_lib())
    env.set_with_grasp_reward(True)
    ret = eval_policy(env, _policy, ("CubeS", "CubeB", "CylinderM"), "random", variant="B", mode="test", seed=3)
    assert set(ret) == {"avg_reward", "avg_rewards", "all_ep_reward_values", "num_success", "success_coords", "fail_coords"}
    v = ret["all_ep_reward_values"]
    assert all(len(v[k]) == n for k in ("total_reward", "finger_reward", "grasp_reward", "lift_reward"))
    tot = np.array(v["total_reward"]); lift = np.array(v["lift_reward"]); grasp = np.array(v["grasp_reward"])
    assert np.array_equal(tot, lift + grasp) and set(np.unique(lift)) <= {0.0, 50.0} and set(np.unique(grasp)) <= {0.0, 5.0}
    assert ret["num_success"] == int((lift == 50.0).sum()) == len(ret["success_coords"]["x"]) and len(ret["fail_coords"]["x"]) == n - ret["num_success"]
    assert 0 < ret["num_success"] < n and grasp.sum() > 0                    # both outcomes occur, the classifier pays somewhere
    assert ret["avg_reward"] == pytest.approx(tot.mean()) and ret["avg_rewards"]["lift_reward"] == pytest.approx(lift.mean())
    ori = np.concatenate([ret["success_coords"]["orientation"], ret["fail_coords"]["orientation"]])
    assert set(ori) <= {"normal", "rotated", "top"} and len(set(ori)) >= 2   # 'random' draws per env (ENV:1208-1220)
    hm, bp = str(tmp_path / "heatmap" / "eval"), str(tmp_path / "boxplot" / "eval")
    save_eval_outputs(ret, 4000, hm, bp)
    for key in ("finger_reward", "grasp_reward", "lift_reward", "total_reward"):
        a = np.load(os.path.join(bp, key + "_4000.npy"))
        assert a.shape == (1, 1, n) and np.array_equal(a[0, 0], np.array(v[key]))      # the [[values]] nesting of MAIN:361-364,510-513
    assert "Total # Success: %d" % ret["num_success"] in open(os.path.join(hm, "heatmap_info.txt")).read()
    if os.path.isdir(REF):
        sys.path.insert(0, os.path.join(REF, "plotting_code"))
        import heatmap_coords as HC                                           # the reference's own writer, unmodified
        ref_dir = str(tmp_path / "ref")
        os.makedirs(ref_dir)
        HC.filter_heatmap_coords(ret["success_coords"], ret["fail_coords"], 4000, ref_dir)
        ours = sorted(os.path.relpath(os.path.join(d, f), hm) for d, _, fs in os.walk(hm) for f in fs)
        theirs = sorted(os.path.relpath(os.path.join(d, f), ref_dir) for d, _, fs in os.walk(ref_dir) for f in fs)
        assert ours == theirs and len(ours) >= 9
        for f in ours:
            if f.endswith(".npy"):
                assert np.array_equal(np.load(os.path.join(hm, f)), np.load(os.path.join(ref_dir, f))), f
            else:
                assert open(os.path.join(hm, f)).read().replace(hm, "") == open(os.path.join(ref_dir, f)).read().replace(ref_dir, "")
    env.h.close()


@pytest.mark.gpu
def test_eval_policy_on_cuda(tmp_path):
    """The same dictionary and files from the CUDA library, 1024 envs, all 21 shapes, with the grasp-classifier reward on; the
    reward columns are consistent with each other and with the success count, and the orientation split covers the batch."""
    n = 1024
    env = shard.make_sharded_env(n, 0, 1, 0)
    env.set_with_grasp_reward(True)
    dev = env.device

    def policy(obs):
        a = torch.zeros((obs.shape[0], 4), device=dev)
        a[:, 1:] = 0.6
        return a
    from kinovagrasping_b200.modelcard import SHAPES
    ret = eval_policy(env, policy, tuple(SHAPES), "random", variant="B", mode="test", seed=5)
    v = ret["all_ep_reward_values"]
    tot = np.array(v["total_reward"]); lift = np.array(v["lift_reward"]); grasp = np.array(v["grasp_reward"])
    assert np.array_equal(tot, lift + grasp) and set(np.unique(lift)) <= {0.0, 50.0} and set(np.unique(grasp)) <= {0.0, 5.0}
    assert ret["num_success"] == int((lift == 50.0).sum()) and 0.05 * n < ret["num_success"] < 0.95 * n and grasp.sum() >= 5.0 * 0.05 * n
    assert len(ret["success_coords"]["x"]) + len(ret["fail_coords"]["x"]) == n
    ori = np.concatenate([ret["success_coords"]["orientation"], ret["fail_coords"]["orientation"]])
    assert all(0.2 * n < int((ori == o).sum()) < 0.5 * n for o in ("normal", "rotated", "top"))      # thresholds 0.333 / 0.667 (ENV:1208-1220)
    save_eval_outputs(ret, 0, str(tmp_path / "heatmap"), str(tmp_path / "boxplot"))
    assert np.load(str(tmp_path / "boxplot" / "total_reward_0.npy")).shape == (1, 1, n)
    env.h.close()


def test_start_position_in_the_hand_frame_is_reset_obs_21_24():
    """MAIN:171-174 `np.matmul(eval_env.Tfw, np.append(obj_coords, 1))[0:3]` == the object entries of the reset observation."""
    kenv._MEM_FILES.clear()
    e = kenv.make(lib=_emu_lib())
    for orient in ("normal", "top"):
        state = e.reset(shape_keys=["CubeB"], hand_orientation=orient, mode="test")
        local = np.matmul(e.Tfw, np.append(e.get_obj_coords(), 1))[0:3]
        assert np.abs(local - np.array(state[21:24])).max() < 1e-5
    e.close()
