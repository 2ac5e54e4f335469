"""The batched rollout loop of kinovagrasping_b200/train.py (main_DDPGfD.py:396-471 per env) on the host emulation:
which steps reach the replay buffer, the lift hand-over, success accounting."""
import numpy as np
import torch

from kinovagrasping_b200 import shard, train
from kinovagrasping_b200.replay import ReplayBuffer_Queue
from conftest import _emu_lib


def test_settled_matches_check_grasp_rule():
    prev = torch.zeros(3, 82); cur = torch.zeros(3, 82)
    cur[0, 9] = 0.001; cur[0, 12] = 0.001; cur[0, 15] = 0.0009       # 0.0029 / 15 = 1.93e-4 < 2e-4
    cur[1, 9] = 0.002; cur[1, 12] = 0.001                            # 0.003 / 15 = 2e-4: not below
    cur[2, 10] = 1.0                                                 # y of finger 1 is not part of the rule
    assert train.settled(prev, cur).tolist() == [True, False, True]


def test_policy_rollout_feeds_replay_and_hands_over_to_lift():
    n = 6
    env = shard.make_sharded_env(n, 0, 1, "cpu", lib=_emu_lib())
    buf = ReplayBuffer_Queue(82, 4, 64, 5, device="cpu", n_envs=n)
    calls = []

    def act(obs):                       # close the fingers at a constant rate, like the naive controller
        calls.append(obs.shape)
        a = torch.zeros(obs.shape[0], 4)
        a[:, 1:] = 0.6
        return a
    succ, stored = train.run_episode(env, act, buf, ("CubeS", "CubeB"), "normal", "B", seed=5)
    assert len(calls) == 30 and succ.shape == (n,)
    # every env either finished by TimeLimit or by a lift; stored policy steps only: at most 30 per env, at least 5 (SKIP_NUM_TS - 1)
    assert 5 * n <= stored <= 30 * n
    assert buf.replay_ep_num == n or buf.replay_ep_num < n        # short episodes are dropped, never more than n
    for i in range(buf.replay_ep_num):
        st, ac, ns, rw, nd = (x.numpy() for x in buf.episode(i))
        assert len(st) > 6 and np.all(ac[:, 0] == 0.0)            # no lifting step is stored: wrist action of policy steps is 0
        assert nd[-1] == 0.0 and np.all(nd[:-1] == 1.0)           # closed by a stored done or by replace
        assert np.array_equal(ns[:-1], st[1:])                    # consecutive transitions of one env
    assert int(buf.st_len.sum()) == 0                             # nothing left staged


def test_expert_rollout_fills_demonstrations():
    n = 4
    env = shard.make_sharded_env(n, 0, 1, "cpu", lib=_emu_lib())
    buf = ReplayBuffer_Queue(82, 4, 64, 5, device="cpu", n_envs=n)
    succ, stored = train.run_episode(env, None, buf, ("CubeS",), "normal", "B", seed=3, expert_mode="combined")
    assert stored > 0 and buf.replay_ep_num >= 1
    for i in range(buf.replay_ep_num):
        st, ac, ns, rw, nd = (x.numpy() for x in buf.episode(i))
        assert np.all(ac[:, 0] == 0.0) and np.all(ac[:, 1:] >= 0.0)
        assert (rw[-1] > 0) == bool(rw.sum() > 0)                  # a reward can only sit on the closing transition
