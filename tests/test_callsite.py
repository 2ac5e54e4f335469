"""Call-site integration (north_star "main_DDPGfD.py / DDPGfD.py and the expert PID controller unchanged at the
call site"): the reference's OWN rollout loops -- expert_data.GenerateExpertPID_JointVel (EXP:690-921) and
main_DDPGfD.update_policy / eval_policy (MAIN:130-272, 333-537), imported unmodified from /root/reference -- run against
kinovagrasping_b200.env.make in place of gym.make, filling this repo's replay ring and training this repo's learner.

The same loops are then run a second time, seeded identically, on the "oracle-driven" stack: the same facade on the fp64
oracle, the reference's own utils.ReplayBuffer_Queue and DDPGfD.  Episode structure must be identical and the stored
transitions agree to the accuracy a free-running fp32-vs-fp64 rollout allows (tolerances at the asserts).

/root/reference exists only in the authoring container, so these are CPU tests on the host emulation of the kernels
(they skip on the GPU box); what the CUDA library adds over the emulation is covered by tests/test_parity.py."""
import argparse
import os

import numpy as np
import pytest
import torch

import _refharness as RH
from conftest import _emu_lib
from kinovagrasping_b200 import env as kenv
from kinovagrasping_b200.learner import DDPGfD
from kinovagrasping_b200.replay import ReplayBuffer_Queue

pytestmark = pytest.mark.skipif(not RH.available(), reason="/root/reference is not mounted here")


def _fresh_files():
    kenv._MEM_FILES.clear()
    kenv._XML_EULER.update({k: [-1.57, 0.0, -1.57] for k in kenv._XML_EULER})
    kenv._XML_EULER["CubeS"] = [-0.06, -0.16, -0.13]


def _episodes(buf, ours):
    out = []
    for i in range(buf.replay_ep_num):
        if ours:
            out.append(tuple(x.cpu().numpy() for x in buf.episode(i)))
        else:
            out.append((np.array(buf.state[i], dtype=np.float32), np.array(buf.action[i], dtype=np.float32), np.array(buf.next_state[i], dtype=np.float32),
                        np.array(buf.reward[i], dtype=np.float32), np.array(buf.not_done[i], dtype=np.float32)))
    return out


def _run_expert(tmp_path, product, n_ep, shapes, orientation, pid_mode):
    _fresh_files()
    lib = _emu_lib()
    RH.install_stubs((lambda: kenv.make(lib=lib)) if product else (lambda: kenv.KinovaGripper_Env(handle=RH.OracleHandle())))
    ED = RH.import_reference("expert_data")
    if product:
        buf = ReplayBuffer_Queue(82, 4, max_episode=100, n_steps=5, device="cpu")
    else:
        buf = RH.import_reference("utils").ReplayBuffer_Queue(82, 4, 100, 5)
        buf.save_replay_buffer = lambda d: d      # the reference's np.save of ragged lists needs numpy < 1.24; not what is compared here
    cwd = os.getcwd()
    os.makedirs(tmp_path, exist_ok=True)
    os.chdir(tmp_path)
    try:
        np.random.seed(7)
        with RH.quiet():
            ret = ED.GenerateExpertPID_JointVel(n_ep, shapes, orientation, False, buf, save=True, render_imgs=False, pid_mode=pid_mode)
    finally:
        os.chdir(cwd)
    return buf, ret


def test_generate_expert_pid_unmodified(tmp_path):
    shapes, n_ep = ["CubeS", "CubeB", "Cone1S"], 4
    ours, ret = _run_expert(tmp_path / "a", True, n_ep, shapes, "random", "combined")
    ref, ret_ref = _run_expert(tmp_path / "b", False, n_ep, shapes, "random", "combined")
    # same bookkeeping: episodes generated / held, per-episode start-end marks, orientation rows, success counts
    assert ours.replay_ep_num == ref.replay_ep_num and ours.size == ref.size
    assert [list(e) if isinstance(e, list) else e for e in ours.episodes] == [list(e) if isinstance(e, list) else e for e in ref.episodes]
    assert ours.orientation_indexes == ref.orientation_indexes
    assert ret[4:6] == ret_ref[4:6] and ret[6] == ret_ref[6]          # num_success, total, coords file path
    a, b = _episodes(ours, True), _episodes(ref, False)
    assert len(a) == len(b) > 0
    for ea, eb in zip(a, b):
        assert ea[0].shape == eb[0].shape
        assert np.array_equal(ea[1], eb[1].reshape(ea[1].shape)) or np.abs(ea[1] - eb[1].reshape(ea[1].shape)).max() < 2e-3   # controller outputs
        assert np.array_equal(ea[3], eb[3]) and np.array_equal(ea[4], eb[4])                                                 # rewards / not_done exact
        assert np.abs(ea[0][0] - eb[0][0]).max() < 1e-5                       # first observation of the episode: reset parity
        assert np.abs(ea[2] - eb[2]).max() < 5e-3                             # free-running fp32 vs fp64 over <= 30 RL steps
    # the files the generator saved load back through the REFERENCE's loader
    U = RH.import_reference("utils")
    back = U.ReplayBuffer_Queue(82, 4, 100, 5)
    with RH.quiet():
        back.store_saved_data_into_replay(str(tmp_path / "a") + "/" + ret[1])      # ret[1] ends with "/", as the loader needs
    assert back.replay_ep_num == ours.replay_ep_num - 1       # the loader pops the trailing empty entry with remove_episode(-1)
    for i in range(ours.replay_ep_num):
        assert np.array_equal(np.array(back.state[i], dtype=np.float32), a[i][0]) and np.array_equal(np.array(back.reward[i], dtype=np.float32), a[i][3])


def _run_update_policy(tmp_path, product, expert_dir):
    _fresh_files()
    lib = _emu_lib()
    RH.install_stubs((lambda: kenv.make(lib=lib)) if product else (lambda: kenv.KinovaGripper_Env(handle=RH.OracleHandle())))
    M = RH.import_reference("main_DDPGfD")
    U = RH.import_reference("utils")
    R = RH.import_reference("DDPGfD")
    R.device = torch.device("cpu"); U.device = torch.device("cpu")
    torch.manual_seed(3)
    ref_policy = R.DDPGfD(82, 4, 0.8, 5, batch_size=4)
    if product:
        policy = DDPGfD(82, 4, 0.8, 5, batch_size=4, device="cpu", process_group=False)
        for mine, theirs in ((policy.actor, ref_policy.actor), (policy.critic, ref_policy.critic),
                             (policy.actor_target, ref_policy.actor_target), (policy.critic_target, ref_policy.critic_target)):
            mine.load_state_dict(theirs.state_dict())         # same initial weights as the oracle-driven run
        agent = ReplayBuffer_Queue(82, 4, max_episode=50, n_steps=5, device="cpu")
        expert = ReplayBuffer_Queue(82, 4, max_episode=50, n_steps=5, device="cpu")
    else:
        policy = ref_policy
        agent = U.ReplayBuffer_Queue(82, 4, 50, 5)
        expert = U.ReplayBuffer_Queue(82, 4, 50, 5)
    with RH.quiet():
        expert.store_saved_data_into_replay(expert_dir)
    # a few valid agent episodes so that the first in-loop update has something to sample (MAIN's update_after = 100 does that)
    rng = np.random.default_rng(5)
    for e in range(3):
        agent.add_episode(1)
        s = rng.normal(0, 0.3, (13, 82)).astype(np.float32)
        for t in range(12):
            agent.add(s[t], rng.uniform(0, 0.8, 4).astype(np.float32), s[t + 1], 0.0, float(t == 11))
        agent.add_episode(0)
    M.args = argparse.Namespace(mode="train", env_name="gym_kinova_gripper:kinovagripper-v0", with_grasp_reward=False, expl_noise=0.1,
                                update_after=0, update_num=2, batch_size=4, eval_freq=1000, eval_num=2, save_freq=1000, seed=2, max_episode=20000)
    M.policy, M.replay_buffer, M.expert_replay_buffer = policy, agent, expert
    M.requested_shapes, M.requested_orientation = ["CubeS", "CubeB"], "normal"
    M.max_action, M.action_dim = 0.8, 4
    M.wrist_lift_velocity, M.finger_lift_velocity = 0.6, 0.5
    M.all_saving_dirs = {"tensorboard_dir": str(tmp_path / "tb")}
    M.env = kenv.KinovaGripper_Env(handle=RH.OracleHandle())      # the module-level env eval_policy reads the orientation from (MAIN:254)
    cwd = os.getcwd()
    os.makedirs(tmp_path, exist_ok=True)
    os.chdir(tmp_path)
    try:
        np.random.seed(11)
        with RH.quiet():
            out = M.update_policy([], 1, 4, 0.3, "train", "run", 30)      # episodes 1..4: updates from the first one on, eval at episode 3
    finally:
        os.chdir(cwd)
    return policy, agent, out


def test_update_policy_and_eval_policy_unmodified(tmp_path):
    """MAIN:333-537 for four episodes with two in-loop train_batch calls per episode and the closing eval_policy (2 grasp
    trials): an untrained actor barely moves the fingers, check_grasp fires at the 6th step, the episode is too short and
    main_DDPGfD.py:470-471 calls remove_episode(-1) -- the path ADVICE r1 found crashing."""
    _, ret = _run_expert(tmp_path / "e", True, 6, ["CubeS", "CubeB"], "normal", "naive")
    expert_dir = str(tmp_path / "e") + "/" + ret[1]
    pol_a, ag_a, out_a = _run_update_policy(tmp_path / "a", True, expert_dir)
    pol_b, ag_b, out_b = _run_update_policy(tmp_path / "b", False, expert_dir)
    assert out_a[1] == out_b[1] == 5 and out_a[2:] == out_b[2:]
    assert ag_a.replay_ep_num == ag_b.replay_ep_num and ag_a.size == ag_b.size
    a, b = _episodes(ag_a, True), _episodes(ag_b, False)
    for ea, eb in zip(a, b):
        assert ea[0].shape == eb[0].shape and np.array_equal(ea[3], eb[3]) and np.array_equal(ea[4], eb[4])
        assert np.abs(ea[0] - eb[0]).max() < 5e-3 and np.abs(ea[1] - eb[1].reshape(ea[1].shape)).max() < 5e-3
    # the episode in progress after the last remove_episode(-1) (the quirk: the short episode stays open)
    la = int(ag_a.lengths[ag_a._row(ag_a.replay_ep_num)]); lb = len(ag_b.state[ag_b.replay_ep_num])
    assert la == lb
    # both learners took the same 8 updates from near-identical batches
    assert pol_a.total_it == pol_b.total_it == 8
    for pa, pb in zip(pol_a.actor.parameters(), pol_b.actor.parameters()):
        assert np.abs(pa.detach().numpy() - pb.detach().numpy()).max() < 2e-3


def test_checkpoint_files_interchange_with_the_reference(tmp_path):
    """DDPG:371-382: the four torch.save files `<name>_{actor,critic,actor_optimizer,critic_optimizer}`.  Reference save ->
    this repo's load -> identical forward and an identical next update; this repo's save -> reference load -> identical
    forward (the policy-stage hand-over of MAIN:776-797 globs for `*_actor_optimizer`)."""
    RH.install_stubs(lambda: None)
    R = RH.import_reference("DDPGfD")
    R.device = torch.device("cpu")
    torch.manual_seed(1)
    ref = R.DDPGfD(82, 4, 0.8, 5, batch_size=8)
    rng = np.random.default_rng(0)
    batch = [torch.from_numpy(rng.normal(0, 0.3, (24, 5, 82)).astype(np.float32)), torch.from_numpy(rng.uniform(0, 0.8, (24, 5, 4)).astype(np.float32)),
             torch.from_numpy(rng.normal(0, 0.3, (24, 5, 82)).astype(np.float32)), torch.from_numpy(rng.uniform(0, 1, (24, 5)).astype(np.float32)),
             torch.ones(24, 5)]
    mine = DDPGfD(82, 4, 0.8, 5, batch_size=8, device="cpu", process_group=False)
    for a, b in ((mine.actor_target, ref.actor_target), (mine.critic_target, ref.critic_target)):
        a.load_state_dict(b.state_dict())
    for _ in range(2):                                    # give Adam some state worth saving
        mine.actor.load_state_dict(ref.actor.state_dict()); mine.critic.load_state_dict(ref.critic.state_dict())
        break
    ref.save(str(tmp_path / "ref_policy"))
    assert sorted(os.listdir(tmp_path)) == ["ref_policy_actor", "ref_policy_actor_optimizer", "ref_policy_critic", "ref_policy_critic_optimizer"]
    mine.load(str(tmp_path / "ref_policy"))
    s = batch[0][:, 0]
    with torch.no_grad():
        assert torch.equal(mine.actor(s), ref.actor(s)) and torch.equal(mine.critic(s, batch[1][:, 0]), ref.critic(s, batch[1][:, 0]))
    assert mine.actor.l1.weight.data_ptr() == mine.actor_flat.data_ptr()      # load() kept the parameters inside the flat bucket
    losses = mine.train_on_batch(*batch)
    mine.save(str(tmp_path / "my_policy"))
    back = R.DDPGfD(82, 4, 0.8, 5, batch_size=8)
    back.load(str(tmp_path / "my_policy"))
    with torch.no_grad():
        assert torch.equal(back.actor(s), mine.actor(s)) and torch.equal(back.critic(s, batch[1][:, 0]), mine.critic(s, batch[1][:, 0]))
    assert back.actor_optimizer.state_dict()["state"][0]["step"] == mine.actor_optimizer.state_dict()["state"][0]["step"]
