#!/usr/bin/env python
"""Generate tests/golden/learner_golden.npz by executing the REFERENCE's own learner code, unmodified, imported from
/root/reference/gym-kinova-gripper (DDPGfD.py, utils.py run on numpy + torch alone).  Runs only in the
authoring container; the fixture is committed so that nothing at test time needs /root/reference.

Pins:
  * Actor / Critic forward on a shipped 82-d checkpoint (policies/exp_NO_graspclassifier_expert_CubeS/..., DDPG:15-50)
  * ReplayBuffer_Queue: add / add_episode / replace / FIFO eviction, and sample_batch_nstep's episode + window draws
    for a seeded np.random (UTIL:34-107, 240-343)
  * three consecutive DDPGfD.train_batch updates (mixed agent/expert sampling, 1-step + n-step critic loss, actor loss
    over all n window states, Adam with weight decay, Polyak update every 10th iteration forced by total_it; DDPG:219-367):
    initial weights, the sampled batches, the 4 losses per update and the weights after each update.
"""
import glob
import os
import sys

import numpy as np
import torch

REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference/gym-kinova-gripper"
OUT = os.path.join(REPO, "tests", "golden", "learner_golden.npz")


def synth_episode(rng, length, lift):
    s = rng.normal(0, 0.3, (length + 1, 82)).astype(np.float32)
    a = rng.uniform(0, 0.8, (length, 4)).astype(np.float32)
    r = np.zeros(length, dtype=np.float32)
    if lift:
        r[-1] = 50.0
    return s, a, r


def fill(buf, rng, n_ep, lens=None):
    eps = []
    for e in range(n_ep):
        length = int(lens[e]) if lens is not None else int(rng.integers(8, 31))
        s, a, r = synth_episode(rng, length, lift=(e % 3 == 0))
        buf.add_episode(1)
        for t in range(length):
            buf.add(s[t], a[t], s[t + 1], float(r[t]), float(t == length - 1))
        buf.add_episode(0)
        eps.append((s, a, r))
    return eps


def sd_to_np(prefix, sd, out):
    for k, v in sd.items():
        out[f"{prefix}.{k}"] = v.detach().cpu().numpy().copy()


def main():
    sys.path.insert(0, REF)
    torch.manual_seed(7)
    np.random.seed(7)
    import DDPGfD as R
    import utils as U
    R.device = torch.device("cpu")
    out = {}

    # ---- checkpoint forward known answers
    ck = sorted(glob.glob(os.path.join(REF, "policies", "exp_NO_graspclassifier_expert_CubeS", "*_actor")))
    ck = [c for c in ck if not c.endswith("optimizer")][0]
    actor = R.Actor(82, 4, 0.8)
    actor.load_state_dict(torch.load(ck, map_location="cpu"))
    critic = R.Critic(82, 4)
    critic.load_state_dict(torch.load(ck[:-len("_actor")] + "_critic", map_location="cpu"))
    rng = np.random.default_rng(11)
    xs = rng.normal(0, 0.5, (64, 82)).astype(np.float32)
    xa = rng.uniform(0, 0.8, (64, 4)).astype(np.float32)
    with torch.no_grad():
        out["ck_state"] = xs; out["ck_action"] = xa
        out["ck_actor_out"] = actor(torch.from_numpy(xs)).numpy()
        out["ck_critic_out"] = critic(torch.from_numpy(xs), torch.from_numpy(xa)).numpy()
    sd_to_np("ck_actor", actor.state_dict(), out)
    sd_to_np("ck_critic", critic.state_dict(), out)

    # ---- replay: FIFO + replace + seeded sampling
    buf = U.ReplayBuffer_Queue(82, 4, max_episode=6, n_steps=5)
    buf.device = torch.device("cpu")
    lens = [12, 9, 30, 7, 16, 11, 20, 8]
    eps = fill(buf, rng, len(lens), lens)
    out["rb_lens_in"] = np.array(lens)
    out["rb_replay_ep_num"] = buf.replay_ep_num
    out["rb_kept_lens"] = np.array([len(buf.state[i]) for i in range(buf.replay_ep_num)])
    out["rb_first_state0"] = np.array(buf.state[0][0])
    # an episode that ends by `replace` (lifted-and-done, EXP:796-797 / MAIN:287)
    s, a, r = synth_episode(rng, 10, lift=False)
    buf.add_episode(1)
    for t in range(10):
        buf.add(s[t], a[t], s[t + 1], 0.0, 0.0)
    old = buf.replace(50.0, True)
    buf.add_episode(0)
    out["rb_replace_old"] = old
    out["rb_after_replace_ep_num"] = buf.replay_ep_num
    out["rb_last_reward_row"] = np.array(buf.reward[buf.replay_ep_num - 1])
    out["rb_last_notdone_row"] = np.array(buf.not_done[buf.replay_ep_num - 1])
    for i in range(buf.replay_ep_num):
        out[f"rb_ep{i}_state"] = np.array(buf.state[i]); out[f"rb_ep{i}_action"] = np.array(buf.action[i])
        out[f"rb_ep{i}_next"] = np.array(buf.next_state[i]); out[f"rb_ep{i}_reward"] = np.array(buf.reward[i])
        out[f"rb_ep{i}_notdone"] = np.array(buf.not_done[i])
    np.random.seed(123)
    st, ac, ns, rw, nd = buf.sample_batch_nstep(4)
    out["rb_sample_seed"] = 123
    out["rb_sample_state"] = st.numpy(); out["rb_sample_action"] = ac.numpy(); out["rb_sample_next"] = ns.numpy()
    out["rb_sample_reward"] = rw.numpy(); out["rb_sample_notdone"] = nd.numpy()

    # ---- three train_batch updates from a seeded initial policy, agent + expert buffers
    torch.manual_seed(3)
    pol = R.DDPGfD(82, 4, 0.8, 5, discount=0.995, tau=0.0005, batch_size=8)
    sd_to_np("init_actor", pol.actor.state_dict(), out)
    sd_to_np("init_critic", pol.critic.state_dict(), out)
    agent = U.ReplayBuffer_Queue(82, 4, max_episode=100, n_steps=5); agent.device = torch.device("cpu")
    expert = U.ReplayBuffer_Queue(82, 4, max_episode=100, n_steps=5); expert.device = torch.device("cpu")
    rng2 = np.random.default_rng(5)
    fill(agent, rng2, 12)
    fill(expert, rng2, 12)
    pol.total_it = 8          # so that the third update (total_it == 10) performs the Polyak target update
    for k in range(3):
        np.random.seed(1000 + k)
        # record the batch this update will draw: same seed, same call order as train_batch (agent first, then expert)
        ab = int(8 * (1 - 0.3)); eb = 8 - ab
        a_s, a_a, a_n, a_r, a_d = agent.sample_batch_nstep(ab)
        e_s, e_a, e_n, e_r, e_d = expert.sample_batch_nstep(eb)
        out[f"tb{k}_state"] = torch.cat((a_s, e_s), 0).numpy(); out[f"tb{k}_action"] = torch.cat((a_a, e_a), 0).numpy()
        out[f"tb{k}_next"] = torch.cat((a_n, e_n), 0).numpy(); out[f"tb{k}_reward"] = torch.cat((a_r, e_r), 0).numpy()
        out[f"tb{k}_notdone"] = torch.cat((a_d, e_d), 0).numpy()
        np.random.seed(1000 + k)
        losses = pol.train_batch(0, expert, agent, 0.3)     # MAIN:484-486 passes prob positionally into num_trajectories
        out[f"tb{k}_losses"] = np.array(losses, dtype=np.float64)
        if k == 2:      # the losses of updates 1 and 2 already pin the weights after updates 0 and 1
            sd_to_np(f"tb{k}_actor", pol.actor.state_dict(), out)
            sd_to_np(f"tb{k}_critic", pol.critic.state_dict(), out)
            sd_to_np(f"tb{k}_actor_target", pol.actor_target.state_dict(), out)
            sd_to_np(f"tb{k}_critic_target", pol.critic_target.state_dict(), out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB;", "train_batch losses", [out[f"tb{k}_losses"] for k in range(3)])


if __name__ == "__main__":
    main()
