"""Learner + replay rows against fixtures produced by the reference's OWN code
(tools/gen_golden_learner.py imports DDPGfD.py / utils.py unmodified): checkpoint forwards, replay FIFO / replace /
seeded n-step sampling, three consecutive train_batch updates.  fp32 on both sides; tolerances written per check.
The CPU run checks the host logic and the mathematics; the @gpu run repeats it on the device the product uses."""
import os

import numpy as np
import pytest
import torch

from kinovagrasping_b200.learner import DDPGfD, Actor, Critic
from kinovagrasping_b200.replay import ReplayBuffer_Queue

DEVICES = [pytest.param("cpu"), pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def G(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "learner_golden.npz")))


def _load_sd(module, G, prefix):
    sd = {k[len(prefix) + 1:]: torch.from_numpy(v) for k, v in G.items() if k.startswith(prefix + ".")}
    assert set(sd) == set(module.state_dict())
    module.load_state_dict(sd)


@pytest.mark.parametrize("device", DEVICES)
def test_actor_critic_forward_on_shipped_checkpoint(G, device):
    """DDPG:15-50 known answers on policies/exp_NO_graspclassifier_expert_CubeS (82-d actor, 86-d critic): 1e-5."""
    a, c = Actor(82, 4, 0.8), Critic(82, 4)
    _load_sd(a, G, "ck_actor"); _load_sd(c, G, "ck_critic")
    a, c = a.to(device), c.to(device)
    s, act = torch.from_numpy(G["ck_state"]).to(device), torch.from_numpy(G["ck_action"]).to(device)
    with torch.no_grad():
        assert np.abs(a(s).cpu().numpy() - G["ck_actor_out"]).max() < 1e-5
        assert np.abs(c(s, act).cpu().numpy() - G["ck_critic_out"]).max() < 1e-5 * max(1.0, np.abs(G["ck_critic_out"]).max())
    assert sum(p.numel() for p in a.parameters()) == 154704 and sum(p.numel() for p in c.parameters()) == 155401


def _synth_episode(rng, length, lift):
    s = rng.normal(0, 0.3, (length + 1, 82)).astype(np.float32)
    a = rng.uniform(0, 0.8, (length, 4)).astype(np.float32)
    r = np.zeros(length, dtype=np.float32)
    if lift:
        r[-1] = 50.0
    return s, a, r


def _fill(buf, rng, n_ep, lens=None):
    for e in range(n_ep):
        length = int(lens[e]) if lens is not None else int(rng.integers(8, 31))
        s, a, r = _synth_episode(rng, length, lift=(e % 3 == 0))
        buf.add_episode(1)
        for t in range(length):
            buf.add(s[t], a[t], s[t + 1], float(r[t]), float(t == length - 1))
        buf.add_episode(0)


@pytest.mark.parametrize("device", DEVICES)
def test_replay_fifo_replace_and_seeded_sampling(G, device):
    """UTIL:34-107, 240-343: the same add/replace sequence and the same np.random seed as the reference run give the
    same stored episodes and bit-identical sampled windows (values are copied, never recomputed: exact)."""
    rng = np.random.default_rng(11)
    rng.normal(0, 0.5, (64, 82)); rng.uniform(0, 0.8, (64, 4))      # the generator's draws before it filled the buffer
    buf = ReplayBuffer_Queue(82, 4, max_episode=6, n_steps=5, device=device)
    lens = list(G["rb_lens_in"])
    _fill(buf, rng, len(lens), lens)
    assert buf.replay_ep_num == int(G["rb_replay_ep_num"])
    assert [buf.episode_length(i) for i in range(buf.replay_ep_num)] == list(G["rb_kept_lens"])
    assert np.array_equal(buf.episode(0)[0][0].cpu().numpy(), G["rb_first_state0"])
    s, a, r = _synth_episode(rng, 10, lift=False)
    buf.add_episode(1)
    for t in range(10):
        buf.add(s[t], a[t], s[t + 1], 0.0, 0.0)
    old = buf.replace(50.0, True)
    buf.add_episode(0)
    assert old == float(G["rb_replace_old"]) and buf.replay_ep_num == int(G["rb_after_replace_ep_num"])
    last = buf.episode(buf.replay_ep_num - 1)
    assert np.array_equal(last[3].cpu().numpy(), G["rb_last_reward_row"]) and np.array_equal(last[4].cpu().numpy(), G["rb_last_notdone_row"])
    for i in range(buf.replay_ep_num):
        st, ac, ns, rw, nd = (x.cpu().numpy() for x in buf.episode(i))
        assert np.array_equal(st, G[f"rb_ep{i}_state"]) and np.array_equal(ac, G[f"rb_ep{i}_action"])
        assert np.array_equal(ns, G[f"rb_ep{i}_next"]) and np.array_equal(rw, G[f"rb_ep{i}_reward"]) and np.array_equal(nd, G[f"rb_ep{i}_notdone"])
    np.random.seed(int(G["rb_sample_seed"]))
    st, ac, ns, rw, nd = buf.sample_batch_nstep(4)
    assert st.shape == G["rb_sample_state"].shape and st.device.type == device
    assert np.array_equal(st.cpu().numpy(), G["rb_sample_state"]) and np.array_equal(ac.cpu().numpy(), G["rb_sample_action"])
    assert np.array_equal(ns.cpu().numpy(), G["rb_sample_next"]) and np.array_equal(rw.cpu().numpy(), G["rb_sample_reward"])
    assert np.array_equal(nd.cpu().numpy(), G["rb_sample_notdone"])
    with pytest.raises(ValueError):
        buf.replace(1.0, False)


@pytest.mark.parametrize("device", DEVICES)
def test_train_batch_three_updates_match_reference(G, device):
    """DDPG:219-367 on the batches the reference drew: 4 losses per update within 1e-4 relative, weights and Polyak
    targets after the third update (total_it == 10) within 2e-5 absolute."""
    pol = DDPGfD(82, 4, 0.8, 5, discount=0.995, tau=0.0005, batch_size=8, device=device, process_group=False)
    _load_sd(pol.actor, G, "init_actor"); _load_sd(pol.critic, G, "init_critic")
    pol.actor_target.load_state_dict(pol.actor.state_dict()); pol.critic_target.load_state_dict(pol.critic.state_dict())
    pol.total_it = 8
    for k in range(3):
        b = [torch.from_numpy(G[f"tb{k}_{n}"]).to(device) for n in ("state", "action", "next", "reward", "notdone")]
        losses = [float(x) for x in pol.train_on_batch(*b)]
        ref = G[f"tb{k}_losses"]
        assert np.allclose(losses, ref, rtol=1e-4, atol=1e-5), (k, losses, ref)
    for name, mod in (("tb2_actor", pol.actor), ("tb2_critic", pol.critic), ("tb2_actor_target", pol.actor_target), ("tb2_critic_target", pol.critic_target)):
        for kk, v in mod.state_dict().items():
            assert np.abs(v.cpu().numpy() - G[f"{name}.{kk}"]).max() < 2e-5, (name, kk)
    # the parameters really are views of one flat buffer (one all-reduce per backward)
    assert pol.actor_flat.numel() == 154704 and pol.critic_flat.numel() == 155401
    assert pol.actor.l1.weight.data_ptr() == pol.actor_flat.data_ptr()


@pytest.mark.parametrize("device", DEVICES)
def test_train_batch_facade_samples_like_reference(G, device):
    """The reference-signature train_batch: agent batch first, then expert, same np.random call order (DDPG:235-241)."""
    pol = DDPGfD(82, 4, 0.8, 5, batch_size=8, device=device, process_group=False)
    _load_sd(pol.actor, G, "init_actor"); _load_sd(pol.critic, G, "init_critic")
    pol.actor_target.load_state_dict(pol.actor.state_dict()); pol.critic_target.load_state_dict(pol.critic.state_dict())
    pol.total_it = 8
    agent = ReplayBuffer_Queue(82, 4, max_episode=100, n_steps=5, device=device)
    expert = ReplayBuffer_Queue(82, 4, max_episode=100, n_steps=5, device=device)
    rng2 = np.random.default_rng(5)
    _fill(agent, rng2, 12); _fill(expert, rng2, 12)
    np.random.seed(1000)
    losses = pol.train_batch(0, expert, agent, 0.3)
    assert np.allclose(losses, G["tb0_losses"], rtol=1e-4, atol=1e-5), (losses, G["tb0_losses"])


@pytest.mark.parametrize("device", DEVICES)
def test_vectorised_feeder_and_device_sampler(device):
    """add_batch: record / replace / drop-short-episode semantics (MAIN:451-471, EXP:794-797) and FIFO over the ring;
    sample_windows returns well-formed windows of stored episodes only."""
    N, T = 6, 12
    buf = ReplayBuffer_Queue(82, 4, max_episode=8, n_steps=5, max_episode_len=T, device=device, n_envs=N)
    g = torch.Generator(device=device); g.manual_seed(0)
    lens = torch.tensor([9, 12, 6, 10, 8, 12], device=device)          # env 2: 6 - 5 <= 1 -> dropped
    committed = 0
    for t in range(T):
        s = torch.full((N, 82), float(t), device=device) + torch.arange(N, device=device)[:, None] * 100
        a = torch.rand((N, 4), device=device, generator=g)
        alive = t < lens
        done = (t == lens - 1)
        rew = torch.where(done, torch.tensor(50.0, device=device), torch.tensor(0.0, device=device))
        # env 3 stops recording after step 6 and is closed by `replace` at its last step
        record = alive & ~((torch.arange(N, device=device) == 3) & (t > 6))
        replace = (torch.arange(N, device=device) == 3) & done
        committed += buf.add_batch(s, a, s + 1, rew, done & record | replace, record=record, replace=replace)
    assert committed == 5 and buf.replay_ep_num == 5
    got = sorted(buf.episode_length(i) for i in range(5))
    assert got == [7, 8, 9, 12, 12]
    for i in range(5):
        st, ac, ns, rw, nd = buf.episode(i)
        assert float(rw[-1]) == 50.0 and float(nd[-1]) == 0.0 and float(nd[:-1].min()) == 1.0
        assert torch.equal(ns, st + 1) and torch.equal(st[1:, 0] - st[:-1, 0], torch.ones(len(st) - 1, device=device))
    st, ac, ns, rw, nd = buf.sample_windows(64, generator=g)
    assert st.shape == (64, 5, 82) and rw.shape == (64, 5)
    assert torch.equal(st[:, 1:, 0] - st[:, :-1, 0], torch.ones((64, 4), device=device))      # consecutive steps of one episode
    # FIFO: 6 more episodes through a ring of 8 keep at most max_episode - 1 complete episodes
    for rep in range(2):
        for t in range(10):
            s = torch.zeros((N, 82), device=device) + 1000 * (rep + 1) + t
            d = torch.full((N,), t == 9, device=device)
            buf.add_batch(s, torch.zeros((N, 4), device=device), s, torch.zeros(N, device=device), d)
    assert buf.replay_ep_num == 7
    assert float(buf.episode(6)[0][0, 0]) == 2000.0 and float(buf.episode(0)[0][0, 0]) >= 1000.0


def test_device_sampler_follows_the_ring_bookkeeping():
    """sample_windows reads head / episode count from a device tensor (so that a captured CUDA graph keeps working as the
    ring moves): after evictions the sampler only returns windows of episodes that are still held."""
    from kinovagrasping_b200.replay import ReplayBuffer_Queue
    n, T = 4, 12
    buf = ReplayBuffer_Queue(3, 1, max_episode=6, n_steps=2, max_episode_len=T, device="cpu", n_envs=n)
    tag = 0
    for sweep in range(5):                                   # 20 episodes through a ring of 5 complete ones
        for t in range(8):
            s = torch.full((n, 3), float(tag)) + torch.arange(n)[:, None] * 0.001
            done = torch.full((n,), t == 7)
            buf.add_batch(s, torch.zeros(n, 1), s + 0.5, torch.zeros(n), done)
        tag += 1
        meta = buf.publish_meta()
        assert int(meta[0]) == buf.head and int(meta[1]) == buf.replay_ep_num - 1
        st = buf.sample_windows(64)[0]
        held = {float(buf.episode(i)[0][0, 0].floor()) for i in range(buf.replay_ep_num - 1)}
        assert set(st[:, 0, 0].floor().tolist()) <= held
    assert buf.replay_ep_num <= 5
