"""N>1 paths on CPU with world_size-2 gloo: (1) env sharding -- two ranks stepping halves of a batch
reproduce the single-process batch bit for bit (pose RNG keyed by global env id, no data-path collective);
(2) the learner's data-parallel update -- one flat all-reduce per backward, equal to the single-process update on the
concatenated batch.  The env side runs on the TEST-ONLY host emulation of the kernel source."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _rollout(env, steps):
    """Expert rollout with the grasp-classifier reward on; every record carries obs, reward (lift + grasp), done, the expert's
    action, the grasp part of the reward and -- in the first record -- the hand orientation the sampler drew per env."""
    env.set_with_grasp_reward(True)
    obs0 = env.reset(("CubeS", "CubeB", "Cone1S", "LemonM"), "random", variant="B", seed=77).clone()
    out = [torch.cat([obs0, torch.from_numpy(env.h.orientation_host().astype(np.float32))[:, None]], 1)]
    for _ in range(steps):
        obs, rew, done, info = env.step(None, "combined")
        out.append(torch.cat([obs, rew[:, None], done.float()[:, None], info["action"], info["grasp_reward"][:, None]], 1).clone())
        env.reset(("CubeS", "CubeB", "Cone1S", "LemonM"), "random", variant="B", seed=77, mask=env.done)
    return out


def _env_worker(rank, world, port, n_total, steps, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import _emu_lib
    from kinovagrasping_b200 import shard
    _init(rank, world, port)
    env = shard.make_sharded_env(n_total, rank, world, "cpu", lib=_emu_lib())
    recs = _rollout(env, steps)
    lo, hi = env.global_range
    t_ms = shard.all_reduce_max([1.0 + rank], "cpu")[0]                  # the timing reduction bench.py uses
    nsucc = shard.all_reduce_sum([float(hi - lo)], "cpu")[0]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, [r.numpy() for r in recs]))
    if rank == 0:
        q.put((gathered, t_ms, nsucc))
    dist.barrier()
    dist.destroy_process_group()


def test_env_sharding_two_ranks_equals_single_process():
    from conftest import _emu_lib
    from kinovagrasping_b200 import shard
    n_total, steps = 10, 14
    assert shard.shard_range(10, 0, 3) == (0, 4) and shard.shard_range(10, 2, 3) == (7, 10)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_env_worker, args=(r, 2, port, n_total, steps, q)) for r in range(2)]
    for p in procs: p.start()
    gathered, t_ms, nsum = q.get()
    for p in procs:
        p.join(120); assert p.exitcode == 0
    assert t_ms == 2.0 and nsum == float(n_total)
    single = _rollout(shard.make_sharded_env(n_total, 0, 1, "cpu", lib=_emu_lib()), steps)
    for lo, hi, recs in gathered:
        for a, b in zip(recs, single):
            assert np.array_equal(a, b[lo:hi].numpy()), (lo, hi)
    assert len(set(single[0][:, -1].tolist())) >= 2                       # 'random' orientation: more than one value drawn ...
    assert sum(float(r[:, -1].sum()) for r in single[1:]) >= 5.0          # ... and the classifier paid at least once


def _learner_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from kinovagrasping_b200.learner import DDPGfD
    _init(rank, world, port)
    torch.manual_seed(0)
    pol = DDPGfD(82, 4, 0.8, 5, batch_size=8, device="cpu")
    assert pol.world == 2
    g = torch.Generator().manual_seed(1)
    B = 16
    batch = [torch.randn(B, 5, 82, generator=g), torch.rand(B, 5, 4, generator=g), torch.randn(B, 5, 82, generator=g),
             torch.rand(B, 5, generator=g), torch.ones(B, 5)]
    half = [x[rank * B // 2:(rank + 1) * B // 2] for x in batch]
    for _ in range(2):
        pol.train_on_batch(*half)
    flat = torch.cat([pol.actor_flat, pol.critic_flat])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        q.put([b.numpy() for b in both])
    dist.barrier()
    dist.destroy_process_group()


def test_learner_data_parallel_equals_full_batch():
    from kinovagrasping_b200.learner import DDPGfD
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    both = q.get()
    for p in procs:
        p.join(120); assert p.exitcode == 0
    assert np.array_equal(both[0], both[1])                      # replicas stay in lock step
    torch.manual_seed(0)
    pol = DDPGfD(82, 4, 0.8, 5, batch_size=8, device="cpu", process_group=False)
    g = torch.Generator().manual_seed(1)
    B = 16
    batch = [torch.randn(B, 5, 82, generator=g), torch.rand(B, 5, 4, generator=g), torch.randn(B, 5, 82, generator=g),
             torch.rand(B, 5, generator=g), torch.ones(B, 5)]
    for _ in range(2):
        pol.train_on_batch(*batch)
    flat = torch.cat([pol.actor_flat, pol.critic_flat]).numpy()
    # mean-of-half-batch gradients == full-batch gradient up to fp32 summation order
    assert np.abs(flat - both[0]).max() < 2e-6, np.abs(flat - both[0]).max()
