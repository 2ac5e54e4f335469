"""Multi-GPU plumbing for the rollout path: environments are independent (kinova_gripper_env.py has no cross-env state),
so they shard by contiguous global index range, one process per GPU, with NO data-path collective.
torch.distributed is used only for the rendezvous, for max-over-ranks timing and for scalar success counters."""
import os

import torch


def dist_env():
    """(rank, local_rank, world) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_total, rank, world):
    """Contiguous global env ids [lo, hi) of `rank`; the first n_total % world ranks get one extra env."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def make_sharded_env(n_total, rank, world, device, lib=None, **kw):
    """This rank's VecKinovaGripperEnv over its slice of the global batch.  The pose sampler's counter-based RNG is
    keyed by the GLOBAL env id (kg_set_env_offset), so a run's draws do not depend on the number of GPUs."""
    from .env import VecKinovaGripperEnv
    lo, hi = shard_range(n_total, rank, world)
    env = VecKinovaGripperEnv(hi - lo, device=device, lib=lib, **kw)
    env.h.set_env_offset(lo)
    env.env_offset = lo
    env.global_range = (lo, hi)
    return env


def all_reduce_max(values, device, group=None):
    """Max over ranks of a list of floats (device timings are always reported as the slowest rank's)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.tolist()]


def all_reduce_sum(values, device, group=None):
    """Sum over ranks (episode / success counters at eval, main_DDPGfD.py:130-272)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.tolist()]


def all_reduce_min(values, device, group=None):
    """Min over ranks: collective yes/no decisions (every rank must take the same branch around a collective)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return [float(x) for x in t.tolist()]
