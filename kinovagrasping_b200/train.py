"""Batched DDPGfD training over the CUDA env -- the `--mode train` flow of main_DDPGfD.py (MAIN:333-537, 602-621) with
N environments stepping in lock step instead of one.

  python -m kinovagrasping_b200.train --envs 4096 --sweeps 20 --updates 100 [--shapes CubeS,CubeB] [--orientation normal]
  torchrun --nproc-per-node 8 -m kinovagrasping_b200.train ...        (envs and replay sharded per rank, gradients all-reduced)

Kept from the reference, per environment:
  * the policy acts (plus Gaussian exploration noise, clipped to [0, max_action], MAIN:443-446) until `check_grasp`
    (EXP:559-593: summed distal x-displacement / 15 < 2e-4, tested from the 6th step on) says the fingers have settled; from
    then on the scripted lift action [0.6, 0.5, 0.5, 0.5] runs (MAIN:275-290);
  * only policy steps are stored; a lifted-and-done step rewrites the last stored reward (`replace`, UTIL:309-343); episodes
    with len - n_steps <= 1 are dropped (MAIN:470-471);
  * expert demonstrations come from the PID controller (on the device here), lifting steps not stored (EXP:794-797);
    the policy is first pre-trained on them alone (an untrained actor barely moves the fingers, `check_grasp` then fires at
    the 6th step and the episode is too short to keep -- the reference runs its pre-train stage first for the same reason);
  * `update_num` calls of the n-step update per sweep, each batch 1 - prob agent windows and prob expert windows
    (DDPG:219-262); evaluation = noise-free episodes on the test pose tables, success = lift reward (MAIN:130-272).
What changes: one sweep = one episode of every env (N episodes), so `update_after`/`eval_freq` count sweeps."""
import argparse
import json
import os
import time

import torch

from .learner import DDPGfD
from .replay import ReplayBuffer_Queue
from .shard import all_reduce_min, all_reduce_sum, dist_env, make_sharded_env

WRIST_LIFT_VELOCITY, FINGER_LIFT_VELOCITY = 0.6, 0.5      # MAIN:945-947
SKIP_NUM_TS = 6                                           # MAIN:417
GRASP_SETTLED = 0.0002                                    # EXP:590
SAMPLING_TIME = 15.0                                      # EXP:570


def settled(prev_obs, obs):
    """check_grasp (EXP:559-593) on [N,82] observations: obs[9:17] are the distal finger-tip coordinates; the reference
    sums the change of entries 0, 3, 6 of that slice."""
    d = (prev_obs[:, 9:17] - obs[:, 9:17]).abs()
    return (d[:, 0] + d[:, 3] + d[:, 6]) / SAMPLING_TIME < GRASP_SETTLED


def run_episode(env, act_fn, replay=None, shapes=("CubeS",), orientation="normal", variant="A", mode="train", seed=0,
                expert_mode=None):
    """One lock-step episode of every env.  act_fn(obs) -> [N,4] policy actions; expert_mode: let the on-device PID
    controller act instead (demonstrations).  Returns (success [N] bool, steps stored)."""
    t = torch
    obs = env.reset(shapes, orientation, mode=mode, variant=variant, seed=seed).clone()
    n, dev = env.n, env.device
    alive = t.ones(n, dtype=t.bool, device=dev)
    ready = t.zeros(n, dtype=t.bool, device=dev)
    success = t.zeros(n, dtype=t.bool, device=dev)
    lift = t.tensor([WRIST_LIFT_VELOCITY, FINGER_LIFT_VELOCITY, FINGER_LIFT_VELOCITY, FINGER_LIFT_VELOCITY], device=dev)
    prev = None
    stored = 0
    for step in range(1, env._max_episode_steps + 1):
        if expert_mode is None:
            if prev is not None and step >= SKIP_NUM_TS:
                ready |= settled(prev, obs)
            action = t.where(ready[:, None], lift[None, :], act_fn(obs))
            nxt, rew, done, info = env.step(action)
            lifting = ready
        else:
            nxt, rew, done, info = env.step(None, expert_mode)
            action = info["action"]
            lifting = action[:, 0] > 0                       # the controller's own lift phase
        nxt = nxt.clone()
        if replay is not None:
            rec = alive & ~lifting
            replay.add_batch(obs, action, nxt, rew, done & alive, record=rec, replace=alive & lifting & done, defer=True)
            stored = stored + rec.sum()
        success |= alive & done & (rew >= 50.0)          # the lift reward (a grasp-classifier +5 alone is not a success)
        alive &= ~done
        prev, obs = obs, nxt
    if replay is not None:
        replay.abandon(alive)
        replay.flush()              # the only host synchronisation of the sweep
        stored = int(stored)
    return success, stored


class GraphedUpdates:
    """`network_repl_freq` (10) consecutive updates -- device-side window sampling from both buffers, the n-step update, the
    target-network mix on the last one -- captured once into a CUDA graph and replayed: an update is ~60 small kernels and
    was launch-bound (3.3 ms each eagerly at batch 4096).  Under torchrun the two NCCL all-reduces of an update (critic
    half, then actor half of the flat gradient bucket; NCCL collectives are capturable) are nodes of the same graph, so the
    data-parallel update replays exactly like the single-GPU one plus two NVLink all-reduces of ~0.6 MB.
    Every rank must call run() the same number of times (train.main decides that collectively).  If a replay buffer
    re-allocates its storage (storage_version), the graph is re-captured."""

    def __init__(self, policy, agent, expert, batch, expert_prob):
        self.policy, self.agent, self.expert = policy, agent, expert
        self.n_agent = int(batch * (1 - expert_prob))
        self.n_expert = batch - self.n_agent
        self.block = policy.network_repl_freq
        self.graph = None
        self.losses = None
        self.versions = None

    def _one(self):
        if self.n_agent:
            a = self.agent.sample_windows(self.n_agent)
            e = self.expert.sample_windows(self.n_expert)
            batch = tuple(torch.cat((x, y), 0) for x, y in zip(a, e))
        else:
            batch = self.expert.sample_windows(self.n_expert)
        return self.policy.train_on_batch(*batch)

    def run(self, n_updates):
        """n_updates rounded up to a multiple of the block; returns the losses of the last update (device tensors)."""
        self.agent.publish_meta(); self.expert.publish_meta()
        versions = (self.agent.storage_version, self.expert.storage_version)
        if self.graph is not None and versions != self.versions:
            self.graph = None                                # the captured addresses are gone
        self.versions = versions
        if self.graph is None:
            assert self.policy.total_it % self.block == 0
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                    # eager warm-up block: optimizer state, cuBLAS workspaces
                for _ in range(self.block):
                    self._one()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                for _ in range(self.block):
                    self.losses = self._one()
            self.policy.total_it -= self.block               # capturing records the block, it does not run it
            n_updates -= self.block                          # the warm-up block did run
        for _ in range((max(n_updates, 0) + self.block - 1) // self.block):
            self.graph.replay()
            self.policy.total_it += self.block
        return self.losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="global number of environments")
    ap.add_argument("--shapes", default="CubeS,CubeB")
    ap.add_argument("--orientation", default="normal")
    ap.add_argument("--variant", default="B", choices=["A", "B"])
    ap.add_argument("--sweeps", type=int, default=20)
    ap.add_argument("--expert-sweeps", type=int, default=1)
    ap.add_argument("--updates", type=int, default=100, help="update_num per sweep (MAIN:910)")
    ap.add_argument("--pretrain-updates", type=int, default=1000,
                    help="expert-only updates before the first rollout (the reference's --mode pre-train stage, MAIN:1130-1160)")
    ap.add_argument("--batch", type=int, default=4096, help="n-step windows per update and rank")
    ap.add_argument("--expert-prob", type=float, default=0.3)
    ap.add_argument("--expl-noise", type=float, default=0.1)
    ap.add_argument("--eval-every", type=int, default=5)
    ap.add_argument("--max-episodes", type=int, default=40000, help="replay capacity in episodes per rank")
    ap.add_argument("--tau", type=float, default=0.0005, help="target-network mixing (DDPG:221 default)")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="run the updates eagerly instead of replaying captured CUDA graphs")
    ap.add_argument("--fp32-matmul", action="store_true", help="keep the actor/critic GEMMs in IEEE fp32 instead of TF32 tensor cores")
    ap.add_argument("--save", default=None)
    ap.add_argument("--eval-out", default=None, help="directory for the reference's eval outputs (heatmap coordinates, boxplot arrays) of every evaluation")
    ap.add_argument("--with-grasp-reward", action="store_true", help="add the grasp classifier's +5 (MAIN:906 --with_grasp_reward, ENV:643-659)")
    args = ap.parse_args()

    rank, local, world = dist_env()
    assert torch.cuda.is_available(), "training needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    if not args.fp32_matmul:
        torch.set_float32_matmul_precision("high")            # TF32 tensor cores for the 20k x 400 x 300 GEMMs of an update
    torch.manual_seed(args.seed)                              # same initial weights on every rank
    env = make_sharded_env(args.envs, rank, world, local)
    if args.with_grasp_reward:
        env.set_with_grasp_reward(True)
    shapes = tuple(args.shapes.split(","))
    max_action = 0.8                                          # action_space.high (ENV:128)
    graphed = not args.no_graph
    policy = DDPGfD(82, 4, max_action, n=5, tau=args.tau, batch_size=args.batch, device=dev, capturable=graphed)
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed * 1000 + rank)
    expert = ReplayBuffer_Queue(82, 4, args.max_episodes, 5, device=dev, n_envs=env.n)
    agent = ReplayBuffer_Queue(82, 4, args.max_episodes, 5, device=dev, n_envs=env.n)

    def log(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    def count(mask):
        s, n = all_reduce_sum([float(mask.sum()), float(mask.numel())], dev)
        return s / n

    t0 = time.perf_counter()
    for k in range(args.expert_sweeps):
        succ, stored = run_episode(env, None, expert, shapes, args.orientation, args.variant, seed=args.seed + 7000 + k, expert_mode="combined")
        log(phase="expert", sweep=k, success=count(succ), episodes_held=expert.replay_ep_num, steps_stored=stored)

    held = all_reduce_min([float(expert.replay_ep_num)], dev)[0]
    assert held > 2, f"a rank holds only {held:.0f} expert episodes: pre-training would sample an empty shard"

    def evaluate(tag, sweep):
        if args.eval_out:      # the reference's eval outputs (MAIN:130-272, 486-528): heatmap coordinates + boxplot arrays, per rank
            from .evalout import eval_policy, save_eval_outputs
            ret = eval_policy(env, lambda o: policy.select_actions(o), shapes, args.orientation, args.variant, mode="test", seed=args.seed + 99)
            out = os.path.join(args.eval_out, f"rank{rank}")
            save_eval_outputs(ret, sweep, os.path.join(out, "heatmap", "eval"), os.path.join(out, "boxplot", "eval"))
            s, n = all_reduce_sum([float(ret["num_success"]), float(env.n)], dev)
            log(phase="eval", tag=tag, sweep=sweep, success=s / n, avg_reward=ret["avg_reward"], wall_s=time.perf_counter() - t0)
            return
        succ, _ = run_episode(env, lambda o: policy.select_actions(o), None, shapes, args.orientation, args.variant, mode="test", seed=args.seed + 99)
        log(phase="eval", tag=tag, sweep=sweep, success=count(succ), wall_s=time.perf_counter() - t0)

    evaluate("untrained", 0)
    losses = None
    graphs = []
    if graphed and args.pretrain_updates:                     # pre-training: demonstrations only (DDPG:227-229)
        graphs.append(GraphedUpdates(policy, agent, expert, args.batch, 1.0))
        losses = graphs[-1].run(args.pretrain_updates)
    else:
        for _ in range(args.pretrain_updates):
            losses = policy.train_on_batch(*expert.sample_windows(args.batch, generator=gen))
    upd = GraphedUpdates(policy, agent, expert, args.batch, args.expert_prob) if graphed else None
    if upd is not None:
        graphs.append(upd)
    if args.pretrain_updates:
        log(phase="pretrain", updates=args.pretrain_updates, losses=[float(x) for x in losses])
        evaluate("pretrained", 0)
    env_steps = 0
    for sweep in range(1, args.sweeps + 1):
        ts = time.perf_counter()
        succ, stored = run_episode(env, lambda o: policy.select_actions(o, args.expl_noise, generator=gen), agent, shapes, args.orientation,
                                   args.variant, seed=args.seed + sweep)
        torch.cuda.synchronize()
        t_roll = time.perf_counter() - ts
        env_steps += env.n * env._max_episode_steps * world
        n_agent = int(args.batch * (1 - args.expert_prob))
        losses = None
        tu = time.perf_counter()
        # every rank takes the same branch (a rank that skipped while the others entered the all-reduce would hang them):
        # the update block runs only if EVERY shard holds enough episodes
        ready = all_reduce_min([1.0 if agent.replay_ep_num > 2 else 0.0], dev)[0] > 0.5
        if ready:
            if upd is not None:
                losses = upd.run(args.updates)
            else:
                for _ in range(args.updates):
                    a = agent.sample_windows(n_agent, generator=gen)
                    e = expert.sample_windows(args.batch - n_agent, generator=gen)
                    s_, a_, n_, r_, d_ = (torch.cat((x, y), 0) for x, y in zip(a, e))
                    losses = policy.train_on_batch(s_, a_, n_, r_, d_)
        torch.cuda.synchronize()
        t_upd = time.perf_counter() - tu
        log(phase="train", sweep=sweep, rollout_success=count(succ), agent_episodes=agent.replay_ep_num, rollout_s=t_roll, update_s=t_upd,
            rl_steps_per_s=env.n * env._max_episode_steps * world / t_roll,
            losses=None if losses is None else [float(x) for x in losses])
        if sweep % args.eval_every == 0 or sweep == args.sweeps:
            evaluate("policy", sweep)
    if args.save and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.save)), exist_ok=True)
        policy.save(args.save)
    # captured graphs hold NCCL work: release them before the process group goes (destroy_process_group otherwise waits forever)
    for g in graphs:
        if g.graph is not None:
            g.graph.reset()
        g.graph = None
    del graphs, upd
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
