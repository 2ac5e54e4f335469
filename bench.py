#!/usr/bin/env python
"""bench.py -- batched grasp env-steps/sec (BASELINE.json metric) on configs[1]:
4096 envs per GPU, CubeS/CubeB alternating, hand_orientation "normal", on-device expert PID (combined),
30-step episodes with auto-reset on done.  One "step" = one RL step (15 physics substeps + obs + reward + controller +
auto-reset: begin | 15 x (narrowphase | per-env substep) kernels per env group, then the masked reset) over the whole batch.
See DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--variant A|B]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_RL_STEP = 600          # read qpos16+qvel15+action4, write qpos16+qvel15+obs82+reward+done (fp32)
ALG_FLOP_PER_SUBSTEP = 3.3e5         # SURVEY 8d canonical workload constants
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
SUBSTEPS = 15


def ncu_traffic(n_envs):
    """DRAM bytes per launch of the dominant kernel (kg_sub_kernel<false>, one physics substep over the batch) from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json), only if it was taken at this env count; else None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if int(d.get("envs", -1)) == int(n_envs):
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def _cpu_worker(args):
    wid, n_env, n_steps, rays_every, variant = args
    import numpy as np
    from oracle import kgo
    from kinovagrasping_b200.modelcard import ModelCard
    m = kgo.Model()
    kgo.lib().kgo_set_rays_every_substep(1 if rays_every else 0)
    pose = ModelCard().pose("train")
    rng = np.random.default_rng(1000 + wid)
    envs = []

    def draw(sh):
        r = int(rng.integers(1, 4500))
        objp = pose[sh, 0, r, 0:3].astype(np.float64)
        euler = pose[sh, 2 if variant == "B" else 0, r, 6:9].astype(np.float64)
        return m.env_reset(sh, euler, objp)
    for i in range(n_env):
        sh = (wid * n_env + i) % 2
        e, o = draw(sh)
        envs.append([e, o, sh])
    # episode phases desynchronised before the clock starts, as the GPU arm does (env i has run i mod 30 steps of its
    # episode): the timed steps see the steady-state mix of approach / grasp / lift, not a batch of contact-free episode starts
    for i, rec in enumerate(envs):
        for _ in range(i % 30):
            a, _ = m.expert_action(rec[0], rec[1], 2)
            o, r, d, _ = m.env_step(rec[0], a)
            rec[1] = o
            if d:
                rec[0], rec[1] = draw(rec[2])
    t0 = time.perf_counter()
    for _ in range(n_steps):
        for rec in envs:
            a, _ = m.expert_action(rec[0], rec[1], 2)
            o, r, d, _ = m.env_step(rec[0], a)
            rec[1] = o
            if d:
                rec[0], rec[1] = draw(rec[2])
    return time.perf_counter() - t0


def cpu_baseline(n_env_per_core, n_steps, rays_every=True, cores=None, variant="B"):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    from oracle import kgo
    kgo.build()
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(w, n_env_per_core, n_steps, rays_every, variant) for w in range(cores)])
        wall = time.perf_counter() - t0
    work = cores * n_env_per_core * n_steps * SUBSTEPS
    return work / max(times), cores, wall


def run_reference(args):
    """--impl reference: the reference's CPU path.  mujoco-py 1.50 / MuJoCo 1.50 cannot be installed (closed binary, no
    network), so this times the fp64 oracle PORT of that path on all host cores, one env stream per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core, chunk = 128, 1
    # every "step" = one RL step over cores*per_core envs
    import multiprocessing as mp
    from oracle import kgo
    kgo.build()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(w, per_core, max(1, args.warmup), True, args.variant) for w in range(cores)])
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(w, per_core, args.steps * chunk, True, args.variant) for w in range(cores)])
        wall = time.perf_counter() - t0
    work = cores * per_core * args.steps * SUBSTEPS
    val = work / max(times)
    line = {"impl": "reference", "metric": "batched grasp env-steps/sec", "value": val, "unit": "physics substeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * max(times) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1] sample: CubeS/CubeB normal-orientation (pose-table variant %s) expert-PID rollouts, %d envs per core" % (args.variant, per_core),
                       "envs": cores * per_core, "note": "mujoco-py 1.50 not installable; fp64 oracle port, rangefinders every substep as mj_step does; episode phases desynchronised before timing"},
            "cpu_baseline": {"value": val, "unit": "physics substeps/s", "cores": cores, "kind": "port",
                             "sample": "%d RL steps x %d envs on %d cores" % (args.steps, cores * per_core, cores)},
            "e2e": {"value": val, "unit": "physics substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "wall_s": wall}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ ours
def flop_per_substep(n_mpr, n_con, n_newton):
    """BASELINE.md section 3 / SURVEY 8d: algorithmic flop of one physics substep for given workload counts."""
    return 8.3e3 + 19e3 * n_mpr + 2e3 + 0.55e3 * n_con + 2e3 + 11e3 * n_newton + 8e3


def ncu_executed_flop():
    """FP32 flop actually executed per env-substep by the kernel chain, from the committed ncu capture
    (profiles/ncu_flops.json: op_fadd + op_fmul + 2 op_ffma thread instructions of every kernel of one RL step); else None."""
    p = os.path.join(ROOT, "profiles", "ncu_flops.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class Workload:
    """configs[1] on one GPU: `n` envs CubeS/CubeB alternating, pose-table variant A or B, on-device combined expert,
    30-step episodes with auto-reset.  Episode phases are DESYNCHRONISED before anything is timed (env i is restarted at
    warm-up step i mod 30), so every timed step sees the same steady-state mix of approach / grasp / lift instead of one
    phase-locked episode."""

    def __init__(self, n, local, rank, variant, seed=1234):
        import torch
        from kinovagrasping_b200.env import VecKinovaGripperEnv
        self.torch, self.n, self.variant, self.seed = torch, n, variant, seed
        self.env = VecKinovaGripperEnv(n, device=local)
        self.env.h.set_env_offset(rank * n)
        self.env.reset(("CubeS", "CubeB"), "normal", variant=variant, seed=seed)
        self.done = self.env.done
        self.v = 1 if variant == "B" else 0
        ids = torch.arange(n, device=self.env.device)
        for k in range(30):
            self.env.step(None, "combined")
            mask = ((ids % 30) == k).to(torch.int32) | self.done
            self.env.h.reset_sampled(mask.data_ptr(), 0, 0, self.v, 0, seed, self.env.obs_soa.data_ptr())
        self.env.h.stats()

    def step(self):
        self.env.step(None, "combined")
        self.env.h.reset_sampled(self.done.data_ptr(), 0, 0, self.v, 0, self.seed, self.env.obs_soa.data_ptr())

    def timed(self, steps, flush):
        """ms of `steps` consecutive steps, CUDA events on the launching stream, L2 flushed before every step (outside the spans)."""
        t = self.torch
        ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()
            ev[k][0].record()
            self.step()
            ev[k][1].record()
        t.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    def close(self):
        self.env.h.close()


def learner_bench(dev, world, updates=50, batch=4096):
    """config 4: ms per DDPGfD update (batch 4096 five-step windows, device-side sampling, CUDA graph of 10 updates; under
    torchrun the two NCCL all-reduces of an update are nodes of that graph) and the all-reduce alone."""
    import torch
    from kinovagrasping_b200.learner import DDPGfD
    from kinovagrasping_b200.replay import ReplayBuffer_Queue
    from kinovagrasping_b200.train import GraphedUpdates
    torch.manual_seed(0)
    torch.set_float32_matmul_precision("high")
    pol = DDPGfD(82, 4, 0.8, 5, batch_size=batch, device=dev, capturable=True)
    bufs = []
    for _ in range(2):
        rb = ReplayBuffer_Queue(82, 4, 4096, 5, device=dev, n_envs=2048)
        s = torch.randn(2048, 83, device=dev)
        for t in range(24):
            st = torch.randn(2048, 82, device=dev)
            rb.add_batch(st, torch.rand(2048, 4, device=dev), st + 0.01, torch.zeros(2048, device=dev), torch.full((2048,), t == 23, device=dev), defer=True)
        rb.flush()
        bufs.append(rb)
    g = GraphedUpdates(pol, bufs[0], bufs[1], batch, 0.3)
    g.run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.run(updates); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / updates
    ar_us = None
    if world > 1:
        import torch.distributed as dist
        for _ in range(5): dist.all_reduce(pol.flat_grad)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50): dist.all_reduce(pol.flat_grad)
        e1.record(); torch.cuda.synchronize()
        ar_us = e0.elapsed_time(e1) / 50 * 1e3
    flop = 2 * 3 * batch * (5 * (82 * 400 + 400 * 300 + 300 * 4) + 6 * (86 * 400 + 400 * 300 + 300))      # rough: fwd+bwd of actor (5 states) and critic
    return {"ms_per_update": ms, "batch_windows": batch, "n_step": 5, "graph_block": 10, "allreduce_us_310105_f32": ar_us,
            "tf32_tflops_est": flop / (ms * 1e-3) / 1e12, "params": 310105}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.envs
    variant = args.variant
    dev = torch.device("cuda", local)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = Workload(n, local, rank, variant)
    env = w.env
    for _ in range(max(3, args.warmup)):
        w.step()
    env.h.stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: K steps per repetition, repeated back to back until >= ~1 s of device time has been measured, so that a
    # single scheduling hiccup cannot move the result (the per-step figure is over all repetitions)
    t_wall = time.perf_counter()
    rep_ms = [float(sum(w.timed(args.steps, flush)))]
    reps = int(min(200, max(1, np.ceil(args.min_ms / max(rep_ms[0], 1e-3)))))
    for _ in range(reps - 1):
        rep_ms.append(float(sum(w.timed(args.steps, flush))))
    barrier()
    t_wall = time.perf_counter() - t_wall
    total_ms = float(sum(rep_ms))
    steps_total = args.steps * len(rep_ms)
    stats = env.h.stats()
    clocks = sampler.stop() if rank == 0 else None
    # device time of the whole kernel chain of a step (CUDA events inside the library on the launch stream) ...
    kt = []
    for _ in range(30):
        flush.zero_()
        env.step(None, "combined")
        kt.append(env.h.last_step_ms())
        env.h.reset_sampled(w.done.data_ptr(), 0, 0, w.v, 0, w.seed, env.obs_soa.data_ptr())
    kernel_ms = float(np.mean(kt))
    # ... and of the dominant kernel alone: kg_sub_kernel<false> (one physics substep over the batch, 14 launches per RL
    # step), CUDA events after every launch on the launching stream (kg_set_timing; groups are not pipelined and no graph is
    # replayed while it is on), averaged over 30 steps of the same steady-state mix
    env.h.set_timing(1)
    brk = []
    for _ in range(30):
        flush.zero_()
        env.step(None, "combined")
        brk.append(env.h.timing_breakdown())
        env.h.reset_sampled(w.done.data_ptr(), 0, 0, w.v, 0, w.seed, env.obs_soa.data_ptr())
    env.h.set_timing(0)
    try:
        fp32_peak_measured = env.h.measure_fp32_peak()         # FMA microbenchmark on this GPU
    except Exception:                                          # never let the roofline denominator break the bench line
        fp32_peak_measured = 0.0
    sub_ms = float(np.mean([b["substep"] for b in brk])) / (SUBSTEPS - 1)
    share = {k: float(np.mean([b[k] for b in brk])) for k in ("schedule", "begin", "narrowphase", "substep", "last", "rays")}

    # ---- e2e: host buffers through the C ABI (kg_step_host): pinned actions in, obs/reward/done out, every step
    act_h = torch.zeros((4, n), dtype=torch.float32).pin_memory()
    obs_h = torch.zeros((82, n), dtype=torch.float32).pin_memory()
    rew_h = torch.zeros(n, dtype=torch.float32).pin_memory()
    done_h = torch.zeros(n, dtype=torch.int32).pin_memory()
    import ctypes as C
    mask_d = torch.zeros(n, dtype=torch.int32, device=dev)
    # The reference's split, every step: action = controller.get_action(obs) (here kg_expert_action_host: the on-device
    # controller, its action read back to pinned host memory), then env.step(action) with that HOST action (kg_step_host: H2D of
    # the action, the step, D2H of obs / reward / done), then the masked reset of finished envs.  Same controller, same actions
    # and therefore the same contact mix as the device-resident `value` loop, with every action crossing the bus both ways.
    def e2e_step():
        env.h._chk(env.h.lib.kg_expert_action_host(env.h.h, 2, C.c_void_p(act_h.data_ptr())), "kg_expert_action_host")
        env.h._chk(env.h.lib.kg_step_host(env.h.h, C.c_void_p(act_h.data_ptr()), 0, C.c_void_p(obs_h.data_ptr()), C.c_void_p(rew_h.data_ptr()),
                                          C.c_void_p(done_h.data_ptr())), "kg_step_host")
        mask_d.copy_(done_h, non_blocking=True)
        env.h.reset_sampled(mask_d.data_ptr(), 0, 0, w.v, 0, w.seed, 0)
    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = max(args.steps, min(400, int(np.ceil(args.min_ms / max(rep_ms[0] / args.steps, 1e-3)))))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    # what the e2e step consists of (untimed extra pass): device time of the step's kernel chain inside kg_step_host, the rest is
    # copies + launch + synchronisation.  stats of the e2e workload next to it.
    env.h.stats()
    e2e_kernel = []
    for _ in range(60):
        e2e_step()
        e2e_kernel.append(env.h.last_step_ms())
    st_e2e = env.h.stats()
    e2e_diag = {"device_ms_per_step_kernel_chain": float(np.mean(e2e_kernel)),
                "mpr_pairs": st_e2e["mpr_pairs"] / max(1, st_e2e["substeps"]), "newton_iters": st_e2e["newton_iters"] / max(1, st_e2e["substeps"]),
                "contacts": st_e2e["contacts"] / max(1, st_e2e["substeps"])}
    groups = 2 if n >= 32768 else (4 if n >= 2048 else (2 if n >= 512 else 1))          # kg_set_pipeline default

    # ---- extra workloads for the record (rank 0, single GPU only): the other pose-table variant and the 32768-env batch
    extra = {}
    if world == 1 and not args.no_extra:
        for tag, (en, ev) in {"variant_" + ("A" if variant == "B" else "B"): (n, "A" if variant == "B" else "B"), "envs_32768": (32768, variant)}.items():
            w2 = Workload(en, local, rank, ev)
            for _ in range(3):
                w2.step()
            w2.env.h.stats()
            ms2 = w2.timed(40, flush)
            st2 = w2.env.h.stats()
            extra[tag] = {"envs": en, "variant": ev, "value": en * SUBSTEPS * len(ms2) / (sum(ms2) * 1e-3), "unit": "physics substeps/s", "ms_per_step": float(np.mean(ms2)),
                          "steps": len(ms2), "mpr_pairs": st2["mpr_pairs"] / max(1, st2["substeps"]), "newton_iters": st2["newton_iters"] / max(1, st2["substeps"]),
                          "contacts": st2["contacts"] / max(1, st2["substeps"]), "contact_overflow": st2["contact_overflow"]}
            w2.close()
    learner = None
    if not args.no_learner:
        try:
            learner = learner_bench(dev, world)
        except Exception as e:                                  # the env metric is the contract; the learner line is an extra
            learner = {"error": repr(e)[:300]}

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms, sub_ms, (learner or {}).get("ms_per_update") or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, sub_ms, learner_ms = [float(x) for x in t.tolist()]
    envs_total = n * world
    value = envs_total * SUBSTEPS * steps_total / (total_ms * 1e-3)
    e2e_value = envs_total * SUBSTEPS * e2e_steps / (e2e_ms * 1e-3)
    if rank == 0:
        hbm_peak, peak_src = peaks()
        alg_bytes = ALG_BYTES_PER_RL_STEP * n // SUBSTEPS          # 40 B per env per physics substep x envs per launch
        achieved = alg_bytes / (sub_ms * 1e-3) / 1e9
        if fp32_peak_measured > 0.5 * FP32_PEAK_TFLOPS:
            fp32_peak, fp32_src = fp32_peak_measured, "measured: kg_measure_fp32_peak FMA microbenchmark on this GPU"
        else:
            fp32_peak, fp32_src = FP32_PEAK_TFLOPS, "computed: 148 SMs x 128 lanes x 2 x 1.965 GHz"
        ss = max(1, stats["substeps"])
        n_mpr, n_con, n_newton = stats["mpr_pairs"] / ss, stats["contacts"] / ss, stats["newton_iters"] / ss
        chain_rate = n * SUBSTEPS / (kernel_ms * 1e-3)             # substeps/s of the step's kernel chain (device time)
        f_meas = flop_per_substep(n_mpr, n_con, n_newton)
        exe = ncu_executed_flop()
        fp32 = {"peak_tflops": fp32_peak, "peak_source": fp32_src, "computed_peak_tflops": FP32_PEAK_TFLOPS,
                # the honest figure: BASELINE.md's formula with the counts THIS run measured
                "frac_measured_counts": chain_rate * f_meas / 1e12 / fp32_peak, "flop_per_substep_measured_counts": f_meas,
                "measured_counts": {"n_mpr": n_mpr, "n_con": n_con, "n_newton": n_newton},
                # what the hardware executed (ncu op_fadd + op_fmul + 2 op_ffma over one RL step's kernels), if the capture is committed
                "frac_executed": (chain_rate * exe["flop_per_env_substep"] / 1e12 / fp32_peak) if exe else None,
                "flop_per_substep_executed": exe["flop_per_env_substep"] if exe else None,
                # SURVEY 8d's canonical near-contact workload (N_mpr 14, N_con 8, N_newton 4): an upper reference, NOT this run's work
                "frac_canonical_workload": chain_rate * ALG_FLOP_PER_SUBSTEP / 1e12 / fp32_peak, "flop_per_substep_canonical": ALG_FLOP_PER_SUBSTEP,
                "basis": "flop per substep x substeps / device time of the step's kernel chain (kg_last_step_ms)"}
        line = {"metric": "batched grasp env-steps/sec", "value": value, "unit": "physics substeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / steps_total, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: %d envs/GPU CubeS/CubeB alternating, hand_orientation normal (pose-table variant %s), on-device "
                                       "expert PID (combined), 30-step episodes desynchronised over their phase, auto-reset on done" % (n, variant),
                           "envs_per_gpu": n, "frame_skip": SUBSTEPS, "rl_steps_per_s": value / SUBSTEPS,
                           "timed": {"steps_per_repetition": args.steps, "repetitions": len(rep_ms), "steps_total": steps_total, "device_ms_total": total_ms,
                                     "repetition_ms_min_max": [min(rep_ms), max(rep_ms)]},
                           "l2": "flushed between timed steps by a 256 MiB write (excluded from the timed spans); state is 2.4 MB per 4096 envs",
                           "parallelism": "env-sharded x%d, no data-path collective" % world},
                "e2e": {"value": e2e_value, "unit": "physics substeps/s", "h2d_bytes_per_step": 16 * n + 4 * n, "d2h_bytes_per_step": (82 + 2 + 4) * 4 * n, "steps": e2e_steps,
                        "ms_per_step": e2e_ms / e2e_steps, "workload": "per step: kg_expert_action_host (combined expert, action D2H) -> kg_step_host (action H2D, obs/reward/done D2H) -> masked sampled reset; same actions as the device-resident loop",
                        **e2e_diag},
                # kernels per step: per env group kg_schedule_kernel + kg_begin_kernel + 15 x (kg_mpr_kernel + kg_sub_kernel) + kg_rays_kernel, + kg_reset_kernel
                # (the chain of a step is replayed as ONE CUDA graph launch)
                "gpu_launches": (groups * (3 + 2 * SUBSTEPS) + 1) * steps_total,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(n),
                             "peak_source": peak_src, "kernel": "kg_sub_kernel<false> (one physics substep over the batch; 14 of the 31 launches per group of a step)",
                             "kernel_ms": sub_ms, "step_chain_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             "step_ms_by_launch_kind": share,
                             "note": "state I/O is not the limiter (SURVEY 8d): the path is FP32-issue/latency bound, see fp32.  traffic (ncu, cold L2) "
                                     "above the algorithmic bytes is the per-env carry record between the kernels of a step: L2-resident in production"},
                "fp32": fp32,
                "stats_per_substep": {"mpr_pairs": n_mpr, "contacts": n_con, "newton_iters": n_newton, "contact_overflow": stats["contact_overflow"]},
                "extra": extra, "learner": dict(learner or {}, ms_per_update_max_over_ranks=learner_ms) if learner else None,
                "clocks": clocks, "wall_s": t_wall}
        if world == 1 and not args.no_cpu:
            v, cores, wall = cpu_baseline(args.cpu_envs, args.cpu_steps, True, None, variant)
            line["cpu_baseline"] = {"value": v, "unit": "physics substeps/s", "cores": cores, "kind": "port",
                                    "sample": "%d RL steps x %d envs/core on %d cores (%.1f s), fp64 oracle port with rangefinders every substep; "
                                              "mujoco-py 1.50 is not installable here" % (args.cpu_steps, args.cpu_envs, cores, wall)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--variant", default="B", choices=["A", "B"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the variant-A and 32768-env side measurements")
    ap.add_argument("--no-learner", action="store_true", help="skip the DDPGfD update microbenchmark")
    ap.add_argument("--min-ms", type=float, default=1000.0, help="repeat the K-step timed region until this much device time is measured")
    ap.add_argument("--cpu-envs", type=int, default=16, help="cpu_baseline sample: env streams per host core")
    ap.add_argument("--cpu-steps", type=int, default=900, help="cpu_baseline sample: RL steps per env stream (~15 s of CPU work)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
