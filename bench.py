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
                       "envs": cores * per_core, "note": "mujoco-py 1.50 not installable; fp64 oracle port, rangefinders every substep as mj_step does"},
            "cpu_baseline": {"value": val, "unit": "physics substeps/s", "cores": cores, "kind": "port",
                             "sample": "%d RL steps x %d envs on %d cores" % (args.steps, cores * per_core, cores)},
            "e2e": {"value": val, "unit": "physics substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "wall_s": wall}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kinovagrasping_b200.env import VecKinovaGripperEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.envs
    env = VecKinovaGripperEnv(n, device=local)
    env.h.set_env_offset(rank * n)
    shapes = ("CubeS", "CubeB")
    variant = args.variant
    seed = 1234
    env.reset(shapes, "normal", variant=variant, seed=seed)
    dev = env.device
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    done_i32 = env.done

    def one_step():
        env.step(None, "combined")
        env.h.reset_sampled(done_i32.data_ptr(), 0, 0, 1 if variant == "B" else 0, 0, seed, env.obs_soa.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        one_step()
    env.h.stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (excluded from the timed spans)
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    stats = env.h.stats()
    clocks = sampler.stop() if rank == 0 else None
    # device time of the whole kernel chain of a step (CUDA events inside the library on the launch stream) ...
    kt = []
    for _ in range(5):
        flush.zero_()
        env.step(None, "combined")
        kt.append(env.h.last_step_ms())
        env.h.reset_sampled(done_i32.data_ptr(), 0, 0, 1 if variant == "B" else 0, 0, seed, env.obs_soa.data_ptr())
    kernel_ms = float(np.median(kt))
    # ... and of the dominant kernel alone: kg_sub_kernel<false> (one physics substep over the batch, 14 launches per RL
    # step), CUDA events after every launch on the launching stream (kg_set_timing; groups are not pipelined while it is on),
    # averaged over one full 30-step episode cycle of the same workload
    env.h.set_timing(1)
    brk = []
    for _ in range(30):
        flush.zero_()
        env.step(None, "combined")
        brk.append(env.h.timing_breakdown())
        env.h.reset_sampled(done_i32.data_ptr(), 0, 0, 1 if variant == "B" else 0, 0, seed, env.obs_soa.data_ptr())
    env.h.set_timing(0)
    try:
        fp32_peak_measured = env.h.measure_fp32_peak()         # FMA microbenchmark on this GPU
    except Exception:                                          # never let the roofline denominator break the bench line
        fp32_peak_measured = 0.0
    sub_ms = float(np.mean([b["substep"] for b in brk])) / (SUBSTEPS - 1)
    share = {k: float(np.mean([b[k] for b in brk])) for k in ("schedule", "begin", "narrowphase", "substep", "last")}

    # ---- e2e: host buffers through the C ABI (kg_step_host): pinned actions in, obs/reward/done out, every step
    act_h = torch.zeros((4, n), dtype=torch.float32).pin_memory()
    act_h[1:] = 0.5
    obs_h = torch.zeros((82, n), dtype=torch.float32).pin_memory()
    rew_h = torch.zeros(n, dtype=torch.float32).pin_memory()
    done_h = torch.zeros(n, dtype=torch.int32).pin_memory()
    import ctypes as C
    mask_d = torch.zeros(n, dtype=torch.int32, device=dev)
    def e2e_step():
        env.h._chk(env.h.lib.kg_step_host(env.h.h, C.c_void_p(act_h.data_ptr()), 0, C.c_void_p(obs_h.data_ptr()), C.c_void_p(rew_h.data_ptr()),
                                          C.c_void_p(done_h.data_ptr())), "kg_step_host")
        mask_d.copy_(done_h, non_blocking=True)
        env.h.reset_sampled(mask_d.data_ptr(), 0, 0, 1 if variant == "B" else 0, 0, seed, 0)
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms, sub_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms, sub_ms = [float(x) for x in t.tolist()]
    groups = 2 if n >= 512 else 1          # kg_set_pipeline default
    envs_total = n * world
    substeps = envs_total * SUBSTEPS * args.steps
    value = substeps / (total_ms * 1e-3)
    e2e_value = substeps / (e2e_ms * 1e-3)
    if rank == 0:
        hbm_peak, peak_src = peaks()
        alg_bytes = ALG_BYTES_PER_RL_STEP * n // SUBSTEPS          # 40 B per env per physics substep x envs per launch
        achieved = alg_bytes / (sub_ms * 1e-3) / 1e9
        fp32_achieved = (n * SUBSTEPS * ALG_FLOP_PER_SUBSTEP) / (kernel_ms * 1e-3) / 1e12       # whole kernel chain of a step
        if fp32_peak_measured > 0.5 * FP32_PEAK_TFLOPS:
            fp32_peak, fp32_src = fp32_peak_measured, "measured: kg_measure_fp32_peak FMA microbenchmark on this GPU"
        else:
            fp32_peak, fp32_src = FP32_PEAK_TFLOPS, "computed: 148 SMs x 128 lanes x 2 x 1.965 GHz"
        line = {"metric": "batched grasp env-steps/sec", "value": value, "unit": "physics substeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: %d envs/GPU CubeS/CubeB alternating, hand_orientation normal (pose-table variant %s), on-device "
                                       "expert PID (combined), 30-step episodes, auto-reset on done" % (n, variant),
                           "envs_per_gpu": n, "frame_skip": SUBSTEPS, "rl_steps_per_s": value / SUBSTEPS,
                           "l2": "flushed between timed steps by a 256 MiB write (excluded from the timed spans); state is 2.4 MB per 4096 envs",
                           "parallelism": "env-sharded x%d, no data-path collective" % world},
                "e2e": {"value": e2e_value, "unit": "physics substeps/s", "h2d_bytes_per_step": 16 * n + 4 * n, "d2h_bytes_per_step": (82 + 2) * 4 * n},
                # per step and env group: kg_schedule_kernel + kg_begin_kernel + 15 x (kg_mpr_kernel + kg_sub_kernel); + kg_reset_kernel
                "gpu_launches": (groups * (2 + 2 * SUBSTEPS) + 1) * args.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(n),
                             "peak_source": peak_src, "kernel": "kg_sub_kernel<false> (one physics substep over the batch; 14 of the 31 launches of a step)",
                             "kernel_ms": sub_ms, "step_chain_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             "step_ms_by_launch_kind": share,
                             "note": "state I/O is not the limiter (SURVEY 8d): the path is FP32-issue/latency bound, see fp32; "
                                     "traffic above the algorithmic bytes is the first touch of the model tables (hulls, BVH, pose rows)"},
                "fp32": {"achieved_tflops": fp32_achieved, "peak_tflops": fp32_peak, "frac": fp32_achieved / fp32_peak,
                         "peak_source": fp32_src, "computed_peak_tflops": FP32_PEAK_TFLOPS,
                         "basis": "3.3e5 algorithmic flop per substep (SURVEY 8d) x substeps / device time of the step's kernel chain"},
                "stats_per_substep": {"mpr_pairs": stats["mpr_pairs"] / max(1, stats["substeps"]), "newton_iters": stats["newton_iters"] / max(1, stats["substeps"]),
                                      "contact_overflow_frac": stats["contact_overflow"] / max(1, stats["substeps"])},
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
    ap.add_argument("--cpu-envs", type=int, default=16, help="cpu_baseline sample: env streams per host core")
    ap.add_argument("--cpu-steps", type=int, default=900, help="cpu_baseline sample: RL steps per env stream (~15 s of CPU work)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
