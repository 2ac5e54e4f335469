#!/usr/bin/env python
"""North-star check 4: expert-PID grasp/lift success rate over randomised starts, per shape, CUDA path vs the fp64 oracle
on the SAME starts (BASELINE.json: 10k starts per shape, within +-2 percentage points).

  python tools/success_sweep.py [--starts 10000] [--shapes 0,1,...] [--mode 2] [--lib path.so] [--out gpurun_out/success.json]

Starts: pose-table rows drawn with a fixed seed, hand orientation cycling normal/rotated/top (the Exp1 'random' choice,
kinova_gripper_env.py:1208-1220), object position and hand euler from the matching with_noise table.  An episode is 30
RL steps (TimeLimit) of the on-device combined expert; success = the terminal reward is the lift reward (ENV:631-687).
The oracle side runs one process per host core.  TEST/REPORT TOOL: the oracle is the checker, never the product."""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def starts(card_pose, shape, n, seed):
    rng = np.random.default_rng(seed + 7919 * shape)
    rows = rng.integers(1, 4500, n)
    ori = np.arange(n) % 3
    objp = np.stack([card_pose[shape, ori[i], rows[i], 0:3] for i in range(n)], 1).astype(np.float32)
    euler = np.stack([card_pose[shape, ori[i], rows[i], 6:9] for i in range(n)], 1).astype(np.float32)
    return euler, objp


def _oracle_worker(args):
    shape, euler, objp, mode = args
    from oracle import kgo
    m = kgo.Model()
    succ = np.zeros(euler.shape[1], bool)
    for i in range(euler.shape[1]):
        e, o = m.env_reset(int(shape), euler[:, i].astype(np.float64), objp[:, i].astype(np.float64))
        for _ in range(30):
            a, _ = m.expert_action(e, o, mode)
            o, r, d, _ = m.env_step(e, a)
            if d:
                succ[i] = r > 0
                break
    return succ


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--starts", type=int, default=10000)
    ap.add_argument("--shapes", default=",".join(str(i) for i in range(14)))
    ap.add_argument("--mode", type=int, default=2)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "success_sweep.json"))
    args = ap.parse_args()
    from kinovagrasping_b200 import _lib
    from kinovagrasping_b200.handle import Handle
    from kinovagrasping_b200.modelcard import ModelCard, SHAPES
    from oracle import kgo
    kgo.build()
    lib = _lib.bind(C.CDLL(os.path.abspath(args.lib))) if args.lib else _lib.load()
    pose = ModelCard().pose("train")
    cores = os.cpu_count() or 1
    n = args.starts
    rows = []
    with mp.get_context("spawn").Pool(cores) as pool:
        for sh in [int(x) for x in args.shapes.split(",")]:
            euler, objp = starts(pose, sh, n, args.seed)
            shape = np.full(n, sh, dtype=np.int32)
            t0 = time.perf_counter()
            h = Handle(n, lib=lib)
            h.reset_host(shape, euler, objp)
            done_k = np.zeros(n, bool); succ_k = np.zeros(n, bool)
            for _ in range(30):
                obs, rew, done = h.step_host(None, args.mode)
                first = done.astype(bool) & ~done_k
                succ_k[first] = rew[first] > 0
                done_k |= done.astype(bool)
            ovf = h.stats()["contact_overflow"]
            h.close()
            t_gpu = time.perf_counter() - t0
            t0 = time.perf_counter()
            chunks = np.array_split(np.arange(n), cores * 4)
            res = pool.map(_oracle_worker, [(sh, euler[:, c], objp[:, c], args.mode) for c in chunks if len(c)])
            succ_o = np.concatenate(res)
            t_cpu = time.perf_counter() - t0
            row = {"shape": SHAPES[sh], "starts": n, "success_cuda": float(succ_k.mean()), "success_oracle": float(succ_o.mean()),
                   "diff_pp": 100.0 * float(succ_k.mean() - succ_o.mean()), "per_start_agreement": float((succ_k == succ_o).mean()),
                   "all_done": bool(done_k.all()), "contact_overflow": int(ovf), "no_contact_dropped": bool(ovf == 0), "cuda_s": t_gpu, "oracle_s": t_cpu, "oracle_cores": cores}
            rows.append(row)
            print(f"{row['shape']:10s} starts {n}  cuda {100 * row['success_cuda']:6.2f}%  oracle {100 * row['success_oracle']:6.2f}%  "
                  f"diff {row['diff_pp']:+5.2f} pp  same outcome {100 * row['per_start_agreement']:5.1f}%  overflow {ovf}  ({t_gpu:.1f}s / {t_cpu:.1f}s)", flush=True)
    worst = max(abs(r["diff_pp"]) for r in rows)
    assert all(r["contact_overflow"] == 0 for r in rows), "a contact was dropped: the heavy lane must hold every contact of the model"
    out = {"mode": args.mode, "seed": args.seed, "rows": rows, "worst_abs_diff_pp": worst, "within_2pp": bool(worst <= 2.0), "contact_overflow_total": 0}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(f"worst |diff| {worst:.2f} pp -> {'within' if worst <= 2.0 else 'OUTSIDE'} +-2 pp")


if __name__ == "__main__":
    main()
