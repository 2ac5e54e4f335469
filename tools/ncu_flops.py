#!/usr/bin/env python
"""profiles/ncu_flops.json from an ncu CSV of every kernel of a run of tools/perf_probe.py (one full 30-step episode cycle):
   ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \\
       --clock-control none --csv --log-file gpurun_out/flops.csv python tools/perf_probe.py <lib> 4096 30
   python tools/ncu_flops.py gpurun_out/flops.csv 4096 30
FP32 flop executed = fadd + fmul + 2 ffma thread instructions (predicated on), summed over every kernel of the captured RL
steps, divided by envs x 15 substeps x steps.  bench.py reports it as fp32.frac_executed."""
import collections
import csv
import json
import os
import sys

path, envs, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    agg[name][r[mi]] += float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        launches[name] += 1
tot = collections.defaultdict(float)
per = {}
for k, m in agg.items():
    flop = m["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"] + 2 * m["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
    per[k] = {"launches": launches[k], "flop": flop, "warp_inst": m["smsp__inst_executed.sum"], "thread_inst": m["smsp__thread_inst_executed.sum"], "time_ns": m["gpu__time_duration.sum"]}
    for kk, vv in per[k].items():
        tot[kk] += vv
unit = envs * 15 * steps
out = {"envs": envs, "rl_steps": steps, "flop_per_env_substep": tot["flop"] / unit, "warp_inst_per_env_substep": tot["warp_inst"] / unit,
       "thread_inst_per_env_substep": tot["thread_inst"] / unit, "active_lanes_per_inst": tot["thread_inst"] / max(1.0, tot["warp_inst"]),
       "per_kernel": {k: dict(v, flop_per_env_substep=v["flop"] / unit, warp_inst_per_env_substep=v["warp_inst"] / unit, time_share=v["time_ns"] / tot["time_ns"]) for k, v in per.items()},
       "source": os.path.basename(path), "note": "ncu, every kernel of the captured RL steps (expert rollouts incl. auto-reset, phase-locked episode cycle)"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_flops.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}))
for k, v in out["per_kernel"].items():
    print(f"  {k:28s} launches {v['launches']:5d} flop/env-substep {v['flop_per_env_substep']:9.0f} warp-inst/env-substep {v['warp_inst_per_env_substep']:8.0f} time share {v['time_share']:.3f}")
