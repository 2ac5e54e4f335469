#!/usr/bin/env python
"""Kernel-time probe for build variants: python tools/perf_probe.py <lib.so> [envs] [steps].
Times kg_step_kernel with the library's own CUDA events (kg_last_step_ms) on the configs[1] workload."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kinovagrasping_b200 import _lib
from kinovagrasping_b200.handle import Handle

lib = _lib.bind(C.CDLL(os.path.abspath(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
variant = int(os.environ.get("KG_VARIANT", "1"))
h = Handle(n, lib=lib)
sched = int(os.environ.get("KG_SCHEDULE", "1"))
h.set_schedule(sched)
mode = int(os.environ.get("KG_MODE", "0"))
h.set_mode(mode)
groups = int(os.environ.get("KG_GROUPS", "0"))
if groups: h.set_pipeline(groups)
if os.environ.get("KG_TIMING"): h.set_timing(1)
if os.environ.get("KG_GRAPH") is not None: h.set_graph(int(os.environ["KG_GRAPH"]))
if os.environ.get("KG_PDL") is not None: h.set_dependent_launch(int(os.environ["KG_PDL"]))
if os.environ.get("KG_GRASP"): h.set_grasp_reward(True)          # grasp-classifier reward kernels behind every step
shape_set = [int(x) for x in os.environ.get("KG_SHAPES", "0,1").split(",")]
orient = int(os.environ.get("KG_ORIENT", "0"))           # 0 normal, 1 rotated, 2 top, 3 random
shape = np.array([shape_set[i % len(shape_set)] for i in range(n)], dtype=np.int32)
import torch
sh = torch.from_numpy(shape).cuda()
done = torch.zeros(n, dtype=torch.int32, device="cuda")
h.reset_sampled(0, sh.data_ptr(), orient, variant, 0, 1234, 0)
ms = []
brk = []
for t in range(steps):
    h.step(0, 2, 0, 0, done.data_ptr())
    ms.append(h.last_step_ms())
    if os.environ.get("KG_TIMING"):
        d = h.timing_breakdown()
        brk.append([0.0, d["schedule"], d["begin"], d["narrowphase"], d["substep"], d["last"], d.get("rays", 0.0)])
    h.reset_sampled(done.data_ptr(), 0, orient, variant, 0, 1234, 0)
h.sync()
st = h.stats()
ms = np.array(ms)
prof = (C.c_longlong * 32)()
lib.kg_get_prof(h.h, prof)
prof = list(prof)
print(f"{os.path.basename(sys.argv[1]):28s} sched {sched} envs {n} kernel ms: mean {ms[3:].mean():.3f} median {np.median(ms[3:]):.3f} min {ms[3:].min():.3f} max {ms[3:].max():.3f} "
      f"-> {n * 15 / ms[3:].mean() * 1e3:.3e} substeps/s | mpr/sub {st['mpr_pairs'] / st['substeps']:.2f} newton/sub {st['newton_iters'] / st['substeps']:.2f} ovf {st['contact_overflow']}"
      f" mode {mode} groups {groups} graph {os.environ.get('KG_GRAPH', 'default')} grasp {os.environ.get('KG_GRASP', '0')} pdl {os.environ.get('KG_PDL', 'default')}")
if brk:
    names = ["", "schedule", "begin", "narrowphase", "substep", "last", "rays"]
    for t in (5, 18):
        if t < len(brk): print(f"   step {t}: total {ms[t]:.3f} ms | " + "  ".join(f"{names[k]} {brk[t][k]:.3f}" for k in range(1, 7)))
if os.environ.get("KG_SERIES"): print("   ms per step:", " ".join(f"{x:.2f}" for x in ms))
tot = sum(prof[:20])
if tot:
    names = ["kin/dyn", "w>coll", "coll.lane", "compact", "w>rows", "rows", "w>solve", "sol.setup", "sol.cost", "sol.hess", "sol.chol", "sol.dirs",
             "sol.ls", "sol.epi", "w>euler", "euler", "rays", "obs", "load+w>dyn", "begin"]
    print("   per-env kernels, warp cycles: " + "  ".join(f"{nm} {100.0 * c / tot:.1f}%" for nm, c in zip(names, prof) if c) + f" | cycles/substep/env {tot / st['substeps']:.0f}"
          f" | narrowphase kernel: {prof[28] / st['substeps']:.0f} warp cycles/substep/env, {prof[29] / st['substeps']:.2f} tasks")
    ss = float(st["substeps"])
    cn = ["support", "climb", "mpr", "mpr_iter", "sepchk", "sep_ok", "plane_sup", "ls_eval"]
    print("   per substep: " + "  ".join(f"{nm} {prof[20 + i] / ss:.2f}" for i, nm in enumerate(cn)) + f" | cycles per MPR iteration {prof[28] / max(1, prof[23]):.0f}")
