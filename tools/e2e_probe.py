#!/usr/bin/env python
"""Where the end-to-end step's time goes (bench.py's e2e loop with a host clock around every call):
python tools/e2e_probe.py [envs] [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
w = bench.Workload(n, 0, 0, "B")
env = w.env
act_h = torch.zeros((4, n), dtype=torch.float32).pin_memory()
obs_h = torch.zeros((82, n), dtype=torch.float32).pin_memory()
rew_h = torch.zeros(n, dtype=torch.float32).pin_memory()
done_h = torch.zeros(n, dtype=torch.int32).pin_memory()
mask_d = torch.zeros(n, dtype=torch.int32, device="cuda")
lib, h = env.h.lib, env.h.h
P = lambda t: C.c_void_p(t.data_ptr())
T = np.zeros((steps, 5))
for k in range(steps + 10):
    t0 = time.perf_counter()
    lib.kg_expert_action_host(h, 2, P(act_h))
    t1 = time.perf_counter()
    lib.kg_step_host(h, P(act_h), 0, P(obs_h), P(rew_h), P(done_h))
    t2 = time.perf_counter()
    mask_d.copy_(done_h, non_blocking=True)
    t3 = time.perf_counter()
    env.h.reset_sampled(mask_d.data_ptr(), 0, 0, w.v, 0, w.seed, 0)
    t4 = time.perf_counter()
    if k >= 10:
        T[k - 10] = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, env.h.last_step_ms() * 1e-3]
m = T.mean(0) * 1e6
print(f"envs {n}: expert_action_host {m[0]:.1f} us | step_host {m[1]:.1f} us (kernel chain {m[4]:.1f} us, rest {m[1] - m[4]:.1f}) | mask copy {m[2]:.1f} us | "
      f"reset launch {m[3]:.1f} us | total {m[:4].sum():.1f} us -> {n * 15 / m[:4].sum() * 1e6:.3e} substeps/s")
