#!/usr/bin/env python
"""Bit-compare two builds of the library on the configs[1] workload: python tools/cmp_libs.py a.so b.so [envs] [steps]."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kinovagrasping_b200 import _lib
from kinovagrasping_b200.handle import Handle
import torch
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
out = []
for path in sys.argv[1:3]:
    lib = _lib.bind(C.CDLL(os.path.abspath(path)))
    h = Handle(n, lib=lib)
    shapes = [int(x) for x in os.environ.get("KG_SHAPES", "0,1").split(",")]
    orient, mode = int(os.environ.get("KG_ORIENT", "0")), int(os.environ.get("KG_MODE", "2"))
    sh = torch.from_numpy(np.array(shapes, np.int32)[np.arange(n) % len(shapes)]).cuda()
    obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
    h.reset_sampled(0, sh.data_ptr(), orient, 1, 0, 1234, obs.data_ptr())
    tr = []
    for _ in range(steps):
        h.step(0, mode, obs.data_ptr(), rew.data_ptr(), done.data_ptr()); h.sync()
        tr += [obs.cpu().numpy().copy(), rew.cpu().numpy().copy()]
        h.reset_sampled(done.data_ptr(), 0, orient, 1, 0, 1234, 0)
    tr += list(h.get_state_host()); st = h.stats()
    out.append((tr, st)); h.close()
same = all(np.array_equal(x, y) for x, y in zip(out[0][0], out[1][0]))
worst = max(float(np.abs(x - y).max()) for x, y in zip(out[0][0], out[1][0]))
frac = float(np.mean(np.abs(out[0][0][-4] - out[1][0][-4]).max(axis=0) <= 1e-3))       # envs whose last obs agree
print(f"{sys.argv[1]} vs {sys.argv[2]}: bit-identical {same}, envs with |obs diff| <= 1e-3 after the last step {frac:.4f}, max |diff| {worst:.3e}; stats {out[0][1]} | {out[1][1]}")
