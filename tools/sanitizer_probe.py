#!/usr/bin/env python
"""Small contact-rich run for compute-sanitizer: python tools/sanitizer_probe.py <envs> <steps> <mode 1|2>
Cone / vase / 45-degree shapes in the top grasp + three envs forced into a 25-contact state (hand pressed on the floor), so
that the light warps, the heavy lane (all-contacts workspace), the narrowphase, the ray kernel, graph replay and the masked
reset are all exercised."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kinovagrasping_b200.handle import Handle
n, steps, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
import torch
h = Handle(n)
h.set_mode(mode)
shapes = [0, 1, 4, 6, 9, 10, 13]
sh = torch.from_numpy(np.array([shapes[i % len(shapes)] for i in range(n)], dtype=np.int32)).cuda()
obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
h.reset_sampled(0, sh.data_ptr(), 2, 0, 0, 3, obs.data_ptr())
qp, qv = h.get_state_host()
for i in range(3):
    qp[0, i], qp[1, i], qp[2, i] = 0.0, -0.04, -0.08
    qp[3, i] = qp[5, i] = qp[7, i] = 1.2; qp[4, i] = qp[6, i] = qp[8, i] = 0.6
h.set_state_host(qp, qv)
for t in range(steps):
    h.step(0, 2, obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    h.reset_sampled(done.data_ptr(), 0, 2, 0, 0, 3, 0)
h.sync()
print("done", h.stats(), float(obs.abs().max()))
