#!/usr/bin/env python
"""Per-env timing of the per-env substep kernel (KG_TRACE build): python tools/trace_probe.py <lib_trace.so> [envs] [step]
For every substep launch of one RL step: mean / p50 / p99 / max warp cycles, and what the slowest warps were doing."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kinovagrasping_b200 import _lib
from kinovagrasping_b200.handle import Handle
import torch
lib = _lib.bind(C.CDLL(os.path.abspath(sys.argv[1])))
lib.kg_get_trace.restype = C.c_int; lib.kg_get_trace.argtypes = [C.c_void_p, C.c_void_p]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
at = int(sys.argv[3]) if len(sys.argv) > 3 else 18
h = Handle(n, lib=lib)
h.set_pipeline(int(os.environ.get("KG_GROUPS", "1")))
sh = torch.from_numpy((np.arange(n) % 2).astype(np.int32)).cuda()
done = torch.zeros(n, dtype=torch.int32, device="cuda")
h.reset_sampled(0, sh.data_ptr(), 0, 1, 0, 1234, 0)
for t in range(at + 1):
    h.step(0, 2, 0, 0, done.data_ptr())
    ms = h.last_step_ms()
    if t < at: h.reset_sampled(done.data_ptr(), 0, 0, 1, 0, 1234, 0)
tr = np.zeros((16, n, 2), dtype=np.int64)
assert lib.kg_get_trace(h.h, tr.ctypes.data_as(C.c_void_p)) == 0
print(f"RL step {at}: {ms:.3f} ms, {n} envs")
for s in range(15):
    w = tr[s, :, 0]; cyc = w >> 20; it = (w >> 12) & 0xff; nc = (w >> 4) & 0xff
    t0 = tr[s, :, 1] & 0xffffffffffff; sm = tr[s, :, 1] >> 48
    ok = cyc > 0
    if not ok.any(): continue
    c = cyc[ok]; span = (t0[ok] + c).max() - t0[ok].min()
    worst = np.argsort(-cyc)[:3]
    print(f" sub {s:2d}: warps {ok.sum():5d} cycles mean {c.mean():8.0f} p50 {np.median(c):8.0f} p99 {np.percentile(c, 99):8.0f} max {c.max():8.0f} | launch span {span:8.0f} "
          f"| start spread {t0[ok].max() - t0[ok].min():7.0f} | newton mean {it[ok].mean():.2f} max {it[ok].max()} | ncon mean {nc[ok].mean():.1f} max {nc[ok].max()} "
          f"| slowest: " + " ".join(f"(it {it[i]} ncon {nc[i]} cyc {cyc[i]})" for i in worst))
    # cycles by Newton iteration count
    if s in (0, 7, 13):
        print("          by newton iters: " + "  ".join(f"{k}: n={int((it[ok] == k).sum())} mean={cyc[ok][it[ok] == k].mean():.0f}" for k in sorted(set(it[ok].tolist()))[:10]))
