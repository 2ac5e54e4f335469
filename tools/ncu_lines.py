#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python tools/ncu_lines.py <rep> [top]
(ncu -i rep --page source --csv --print-source sass,cuda; aggregates the CUDA-line rows)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or len(r) < len(hdr): continue
    d = dict(zip(hdr[4:], r[4:]))
    def I(k):
        try: return int(d.get(k) or 0)
        except ValueError: return 0
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0", "-")}
    out.append((cur, int(r[0]), r[1].strip()[:90], I("Instructions Executed"), I("# Samples"), I("Thread Instructions Executed"), st))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total warp-inst {ti}  samples {ts}")
# per function-ish buckets by line ranges are left to the reader; print top lines by samples and by instructions
for key, name in ((4, "samples"), (3, "instructions")):
    print(f"--- top {top} lines by {name}")
    for o in sorted(out, key=lambda o: -o[key])[:top]:
        st = " ".join(f"{k}:{v}" for k, v in sorted(o[6].items(), key=lambda kv: -kv[1])[:3])
        print(f"{o[0]}:{o[1]:5d} inst {100.0*o[3]/ti:5.2f}% smp {100.0*o[4]/ts:5.2f}% lanes {o[5]/max(1,o[3]):4.1f} | {st} | {o[2]}")
import json
json.dump([(o[0], o[1], o[3], o[4], o[5], o[6]) for o in out], open("/tmp/ncu_lines.json", "w"))
