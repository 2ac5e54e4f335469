#!/usr/bin/env python
"""Bucket /tmp/ncu_lines.json (written by tools/ncu_lines.py) by function, using the 'KG_DEV|KG_NOINL ... name(' definitions in kg_core.cuh."""
import json, re, sys, collections
src = open("kinovagrasping_b200/csrc/kg_core.cuh").read().split("\n")
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <class WS> )?(KG_DEV|KG_NOINL)\s+[\w\s\*&]+?\b(kg_\w+)\s*\(", l)
    if m: starts.append((i, m.group(2)))
def fn(line):
    name = "?"
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
d = json.load(open("/tmp/ncu_lines.json"))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for f, line, inst, smp, tinst, st in d:
    k = fn(line) if f == "kg_core.cuh" else f
    a = agg[k]; a[0] += inst; a[1] += smp; a[2] += tinst
    for s, v in st.items(): a[3][s] += v
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[0] == 0 and a[1] == 0: continue
    st = " ".join(f"{s}:{100.0*v/max(1,a[1]):.0f}%" for s, v in a[3].most_common(4))
    print(f"{k:28s} inst {100.0*a[0]/ti:5.1f}%  samples {100.0*a[1]/ts:5.1f}%  lanes {a[2]/max(1,a[0]):4.1f}  | {st}")
