#!/usr/bin/env python
"""Static SASS size per source function of the profiled kernel: python tools/ncu_static.py <rep>"""
import csv, io, re, subprocess, sys, collections
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
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
cur = None; curline = None; cnt = collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0].isdigit(): curline = int(r[0]); continue
    if r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
        cnt[fn(curline) if cur == "kg_core.cuh" else cur] += 1
tot = sum(cnt.values()); print("static SASS instructions (rows with samples/exec info):", tot, f"= {tot*16/1024:.0f} KB")
for k, v in cnt.most_common(40): print(f"{k:28s} {v:6d} {v*16/1024:6.1f} KB")
