"""SASS size of one kernel bucketed by the kg_* source function its instructions come from (inlining flattened).

    cuobjdump -xelf all kinovagrasping_b200/libkg_b200.so      # -> kg_api.sm_100a.cubin
    nvdisasm -g kg_api.sm_100a.cubin > all.sass
    python tools/sass_hist.py all.sass kg_sub_kernelILb0

Prints instructions (16 bytes each on sm_100) per function: what to shrink when the kernel outgrows the 32 KB L1.5
instruction cache (profiles/README.md: stall_no_instruction)."""
import collections
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def function_ranges(path):
    starts = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:KG_DEV|KG_NOINL|__global__|__device__|static|inline)\b.*?\b(kg_\w+|launch_\w+|emu_\w+)\s*\(")
    with open(path) as f:
        for i, line in enumerate(f, 1):
            m = pat.match(line)
            if m:
                starts.append((i, m.group(1)))
    return starts


def owner(starts, line):
    name = "?"
    for s, n in starts:
        if s <= line:
            name = n
        else:
            break
    return name


def main():
    sass, kernel = sys.argv[1], sys.argv[2]
    by_line = len(sys.argv) > 3 and sys.argv[3] == "--lines"
    ranges = {}
    for fn in ("kinovagrasping_b200/csrc/kg_core.cuh", "kinovagrasping_b200/csrc/kg_api.cu", "kinovagrasping_b200/csrc/kg_model.h"):
        ranges[os.path.basename(fn)] = function_ranges(os.path.join(ROOT, fn))
    hist = collections.Counter()
    lines = collections.Counter()
    inside = False
    cur = ("?", 0)
    total = 0
    with open(sass) as f:
        for ln in f:
            if ln.startswith(".text."):
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                total += 1
                f_, l_ = cur
                name = owner(ranges.get(f_, []), l_) if f_ in ranges else f_
                hist[name] += 1
                lines[(f_, l_)] += 1
    print(f"{kernel}: {total} instructions = {total * 16 / 1024:.1f} KB")
    for name, c in hist.most_common(40):
        print(f"  {name:28s} {c:7d}  {c * 16 / 1024:6.1f} KB  {100.0 * c / total:5.1f} %")
    if by_line:
        for (f_, l_), c in lines.most_common(60):
            print(f"  {f_}:{l_:5d} {c:6d}")


if __name__ == "__main__":
    main()
