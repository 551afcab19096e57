"""Per-function share of warp instructions / stall samples of one kernel: ncu `--page source --csv` rows joined with
`nvdisasm --print-line-info` of the same cubin (innermost inlined function of each SASS instruction).

    python tools/ncu_funcs.py gpurun_out/src.csv /tmp/sass/scan.sass readsb_b200/csrc/scan_kernel.cu [stall-column ...]
"""
import csv
import re
import sys
from collections import defaultdict

from sass_util import function_lines


def main():
    src_csv, sass, cu = sys.argv[1:4]
    rows = list(csv.reader(open(src_csv))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
    insts = [r for r in rows[2:] if len(r) >= len(hdr)]
    stack_re = re.compile(r'//## File "([^"]+)", line (\d+)'); ins_re = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
    locs, chain, cur = [], False, None
    for ln in function_lines(sass, len(insts)):
        m = stack_re.search(ln)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            if not chain:
                cur = loc
            chain = True
            continue
        if ins_re.match(ln):
            chain = False; locs.append(cur)
    print(len(locs), "sass instructions,", len(insts), "ncu rows")
    lines = open(cu).read().splitlines()
    fname = cu.split("/")[-1]
    funcs = [(i, m.group(1)) for i, l in enumerate(lines, 1) for m in [re.match(r'^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__).*?\b(\w+)\(', l)] if m]

    def fn(line):
        name = "?"
        for i, n in funcs:
            if i <= line:
                name = n
        return name
    extra = [c for c in hdr if c.startswith("stall_")] if len(sys.argv) > 4 and sys.argv[4] == "stalls" else []
    acc = defaultdict(lambda: [0, 0, 0] + [0] * len(extra)); tot = 0
    for loc, r in zip(locs, insts):
        ie = int(r[col["Instructions Executed"]] or 0); te = int(r[col["Thread Instructions Executed"]] or 0); smp = int(r[col["# Samples"]] or 0)
        tot += ie
        name = fn(loc[1]) if loc and loc[0] == fname else (loc[0] if loc else "none")
        a = acc[name]; a[0] += ie; a[1] += te; a[2] += smp
        for j, c in enumerate(extra):
            a[3 + j] += int(r[col[c]] or 0)
    ts = sum(a[2] for a in acc.values())
    for n, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        ex = "  ".join(f"{c[6:]}={a[3 + j]}" for j, c in enumerate(extra) if a[3 + j] > ts / 200)
        print(f"{n:>24} {a[0]:>11} {100 * a[0] / tot:5.1f}%  thr/inst {a[1] / max(a[0], 1):5.1f}  stalls {100 * a[2] / max(ts, 1):5.1f}%  {ex}")


if __name__ == "__main__":
    main()
