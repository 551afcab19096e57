"""Correlates an ncu `--page source --csv` dump (per-SASS-instruction counters) with source lines using
nvdisasm --print-line-info on the cubin extracted from the built library.

    python tools/ncu_lines.py gpurun_out/src.csv /tmp/sass/scan.sass [top_n]
"""
import csv
import re
import sys
from collections import defaultdict

from sass_util import function_lines


def main():
    src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    insts = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        insts.append(r)
    # nvdisasm: lines "//## File "...", line N" precede instructions "/*0010*/  OP ..."
    lines = []
    cur = None
    stack_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins_re = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
    for ln in function_lines(sass, len(insts)):
        m = stack_re.search(ln)
        if m:
            if "inlined at" in ln:
                # keep the innermost (first) location of an inlining chain only
                pass
            cur_candidate = (m.group(1).split("/")[-1], int(m.group(2)))
            if "inlined at" not in ln or cur is None or True:
                if not getattr(main, "_chain", False):
                    cur = cur_candidate
                main._chain = True
            continue
        m = ins_re.match(ln)
        if m:
            main._chain = False
            lines.append((int(m.group(1), 16), cur, m.group(2)))
    if len(lines) != len(insts):
        print(f"warning: {len(lines)} sass instructions vs {len(insts)} ncu rows", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0, 0])
    total = 0
    for (off, loc, text), r in zip(lines, insts):
        ie = int(r[col["Instructions Executed"]] or 0)
        te = int(r[col["Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        wf = int(r[col["L1 Wavefronts Shared"]] or 0)
        a = agg[loc]
        a[0] += ie; a[1] += te; a[2] += smp; a[3] += wf
        total += ie
    tot_s = sum(a[2] for a in agg.values())
    print(f"total warp instructions {total}, samples {tot_s}")
    src_cache = {}
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            try:
                if loc[0] not in src_cache:
                    import glob
                    cand = glob.glob(f"readsb_b200/csrc/{loc[0]}")
                    src_cache[loc[0]] = open(cand[0]).read().splitlines() if cand else []
                text = src_cache[loc[0]][loc[1] - 1].strip()[:110]
            except Exception:
                pass
        print(f"{a[0]:>12} {100 * a[0] / total:5.1f}%  thr/inst {a[1] / max(a[0], 1):5.1f}  stall-samples {100 * a[2] / max(tot_s, 1):5.1f}%  smem-wf {a[3]:>10}  {loc}  {text}")


if __name__ == "__main__":
    main()


def grouped(src_csv, sass, groups):
    """groups: list of (name, file, lo, hi) line ranges; prints warp-instruction share per group."""
    import csv as _csv
    rows = list(_csv.reader(open(src_csv)))
    hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
    insts = [r for r in rows[2:] if len(r) >= len(hdr)]
    stack_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins_re = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
    locs, chain, cur, outer = [], False, None, None
    for ln in open(sass):
        m = stack_re.search(ln)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            if not chain:
                cur = loc
            outer = loc          # last location of an inlining chain = outermost caller
            chain = True
            continue
        m = ins_re.match(ln)
        if m:
            chain = False
            locs.append((cur, outer))
    tot = 0
    acc = defaultdict(lambda: [0, 0, 0])
    for (cur, outer), r in zip(locs, insts):
        ie = int(r[col["Instructions Executed"]] or 0); smp = int(r[col["# Samples"]] or 0); wf = int(r[col["L1 Wavefronts Shared"]] or 0)
        tot += ie
        name = "other"
        for g, f, lo, hi in groups:
            if outer and outer[0] == f and lo <= outer[1] <= hi:
                name = g; break
        acc[name][0] += ie; acc[name][1] += smp; acc[name][2] += wf
    ts = sum(a[1] for a in acc.values())
    for name, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:>22}: {a[0]:>12} warp-instr {100*a[0]/tot:5.1f}%   stall samples {100*a[1]/max(ts,1):5.1f}%   smem wavefronts {a[2]}")
