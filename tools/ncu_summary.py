"""Condenses an ncu report (`ncu --set full ... -o X`) into the CSV kept under profiles/: the metrics DESIGN.md argues with,
the stall-reason totals of the source page, and (optionally) a scan_traffic.json for bench.py's roofline.traffic.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/r02_X_ncu_summary.csv [--traffic profiles/scan_traffic.json "what was captured"]
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2]


def stall_totals(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return {}, 0
    hdr = rows[1]
    insts = [r for r in rows[2:] if len(r) >= len(hdr)]
    tot = {h: sum(int(r[i] or 0) for r in insts) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    return tot, len(insts)


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    names, units, vals = raw_page(rep)
    col = {n: i for i, n in enumerate(names)}
    lines = [("metric", "unit", "value")]
    kernel = vals[col["Kernel Name"]] if "Kernel Name" in col else ""
    lines.append(("kernel", "", kernel))
    for k in KEEP:
        if k in col:
            lines.append((k, units[col[k]], vals[col[k]]))
    st, n_sass = stall_totals(rep)
    total = sum(st.values()) or 1
    lines.append(("sass_instructions", "", str(n_sass)))
    for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
        lines.append((f"pcsamp_{k}", "samples (% of all)", f"{v} ({100.0 * v / total:.1f}%)"))
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    print(dst, "kernel", kernel[:60], "us", vals[col["gpu__time_duration.sum"]])
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        import bench
        rd, wr = float(vals[col["dram__bytes_read.sum"]]), float(vals[col["dram__bytes_write.sum"]])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd *= scale[units[col["dram__bytes_read.sum"]]]; wr *= scale[units[col["dram__bytes_write.sum"]]]
        json.dump({"kernel": "scan_kernel", "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                   "algorithmic_bytes_per_launch": 268435456, "source_digest": bench.scan_source_digest(), "capture": sys.argv[i + 2],
                   "summary": str(Path(dst).name)}, open(sys.argv[i + 1], "w"), indent=1)
        print("traffic", (rd + wr) / 268435456.0, "x algorithmic")


if __name__ == "__main__":
    main()
