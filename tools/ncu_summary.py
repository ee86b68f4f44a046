#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) into a small markdown file for profiles/ (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_euler_rNN.md "title" """
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__ops_path_tensor_src_fp64.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__shared_mem_per_block_dynamic"]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu summary"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source report: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        lines += [f"## {name}", "", "| metric | unit | value |", "|---|---|---|"]
        for k in KEYS:
            if k in h:
                lines.append(f"| `{k}` | {units[h.index(k)]} | {r[h.index(k)]} |")
        st = []
        for i, n in enumerate(h):
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        lines += ["", "warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]), ""]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    try:
        hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
        sh = srows[hi]
        si, ci = sh.index("# Samples"), sh.index("Source")
        data = [r for r in srows[hi + 1:] if len(r) > si and r[si].isdigit()]
        tot = sum(int(r[si]) for r in data)
        lines += ["### hottest SASS instructions (warp-stall samples)", "", "| samples | share | instruction |", "|---|---|---|"]
        for r in sorted(data, key=lambda r: -int(r[si]))[:15]:
            lines.append(f"| {r[si]} | {int(r[si]) / tot:.3f} | `{r[ci].strip()[:80]}` |")
        ops = {}
        for r in data:
            op = r[ci].strip().split()[0] if r[ci].strip() else "?"
            if op.startswith("@"):
                op = r[ci].strip().split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[si])
        lines += ["", "samples by opcode: " + ", ".join(f"{k} {v / tot:.3f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10]), ""]
    except Exception as e:  # noqa: BLE001
        lines.append(f"(source page unavailable: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
