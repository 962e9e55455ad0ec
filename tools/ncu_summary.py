"""Condense an .ncu-rep (read here on the CPU box with `ncu -i`) into the few numbers the design
talks about; writes a markdown table.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep out.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{rep}` (`--set full --clock-control none`, one column per captured launch)", ""]
    names = [d[ix["Kernel Name"]].replace("void ", "").split("(")[0] for d in data]
    lines.append("| metric | unit | " + " | ".join(names) + " |")
    lines.append("|---|---|" + "---|" * len(names))
    for key, label in WANT:
        if key not in ix:
            continue
        vals = []
        for d in data:
            v = d[ix[key]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {label} (`{key}`) | {units[ix[key]]} | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
