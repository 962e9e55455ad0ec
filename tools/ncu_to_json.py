"""profiles/score_kernel_ncu.json from an `ncu --set full` report of the scoring kernel (read here on the CPU box):
the utilisation figures bench.py quotes as the kernel's binding resource, tagged with the sha of the kernel source so
that bench.py can tell when the capture is stale.
  python tools/ncu_to_json.py gpurun_out/prof.ncu-rep "capture description" [kernel regex]"""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, capture = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else "score_cell_kernel")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if pat.search(r[rows[0].index("Kernel Name")])]
ix = {h: i for i, h in enumerate(hdr)}


def col(name, scale=1.0):
    vals = []
    for d in data:
        v = float(d[ix[name]].replace(",", ""))
        u = units[ix[name]]
        if u == "Mbyte":
            v *= 1e6
        elif u == "Kbyte":
            v *= 1e3
        elif u == "Gbyte":
            v *= 1e9
        elif u in ("us", "usecond"):
            v *= 1e-3
        elif u in ("ns", "nsecond"):
            v *= 1e-6
        vals.append(v * scale)
    return vals


def mean(v):
    return sum(v) / len(v)


names = [d[ix["Kernel Name"]] for d in data]
out = {
    "capture": capture,
    "kernels": names,
    "kernel_source_sha16": hashlib.sha256(open(os.path.join(ROOT, "misc3d_b200", "csrc", "score_cell.cuh"), "rb").read()).hexdigest()[:16],
    "duration_ms": mean(col("gpu__time_duration.sum")),
    "duration_ms_each": col("gpu__time_duration.sum"),
    "issue_slots_busy_frac": mean(col("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.01)),
    "fma_pipe_frac": mean(col("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0.01)),
    "alu_pipe_frac": mean(col("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0.01)),
    "smem_wavefront_frac": mean(col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 0.01)),
    "warps_active_frac": mean(col("sm__warps_active.avg.pct_of_peak_sustained_active", 0.01)),
    "warp_instructions_per_launch": mean(col("smsp__inst_executed.sum")),
    "dram_bytes_per_launch": mean([a + b for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]),
    "dram_throughput_frac": mean(col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.01)),
    "registers_per_thread": mean(col("launch__registers_per_thread")),
    "each": {"issue_slots_busy_frac": col("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.01),
             "warp_instructions": col("smsp__inst_executed.sum"),
             "smem_wavefronts": col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")},
}
json.dump(out, open(os.path.join(ROOT, "profiles", "score_kernel_ncu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
