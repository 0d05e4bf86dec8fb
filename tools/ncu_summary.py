#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import re
import subprocess
import sys
import collections

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = hdr.index("Kernel Name")
    for r in data:
        print(f"kernel: {r[name]}")
        for k in KEYS + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]:
            if k in hdr:
                i = hdr.index(k)
                if r[i] not in ("0", ""):
                    print(f"  {k:88s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr, data = rows[1], rows[2:]
        ix = {k: i for i, k in enumerate(hdr)}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
        tot = sum(f(r, "Instructions Executed") for r in data) or 1.0
        for r in data:
            m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
            a = agg[m.group(2) if m else "?"]
            a[0] += f(r, "Instructions Executed")
            a[1] += f(r, "L1 Wavefronts Shared")
            a[2] += f(r, "L1 Wavefronts Shared Ideal")
        print("  SASS opcode mix (warp instructions executed, shared-memory wavefronts actual/ideal):")
        for op, a in sorted(agg.items(), key=lambda x: -x[1][0])[:12]:
            print(f"    {op:12s} {a[0]:14.0f} {100 * a[0] / tot:5.1f}%   wf {a[1]:12.0f} / {a[2]:12.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
