#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`) into a small markdown table for profiles/."""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem_wavefronts_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in ci]
    print("| # | kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, r in enumerate(data):
        name = r[ci["Kernel Name"]]
        m = re.search(r"conv_tc_kernel<(.*?)>", name)
        short = ("conv_tc<" + m.group(1).replace("(int)", "").replace("(bool)", "") + ">") if m else re.sub(r"\(.*", "", name).replace("void ", "").replace("ganrev::", "")[-48:]
        vals = []
        for mname, _ in cols:
            v = r[ci[mname]]
            u = units[ci[mname]]
            try:
                f = float(v.replace(",", ""))
                if u == "Gbyte": f *= 1000.0
                if u == "Kbyte": f /= 1000.0
                if u == "byte": f /= 1e6
                if u == "ms": f *= 1000.0
                if u == "ns": f /= 1000.0
                vals.append("%.1f" % f)
            except ValueError:
                vals.append(v)
        print(f"| {k} | {short} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
