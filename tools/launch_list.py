#!/usr/bin/env python
"""Summarise the `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table for profiles/."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("(int)", "").replace("(bool)", "")
    name = name.split(">(")[0] + ">" if ">(" in name else re.sub(r"\(.*$", "", name)   # drop the argument list
    name = name.replace("void ", "").replace("stc::", "search_tc::").replace("ktc::", "kmeans_tc::").replace("ltc::", "label_tc::").replace("ganrev::", "").replace("tc::", "").replace("scan::", "").replace("conv_tc_kernel", "conv_tc")
    return name


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in data:
        k = short(r[ci["Kernel Name"]])
        e = agg.setdefault(k, {"n": 0, "ns": 0.0, "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]]})
        e["n"] += 1
        e["ns"] += float(r[ci["Metric Value"]])
    tot = sum(e["ns"] for e in agg.values()) or 1.0
    print("| kernel (conv_tc template args: NT, MT, NDY, BRES, ACT, POOL, FP32OUT, CG) | launches | total us | share | grid | block |")
    print("|---|---|---|---|---|---|")
    for k, e in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"| {k} | {e['n']} | {e['ns'] / 1e3:.1f} | {e['ns'] / tot:.3f} | {e['grid']} | {e['block']} |")


if __name__ == "__main__":
    main(sys.argv[1])
