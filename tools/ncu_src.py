#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per kernel, stall mix and hottest instructions."""
import csv
import sys

def main(path, topn=30, want=None):
    rows = list(csv.reader(open(path)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for k, s in enumerate(secs):
        if want is not None and k != want:
            continue
        hdr, data = s["hdr"], s["data"]
        ci = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ci["# Samples"]]) for r in data) or 1
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(int(r[ci[c]]) for r in data), c[6:]) for c in stall_cols), reverse=True)[:8]
        print(f"== [{k}] {s['name'][:110]}\n   samples {tot}, instrs {len(data)}, stalls {[(n, round(100*v/tot,1)) for v, n in agg]}")
        top = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]]))[:topn]
        for idx in sorted(top):
            r = data[idx]
            sm = int(r[ci["# Samples"]])
            st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
            print(f"   {idx:5d} {r[ci['Source']].strip()[:66]:66s} {sm:6d} {100*sm/tot:5.1f}% x{r[ci['Instructions Executed']]:>8s} {st}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, int(sys.argv[3]) if len(sys.argv) > 3 else None)
