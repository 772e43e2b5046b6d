#!/usr/bin/env python
"""Torch7 `.net` checkpoint -> raw float32 weight blob (+ .json geometry) for the Lua shim.

  python tools/net2blob.py logs/adversarial.net            -> adversarial.G.blob (+ .json)
  python tools/net2blob.py logs/r_32x32_nd100_normal.net --R 1,32,32,100   -> *.R.blob

`lua/ganrev.lua` read_blob() maps the file with torch.FloatStorage, so the Lua process needs neither nn nor
cudnn to deserialise a trained model (apply_r.lua:62-69, 92-103; north_star: "no cutorch/cudnn").  Host-only:
the parser is gan-reverser_b200/t7.py.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("net")
    ap.add_argument("--R", default="", help="C,H,W,noiseDim of an R / R_fixer checkpoint (a G checkpoint carries its own opt)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    t7 = load_package().t7
    ckpt = t7.load(a.net)
    stem = a.out or os.path.splitext(a.net)[0]
    if a.R:
        C, H, W, nd = (int(v) for v in a.R.split(","))
        blob, kind = t7.r_blob(ckpt, C, H, W, nd), "R"
    else:
        C, H, W, nd, blob = t7.g_blob(ckpt)
        kind = "G"
    path = f"{stem}.{kind}.blob"
    blob.astype("<f4").tofile(path)
    json.dump({"kind": kind, "C": C, "H": H, "W": W, "noiseDim": nd, "floats": int(blob.size)}, open(path + ".json", "w"))
    print(path, blob.size, "floats")


if __name__ == "__main__":
    main()
