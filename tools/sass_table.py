"""SASS mnemonic counts per kernel of the built library: python tools/sass_table.py > profiles/rNN/sass_table.md"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "gan-reverser_b200", "libganrev_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "FFMA2", "FFMA", "MUFU", "F2I"]
rows, cur = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ganrev::", "").replace("(int)", "").replace("(bool)", "")
        cur = rows.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        for c in COLS:
            if op == c or (c != "FFMA" and op.startswith(c)):
                cur[c] += 1
                break
print("# SASS evidence (round 2, final library)\n")
print("`cuobjdump -sass gan-reverser_b200/libganrev_cuda.so`, mnemonic counts per kernel (`tools/sass_table.py`). tcgen05.mma = `UTCHMMA`, tcgen05.commit = `UTCBAR`, "
      "tcgen05.ld = `LDTM`, TMA load / store = `UTMALDG` / `UTMASTG`, cp.async = `LDGSTS`, fma.rn.f32x2 = `FFMA2`; the legacy tensor path `HMMA` (mma.sync) must be 0.\n")
print("| kernel | " + " | ".join(COLS) + " |")
print("|---|" + "---|" * len(COLS))
tot = collections.Counter()
for k, c in rows.items():
    tot.update(c)
    print(f"| {k} | " + " | ".join(str(c[x]) for x in COLS) + " |")
print(f"| **total** | " + " | ".join(str(tot[x]) for x in COLS) + " |")
