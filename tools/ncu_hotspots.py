#!/usr/bin/env python
"""Aggregate the per-instruction stall samples of an ncu report (source page, SASS view) by code region.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ;  python tools/ncu_hotspots.py src.csv [lo:hi:name ...]

Regions are offsets from the kernel's first instruction (hex), e.g. 0x3910:0x51d0:backward.  Without regions: loops found
from backward branches.  Prints samples, share, executed instructions and the stall mix per region."""
import csv
import re
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
base = None
for r in rows[2:]:
    if len(r) < len(hdr) - 5:
        continue
    a = int(r[col["Address"]], 16)
    base = a if base is None else base
    data.append(dict(off=a - base, src=r[col["Source"]].strip(), n=int(r[col["# Samples"]] or 0), ex=int(r[col["Instructions Executed"]] or 0),
                     st={s: int(r[col[s]] or 0) for s in stall_cols}))
tot = sum(d["n"] for d in data)
regions = []
for a in [x for x in sys.argv[2:] if not x.startswith("--")]:
    lo, hi, name = a.split(":"); regions.append((int(lo, 16), int(hi, 16), name))
if not regions:
    for d in data:
        m = re.search(r"BRA(?:\.\w+)* .*?(0x[0-9a-f]+)\s*;?$", d["src"])
        if m:
            t = int(m.group(1), 16) - base                            # targets are absolute addresses
            if 0 <= t < d["off"] and d["off"] - t > 16 * 40:
                regions.append((t, d["off"], f"loop@{t:#x}"))
    # innermost loops only (plus everything else as "other")
    regions = [r for r in sorted(set(regions)) if not any(o != r and r[0] <= o[0] and o[1] <= r[1] for o in set(regions))]
print(f"total samples {tot}")
for lo, hi, name in regions:
    sel = [d for d in data if lo <= d["off"] <= hi]
    n = sum(d["n"] for d in sel); ex = sum(d["ex"] for d in sel)
    mix = collections.Counter()
    for d in sel:
        mix.update(d["st"])
    top = " ".join(f"{k[6:]}:{100 * v / max(1, n):.0f}%" for k, v in mix.most_common(7))
    print(f"{name:14s} [{lo:#x},{hi:#x}] samples {n} ({100 * n / tot:.1f}%) instr-exec {ex}  {top}")
covered = set()
for lo, hi, _ in regions:
    covered.update(d["off"] for d in data if lo <= d["off"] <= hi)
rest = [d for d in data if d["off"] not in covered]
print(f"{'other':14s} samples {sum(d['n'] for d in rest)} ({100 * sum(d['n'] for d in rest) / tot:.1f}%) instr-exec {sum(d['ex'] for d in rest)}")
if "--top" in sys.argv:
    for d in sorted(data, key=lambda d: -d["n"])[:40]:
        print(f"  {d['off']:#07x} n={d['n']:5d} ex={d['ex']:7d} {d['src'][:70]}  " + " ".join(f"{k[6:]}:{v}" for k, v in sorted(d['st'].items(), key=lambda kv: -kv[1])[:3] if v))
