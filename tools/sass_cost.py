#!/usr/bin/env python
"""Static cost model of a compiled kernel from its SASS (no GPU needed).

    python tools/sass_cost.py <lib.so> <kernel-name-substring> [--loops] [--range lo hi]

For every loop (backward branch) of the kernel — or an explicit address range — prints the instruction count, the
number of fp64-pipe instructions (DFMA/DADD/DMUL/DSETP: 2 issue cycles each on B200, 0.5 warp-instr/clk/SMSP), and the
sum of the per-instruction stall counts from the control words (bits [105:109) of each 128-bit instruction: the cycles
ONE warp needs to walk the region when every scoreboard wait is already satisfied).  sum(stall) / (2 * n_fp64) is the
number of co-resident warps per scheduler needed to saturate the fp64 pipe if nothing else stalls: the closer to 1,
the better the instruction-level parallelism of the schedule.  A planning tool: real numbers come from ncu.
"""
import re
import subprocess
import sys
import collections

FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def disasm(lib):
    return subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout


def kernels(text):
    out, name, cur = {}, None, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); cur = out.setdefault(name, []); continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
        if m:
            cur.append([int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]); continue
        m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
        if m and cur and cur[-1][3] is None:
            cur[-1][3] = int(m.group(1), 16)
    return out


def decode(ins):
    addr, txt, lo, hi = ins
    hi = hi or 0
    toks = txt.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ctrl = hi >> 41
    return dict(addr=addr, op=op.split(".")[0], full=op, txt=txt, stall=ctrl & 0xf, yld=(ctrl >> 4) & 1,
                wbar=(ctrl >> 5) & 7, rbar=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3f)


def region_cost(ins, lo, hi):
    sel = [i for i in ins if lo <= i["addr"] <= hi]
    n = len(sel); nfp = sum(i["op"] in FP64 for i in sel); st = sum(max(1, i["stall"]) for i in sel)
    waits = sum(1 for i in sel if i["wait"])
    hist = collections.Counter(i["op"] for i in sel)
    return dict(n=n, fp64=nfp, stall=st, waits=waits, hist=hist)


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    ks = kernels(disasm(lib))
    for name, raw in ks.items():
        if pat not in name:
            continue
        ins = [decode(i) for i in raw]
        print(f"== {name}: {len(ins)} instructions, {len(ins) * 16 / 1024:.1f} KB")
        if "--range" in sys.argv:
            k = sys.argv.index("--range"); ranges = [(int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16))]
        else:
            ranges = []
            for i in ins:
                if i["op"] == "BRA":
                    m = re.search(r"(0x[0-9a-f]+)\s*$", i["txt"])
                    if m and int(m.group(1), 16) < i["addr"]:
                        ranges.append((int(m.group(1), 16), i["addr"]))
        whole = region_cost(ins, 0, 1 << 40)
        print(f"   whole kernel: n={whole['n']} fp64={whole['fp64']} sum_stall={whole['stall']}")
        for lo, hi in sorted(set(ranges)):
            c = region_cost(ins, lo, hi)
            if c["n"] < 24:
                continue
            need = c["stall"] / max(1, 2 * c["fp64"])
            top = " ".join(f"{k}:{v}" for k, v in c["hist"].most_common(9))
            print(f"   loop {lo:#x}-{hi:#x}: n={c['n']} ({c['n'] * 16 / 1024:.1f} KB) fp64={c['fp64']} sum_stall={c['stall']} "
                  f"sb_waits={c['waits']} warps_to_saturate_fp64={need:.2f}   [{top}]")


if __name__ == "__main__":
    main()
