#!/usr/bin/env python
"""Count the fp64-pipe instructions one active-set refinement / one ADMM iteration executes per instance, from the SASS of the
built library (no GPU needed), and write profiles/fp64_ops.json for bench.py's roofline.

    python tools/count_fp64_ops.py            # pendulum shape (4,1,20,20)

Method: in k_tpi_pol the innermost stage loops of the backward and the forward Riccati sweep are the two innermost loops with
the most fp64 instructions (the kernel holds two copies, phase A / phase B: the first is taken); one refinement runs each
Np times.  In k_tpi_admm the iteration body is the largest loop.  Thread-level counts (every lane of a warp is one instance).
DRAM bytes per launch come from the ncu captures named in the output (filled in by hand after a profile run)."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_cost as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pympc_b200", "libbmpc.so")
ks = sc.kernels(sc.disasm(lib))


def loops_of(name_part):
    name = [k for k in ks if name_part in k][0]
    ins = [sc.decode(i) for i in ks[name]]
    loops = []
    for i in ins:
        if i["op"] == "BRA":
            m = re.search(r"(0x[0-9a-f]+)\s*$", i["txt"])
            if m and int(m.group(1), 16) < i["addr"]:
                loops.append((int(m.group(1), 16), i["addr"]))
    loops = sorted(set(loops))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    return name, ins, loops, inner


Np = 20
name, ins, loops, inner = loops_of("k_tpi_polI8TpiShapeILi4ELi1ELi20ELi20EELb0")
cand = sorted(((sc.region_cost(ins, lo, hi)["fp64"], lo, hi) for lo, hi in inner), reverse=True)
cand = [c for c in cand if c[0] > 40]
cand.sort(key=lambda c: c[1])
first_two = sorted(cand[:2], reverse=True)                 # phase A copy: backward (more fp64), forward
bw, fw = first_two[0][0], first_two[1][0]
an, ains, aloops, ainner = loops_of("k_tpi_admmI8TpiShapeILi4ELi1ELi20ELi20EELb0")
admm = max(sc.region_cost(ains, lo, hi)["fp64"] for lo, hi in aloops)
out = {"pendulum_4_1_20_20": {
    "polish_kernel": "k_tpi_pol<TpiShape<4,1,20,20>>", "admm_kernel": "k_tpi_admm<TpiShape<4,1,20,20>>",
    "fp64_backward_per_stage": bw, "fp64_forward_per_stage": fw, "fp64_per_refinement": Np * (bw + fw),
    "fp64_per_admm_iteration": admm,
    "how": "tools/count_fp64_ops.py: DFMA+DADD+DMUL+DSETP in the stage loops of the compiled SASS x Np stages",
}}
# multi-input fast path (MIMO reference-governor shape, pattern-specialised instantiation): the backward and the forward stage loops
# of k_tpm_pol are its two loops with the most fp64 instructions (sub-steps are unrolled inside them)
try:
    mn, mins, mloops, minner = loops_of("k_tpm_polI14TpmSparseShapeILi8ELi4ELi40ELi40E")
    mc = sorted((sc.region_cost(mins, lo, hi)["fp64"] for lo, hi in mloops), reverse=True)
    # (the refinement loop encloses both sweeps: its count is their sum, skip it)
    sweeps = [c for c in mc if c < mc[0]][:2] if len(mc) > 2 and mc[0] >= mc[1] + mc[2] - 8 else mc[:2]
    out["mimo_8_4_40_40"] = {
        "polish_kernel": "k_tpm_pol<TpmSparseShape<8,4,40,40,...>>", "admm_kernel": "k_admm_tile<8,2,8,4>",
        "fp64_backward_per_stage": sweeps[0], "fp64_forward_per_stage": sweeps[1], "fp64_per_refinement": 40 * (sweeps[0] + sweeps[1]),
        "how": "tools/count_fp64_ops.py: DFMA+DADD+DMUL+DSETP in the two stage loops of the compiled SASS x Np stages (the straggler rounds' Schur-form "
               "refinements and the tile ADMM are not counted: their entries carry time shares only)",
    }
except Exception as exc:                                    # shape not in the table of this build
    print("mimo shape:", exc)
path = os.path.join(ROOT, "profiles", "fp64_ops.json")
old = {}
try:
    old = json.load(open(path))
except Exception:
    pass
for k, v in out.items():                                   # keep hand-entered DRAM figures
    for kk in ("polish_dram_bytes_per_launch", "admm_dram_bytes_per_launch", "dram_source"):
        if kk in old.get(k, {}):
            v[kk] = old[k][kk]
old.update(out)
json.dump(old, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
