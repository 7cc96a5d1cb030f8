"""Host-emulation study of the multi-input polish's working-set search (bmpc_tpm.cuh) on the MIMO side bench's transients:
closed loops from random x0, every warm step first tried from the previous working sets shifted one stage with `cap` refinements;
reports the share of warm solves verified within the cap and the mean refinements; then, for the stragglers of that first attempt,
what the device's straggler round does with them (ADMM chunk from the previous v*, then the polish from the iterate): all-at-once
updates capped at 4 (round 2's policy) against single exchanges capped at 12 (tpm_forward<S, true>).
Usage: python tools/tpm_search_study.py [n] [steps] [cap]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
from emu import EmuSystem                       # noqa: E402
from pympc_b200.workloads import mimo           # noqa: E402


POLICIES = (("all-at-once, cap 4", 4, -1), ("all-at-once, cap 12", 12, -1), ("single exchanges, cap 12", 12, 0))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    cfg = mimo(); rng = np.random.default_rng(4)
    ok = tot = 0; used = []; per_step = np.zeros((steps, 2), int)
    strag = {p[0]: [0, 0] for p in POLICIES}
    for b in range(n):
        E = EmuSystem(cfg); x = 0.3 * rng.standard_normal(8); um1 = np.zeros(4)
        for t in range(steps):
            if t == 0:
                Ut, st, *_ = E.solve(x, um1, cfg["xref"]); assert st == 1
                U, ps = E.tpm_step(x, um1, cfg["xref"], mode=2, max_ref=2); assert ps > 0
            else:
                E.mcodes, E.Uplan = codes.copy(), plan.copy(); vprev = E.v.copy()
                U, ps = E.tpm_step(x, um1, cfg["xref"], mode=1, max_ref=cap)
                tot += 1; ok += ps > 0; per_step[t] += (1, ps > 0)
                if ps > 0:
                    used.append(ps)
                else:
                    for name, c, xf in POLICIES:
                        E.v = vprev.copy(); E.x = plan.copy(); E.cold = 0; E.lvl = 2
                        E.admm_only(x, um1, cfg["xref"], 100)
                        U2, ps2 = E.tpm_step(x, um1, cfg["xref"], mode=2, max_ref=c, exchange_from=xf)
                        strag[name][0] += ps2 > 0; strag[name][1] += 1
                    E.v = vprev.copy(); E.x = plan.copy(); E.cold = 0
                    Ut, st, *_ = E.solve(x, um1, cfg["xref"]); assert st == 1
                    U, ps = E.tpm_step(x, um1, cfg["xref"], mode=2, max_ref=4); assert ps > 0, (b, t)
            codes, plan = E.mcodes.copy(), U.copy()
            x = cfg["Ad"] @ x + cfg["Bd"] @ U[:4]; um1 = U[:4].copy()
    print(f"cap {cap}: verified {ok}/{tot} = {100.0 * ok / tot:.1f} %, mean refinements of the verified {np.mean(used):.2f}")
    print("per step (tried, verified):", " ".join(f"{a}/{b}" for a, b in per_step[1:]))
    for name, (a, b) in strag.items():
        print(f"stragglers, first round [ADMM 100 -> {name}]: verified {a}/{b}")


if __name__ == "__main__":
    main()
