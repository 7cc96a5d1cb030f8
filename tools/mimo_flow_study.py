"""Host emulation of the multi-input fast path's round loop (bmpc.cu::enqueue_round) on the MIMO side bench's workload, with the
device code compiled for the host: cold solve = [tile ADMM 25 -> polish from the iterate, all-at-once updates, cap 24] then straggler
rounds [ADMM chunk doubling the total -> polish in exchange mode, cap 24]; warm solve = [polish from the shifted sets, all-at-once,
cap 12] then straggler rounds [ADMM 100, 100, 200 -> polish from the iterate in exchange mode, cap 12].  Prints the distribution of
rounds per solve and how many instances would fall through to the Schur-form polish (after 200 iterations).
Usage: python tools/mimo_flow_study.py [instances] [warm steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
from emu import EmuSystem                       # noqa: E402
from pympc_b200.workloads import mimo           # noqa: E402

COLD_CHUNK, COLD_CAP, FIRST_CAP, ROUND_CAP, WARM_CHUNK = 25, 24, 12, 12, 100      # bmpc.cu: cold chunk, tpm_cold_cap, tpm_first_cap, tpm_round_cap, tpm_chunk


def cold_solve(E, cfg, x, um1):
    """returns (U or None when the instance would go on to the Schur-form polish, rounds, ADMM iterations)"""
    total = 0; chunk = COLD_CHUNK; r = 0
    while total + chunk <= 200:
        E.admm_only(x, um1, cfg["xref"], chunk); total += chunk
        vs = E.v.copy()
        if hasattr(E, "mcodes"):
            del E.mcodes
        U, ps = E.tpm_step(x, um1, cfg["xref"], mode=2, max_ref=COLD_CAP, exchange_from=-1 if r == 0 else 0)
        r += 1
        if ps > 0:
            return U, r, total
        E.v = vs; chunk = total
    return None, r, total


def warm_solve(E, cfg, x, um1, codes, plan):
    E.mcodes, E.Uplan = codes.copy(), plan.copy(); vprev = E.v.copy()
    U, ps = E.tpm_step(x, um1, cfg["xref"], mode=1, max_ref=FIRST_CAP)
    if ps > 0:
        return U, 1, 0
    E.v = vprev; E.x = plan.copy(); E.cold = 0; E.lvl = 2
    total = 0; chunk = WARM_CHUNK; r = 1
    while total + chunk <= 200:
        E.admm_only(x, um1, cfg["xref"], chunk); total += chunk
        vs = E.v.copy()
        U, ps = E.tpm_step(x, um1, cfg["xref"], mode=2, max_ref=ROUND_CAP, exchange_from=0)
        r += 1
        if ps > 0:
            return U, r, total
        E.v = vs; chunk = total
    return None, r, total


def run(n, steps, seed=4):
    cfg = mimo(); rng = np.random.default_rng(seed); X0 = 0.3 * rng.standard_normal((16384, 8))
    rc, rw, itc, itw = [], [], [], []; fall = 0
    for b in range(n):
        E = EmuSystem(cfg); x = X0[b].copy(); um1 = np.zeros(4)
        U, r, it = cold_solve(E, cfg, x, um1); rc.append(r); itc.append(it)
        if U is None:
            fall += 1; continue
        for t in range(steps):
            codes, plan = E.mcodes.copy(), U.copy()
            um1 = U[:4].copy(); x = cfg["Ad"] @ x + cfg["Bd"] @ um1
            U, r, it = warm_solve(E, cfg, x, um1, codes, plan); rw.append(r); itw.append(it)
            if U is None:
                fall += 1; break
    return dict(fallthrough=fall, cold_rounds=np.bincount(rc), warm_rounds=np.bincount(rw), cold_iters=float(np.mean(itc)),
                warm_iters=float(np.mean(itw)) if itw else 0.0)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    t0 = time.time(); r = run(n, steps)
    print(f"instances {n}: to the Schur-form fallback {r['fallthrough']} | cold solves by rounds {r['cold_rounds']}, {r['cold_iters']:.0f} ADMM iterations"
          f" | warm solves by rounds {r['warm_rounds']}, {r['warm_iters']:.1f} iterations  [{time.time() - t0:.0f} s]")
