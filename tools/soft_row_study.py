"""Host-emulation study of the slow soft-row regime: random small systems, x0 up to `scale` x OUTSIDE the
soft state box, large eps_feas.  Runs the generic (team) core of the device code compiled for the host (tests/hostemu) and
tallies verified (1) / solved-unpolished (2) / max-iter (-2) plus the distance of every answer from the oracle's exact
minimiser.  Usage: python tools/soft_row_study.py [n_systems] [eps_feas] [scale] [nx,nu,Np]
(shape: random small shapes by default; "4,1,20" = pendulum-size problems, polish capacity as on the device).  With OSQP=1 in the
environment the oracle's OSQP restatement (the reference's own solver path, eps 1e-3, max_iter 4000) runs on the same QPs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
from emu import EmuSystem                       # noqa: E402
from oracle.qp_assembly import QPData           # noqa: E402
from oracle.kkt import solve_exact              # noqa: E402


def random_system(rng, nx, nu, Np, Nc, eps_feas):
    A = rng.standard_normal((nx, nx)); A *= min(1.0, 1.05 / max(abs(np.linalg.eigvals(A))))
    return dict(Ad=A, Bd=rng.standard_normal((nx, nu)), Np=Np, Nc=Nc or Np, Qx=np.diag(rng.uniform(0.1, 2.0, nx)),
                QxN=np.diag(rng.uniform(0.1, 2.0, nx)), Qu=np.diag(rng.uniform(0.0, 0.5, nu)), QDu=np.diag(rng.uniform(0.05, 1.0, nu)),
                xmin=-rng.uniform(0.5, 3.0, nx), xmax=rng.uniform(0.5, 3.0, nx), umin=-rng.uniform(0.3, 2.0, nu),
                umax=rng.uniform(0.3, 2.0, nu), Dumin=-rng.uniform(0.2, 1.0, nu), Dumax=rng.uniform(0.2, 1.0, nu), eps_feas=eps_feas,
                uref=np.zeros(nu))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    eps_feas = float(sys.argv[2]) if len(sys.argv) > 2 else 1e5
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5
    shape = tuple(int(v) for v in sys.argv[4].split(",")) if len(sys.argv) > 4 else None
    rng = np.random.default_rng(2024)
    tally = {1: 0, 2: 0, -2: 0}; err2 = []; errf = []; its = []; osqp_rows = []
    t0 = time.time()
    for i in range(n):
        nx, nu, Np = shape if shape else (int(rng.integers(2, 5)), int(rng.integers(1, 3)), int(rng.integers(4, 10)))
        c = random_system(rng, nx, nu, Np, None, eps_feas)
        c["x0"] = rng.uniform(scale * c["xmin"], scale * c["xmax"]); c["xref"] = 0.5 * rng.standard_normal(nx); c["uminus1"] = np.zeros(nu)
        E = EmuSystem(c)
        U, st, it, ps, res = E.solve(c["x0"], c["uminus1"], c["xref"], rmax=min(E.mc, 128), pdas_steps=int(os.environ.get("PDAS", 10)),
                                      first_iters=50 if shape else 10)
        Q = QPData(**c); tol = 1e-6
        try:
            z, y, r = solve_exact(Q.P, Q.q, Q.A, Q.l, Q.u); ref = z[Q.NX:Q.NX + Q.NU]
        except RuntimeError:                    # the ADMM -> active-set oracle cannot certify some of the stiffest QPs: the independent exact solver takes over
            from oracle.ldp import solve_mpc
            ref = solve_mpc(Q); tol = 1e-4      # (to NNLS accuracy)
        e = np.max(np.abs(U - ref)) / (1 + np.max(np.abs(ref)))
        if os.environ.get("OSQP"):
            import scipy.sparse as sp
            from oracle import osqp_port
            S = osqp_port.OSQP(); S.setup(sp.csc_matrix(Q.P), Q.q, sp.csc_matrix(Q.A), Q.l, Q.u, warm_start=True, verbose=False, eps_abs=1e-3, eps_rel=1e-3)
            ro = S.solve()
            osqp_rows.append((ro.info.status_val, ro.info.iter, np.max(np.abs(ro.x[Q.NX:Q.NX + Q.NU] - ref)) / (1 + np.max(np.abs(ref)))))
        tally[st if st in tally else -2] += 1; its.append(it)
        if st == 1:
            assert e < tol, (i, e)
        elif st == 2:
            err2.append(e)
        else:
            errf.append(e)
    if osqp_rows:
        st = [r[0] for r in osqp_rows]; ok = [r for r in osqp_rows if r[0] == 1]
        print(f"OSQP restatement on the same QPs: solved {len(ok)}, max-iter {sum(1 for v in st if v == -2)}, other {sum(1 for v in st if v not in (1, -2))};"
              f" mean iters {np.mean([r[1] for r in osqp_rows]):.0f}; relative error of its 'solved' answers median {np.median([r[2] for r in ok]):.1e} max {max(r[2] for r in ok):.1e}")
    print(f"n={n} eps_feas={eps_feas:g} scale={scale}: verified {tally[1]}, solved-unpolished {tally[2]}, max-iter {tally[-2]};"
          f" mean iters {np.mean(its):.0f}; status-2 err max {max(err2, default=0):.2e} median {np.median(err2) if err2 else 0:.2e};"
          f" failed err max {max(errf, default=0):.2e}  [{time.time() - t0:.0f} s]")


if __name__ == "__main__":
    main()
