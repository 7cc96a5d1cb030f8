"""Solver-independent KKT certificate + an exact dense solver for  min ½zᵀPz+qᵀz  s.t. l ≤ Az ≤ u.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used to (a) generate the committed golden
vectors (tests/golden/make_golden.py) and (b) certify any candidate solution on the
reference-assembled (P, q, A, l, u) of ``/root/reference/pyMPC/mpc.py:597-606`` without
trusting a particular solver: a point that satisfies the KKT conditions of a convex QP is a
global minimiser, and it is unique in the u-block whenever P_U ≻ 0.

The exact solver is a plain dense ADMM warm-up (OSQP paper, Alg. 1, unscaled, fixed rho)
followed by a primal-dual active-set iteration; it is deliberately *not* the algorithm the
CUDA path uses (that one works on the condensed form), so agreement between the two is
evidence, not tautology.
"""
import numpy as np
import scipy.linalg as sla

INF_BOUND = 1e20


def kkt_residuals(P, q, A, l, u, z, y):
    """Returns dict(stat, prim, comp, sign): ∞-norm KKT residuals of (z, y)."""
    Az = A @ z
    stat = np.max(np.abs(P @ z + q + A.T @ y))
    prim = max(0.0, np.max(l - Az), np.max(Az - u))
    yp, ym = np.maximum(y, 0), np.minimum(y, 0)
    gap_u = np.where(np.isfinite(u), np.abs(u - Az), np.inf)
    gap_l = np.where(np.isfinite(l), np.abs(Az - l), np.inf)
    comp = max(np.max(np.minimum(yp, gap_u)), np.max(np.minimum(-ym, gap_l)))
    return dict(stat=float(stat), prim=float(prim), comp=float(comp))


def _pdas(P, q, A, l, u, lo_act, up_act, max_steps=50, delta=1e-10, tol=1e-9):
    n, m = P.shape[0], A.shape[0]
    eq = np.isfinite(l) & (l == u)
    for _ in range(max_steps):
        up = up_act | eq
        lo = lo_act & ~up
        act = up | lo
        b = np.where(up, u, l)[act]
        Aa = A[act]
        na = Aa.shape[0]
        KK = np.zeros((n + na, n + na))
        KK[:n, :n] = P; KK[:n, n:] = Aa.T; KK[n:, :n] = Aa
        Kr = KK.copy()
        Kr[np.arange(n), np.arange(n)] += delta
        Kr[np.arange(n, n + na), np.arange(n, n + na)] -= delta
        rhs = np.hstack([-q, b])
        try:
            lu = sla.lu_factor(Kr)
            sol = sla.lu_solve(lu, rhs)
            for _r in range(8):
                sol = sol + sla.lu_solve(lu, rhs - KK @ sol)
        except (np.linalg.LinAlgError, ValueError):
            return None
        if not np.all(np.isfinite(sol)):
            return None
        z = sol[:n]
        y = np.zeros(m); y[act] = sol[n:]
        Az = A @ z
        sc = tol * (1 + np.abs(Az))
        viol_u = Az > u + sc
        viol_l = Az < l - sc
        bad_u = up & ~eq & (y < -tol * (1 + np.abs(y)))
        bad_l = lo & (y > tol * (1 + np.abs(y)))
        if not (viol_u.any() or viol_l.any() or bad_u.any() or bad_l.any()):
            return z, y
        up_act = ((up & ~eq & ~bad_u) | viol_u)
        lo_act = ((lo & ~bad_l) | viol_l) & ~up_act
    return None


def solve_exact(P, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, max_rounds=200, warm=None, tol=1e-9):
    """Exact minimiser (z, y, residuals) with KKT residuals < ~1e-9, or raises RuntimeError."""
    P = np.asarray(P, float); A = np.asarray(A, float)
    l = np.where(np.asarray(l) < -INF_BOUND, -np.inf, l).astype(float)
    u = np.where(np.asarray(u) > INF_BOUND, np.inf, u).astype(float)
    n, m = P.shape[0], A.shape[0]
    eq = np.isfinite(l) & (l == u)
    free = ~np.isfinite(l) & ~np.isfinite(u)

    def make(rho):
        rv = np.where(eq, 1e3 * rho, rho)
        rv = np.where(free, 1e-6, rv)
        return rv, np.linalg.inv(P + sigma * np.eye(n) + A.T @ (rv[:, None] * A))

    rv, Kinv = make(rho)
    x = np.zeros(n) if warm is None else warm[0].copy()
    y = np.zeros(m) if warm is None else warm[1].copy()
    zc = np.clip(A @ x, l, u)
    for rnd in range(max_rounds):
        for _ in range(50):
            xt = Kinv @ (sigma * x - q + A.T @ (rv * zc - y))
            zt = A @ xt
            x = alpha * xt + (1 - alpha) * x
            zr = alpha * zt + (1 - alpha) * zc
            zn = np.clip(zr + y / rv, l, u)
            y = y + rv * (zr - zn)
            zc = zn
        Ax = A @ x
        rp = np.max(np.abs(Ax - zc)) / max(np.max(np.abs(Ax)), np.max(np.abs(zc)), 1e-12)
        Px, Aty = P @ x, A.T @ y
        rd = np.max(np.abs(Px + q + Aty)) / max(np.max(np.abs(Px)), np.max(np.abs(Aty)), np.max(np.abs(q)), 1e-12)
        if max(rp, rd) < 1e-2:
            # OSQP's polish guess (OSQP paper §4): row active iff the dual pushes past the bound
            lo_act = (zc - l < -y) & ~eq
            up_act = (u - zc < y) & ~eq
            out = _pdas(P, q, A, l, u, lo_act, up_act, max_steps=6, tol=tol)
            if out is not None:
                z, yy = out
                r = kkt_residuals(P, q, A, l, u, z, yy)
                if max(r.values()) < 1e-7:
                    return z, yy, r
        # OSQP's adaptive-rho rule (OSQP paper §5.2), applied once per round
        rho_new = float(np.clip(rho * np.sqrt(rp / max(rd, 1e-300)), 1e-6, 1e6))
        if rho_new > 2 * rho or rho_new < rho / 2:
            rho = rho_new
            rv, Kinv = make(rho)
    raise RuntimeError("solve_exact: no KKT-certified solution found")
