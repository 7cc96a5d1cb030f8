"""Second, algorithmically independent exact solver for the golden vectors (TEST INFRASTRUCTURE).

The goldens of tests/golden/ are produced by oracle/kkt.py::solve_exact (ADMM -> active set on the reference-form QP).
This module solves the same MPC problem by a different route, so that a systematic error of that family would show:

  1. the QP is condensed with explicit slack (variables [U ; E], E = slack of the soft state rows), which makes it a
     strictly convex QP with inequality constraints only;
  2. that QP is a least-distance problem (Lawson & Hanson, "Solving Least Squares Problems", ch. 23: LDP), which is solved
     through ONE non-negative least-squares problem — scipy.optimize.nnls, Lawson-Hanson's active-set NNLS, exact in finite
     steps: no iteration to a tolerance, no ADMM, no Schur-complement active-set polish.

Reference of the problem data: /root/reference/pyMPC/mpc.py:456-615 via oracle/qp_assembly.QPData."""
import numpy as np
from scipy.optimize import nnls


def solve_strictly_convex_qp(G, a, C, d):
    """min 1/2 x'Gx + a'x  s.t.  C x <= d   (G positive definite).  Returns x."""
    L = np.linalg.cholesky(G)
    c = np.linalg.solve(L, a)                       # y = L'x + c  ->  1/2 |y|^2 - 1/2 |c|^2
    M = np.linalg.solve(L, C.T).T                   # C L^-T
    Gm, h = -M, -(d + M @ c)                        # LDP: min |y| s.t. Gm y >= h
    n = G.shape[0]
    E = np.vstack([Gm.T, h[None, :]])
    f = np.zeros(n + 1); f[n] = 1.0
    u, _ = nnls(E, f, maxiter=50 * E.shape[1])
    r = E @ u - f
    if abs(r[n]) < 1e-14:
        raise ValueError("LDP: constraints incompatible")
    y = -r[:n] / r[n]
    return np.linalg.solve(L.T, y - c)


def solve_mpc(Q):
    """u-sequence (NU) of the MPC QP held by a oracle.qp_assembly.QPData (soft state rows with weight eps_feas)."""
    NX, NU = Q.NX, Q.NU
    Ax = Q.A[:NX, :NX]; Bu = Q.A[:NX, NX:NX + NU]
    Bcal = -np.linalg.solve(Ax, Bu)
    cc = np.linalg.solve(Ax, Q.l[:NX])             # dynamics rows: Ax X + Bu U = l  ->  X = cc + Bcal U
    PX = Q.P[:NX, :NX]; PU = Q.P[NX:NX + NU, NX:NX + NU]
    qX = Q.q[:NX]; qU = Q.q[NX:NX + NU]
    H = Bcal.T @ PX @ Bcal + PU
    g = Bcal.T @ (PX @ cc + qX) + qU
    rho_e = Q.P[NX + NU, NX + NU]
    n = NU + NX
    G = np.zeros((n, n)); G[:NU, :NU] = 0.5 * (H + H.T); G[NU:, NU:] = rho_e * np.eye(NX)
    a = np.concatenate([g, np.zeros(NX)])
    rows, rhs = [], []
    xl, xu = Q.l[NX:2 * NX], Q.u[NX:2 * NX]        # xmin <= X + E <= xmax
    S = np.hstack([Bcal, np.eye(NX)])
    for i in range(NX):
        if np.isfinite(xu[i]): rows.append(S[i]); rhs.append(xu[i] - cc[i])
        if np.isfinite(xl[i]): rows.append(-S[i]); rhs.append(-(xl[i] - cc[i]))
    Ah = Q.A[2 * NX:, NX:NX + NU]                  # input rows and the reference's delta-u rows (hard)
    hl, hu = Q.l[2 * NX:], Q.u[2 * NX:]
    for i in range(Ah.shape[0]):
        r = np.concatenate([Ah[i], np.zeros(NX)])
        if np.isfinite(hu[i]): rows.append(r); rhs.append(hu[i])
        if np.isfinite(hl[i]): rows.append(-r); rhs.append(-hl[i])
    x = solve_strictly_convex_qp(G, a, np.array(rows), np.array(rhs))
    return x[:NU]
