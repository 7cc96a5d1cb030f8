"""numpy restatement of the QP that ``pyMPC.mpc.MPCController`` assembles.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, in index form,
``/root/reference/pyMPC/mpc.py:456-615`` (``_compute_QP_matrices_``) and ``:386-454``
(``_update_QP_matrices_``); math in ``/root/reference/doc/latex/main.tex:129-499``.

Variables  z = [x_0..x_Np | u_0..u_{Nc-1} | eps_0..eps_Np]            (mpc.py:479,598)
Rows       [dynamics | x-bounds(+eps) | u-bounds | "delta-u" rows]     (mpc.py:592-594)

Everything is returned dense (the problems are small); ``to_csc`` converts for the C solver.
The reference's quirks are reproduced on purpose:
  * the delta-u block is ``-I + eye(k=1)`` on the *scalar* stacking of U with Nc*nu rows
    (mpc.py:569-571), i.e. it shifts by one scalar, not by nu;
  * Nc < Np repeats the last input (mpc.py:538-544) and weights Qu by (Np-Nc+1) on it
    (mpc.py:513-517).
"""
import numpy as np
import scipy.sparse as sp


def _dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M, dtype=float)


class QPData:
    """Dense (P, q, A, l, u) of one MPC instance + the pieces needed for per-step updates."""

    def __init__(self, Ad, Bd, Np=20, Nc=None, x0=None, xref=None, uref=None, uminus1=None,
                 Qx=None, QxN=None, Qu=None, QDu=None, xmin=None, xmax=None, umin=None,
                 umax=None, Dumin=None, Dumax=None, eps_feas=1e6, soft=True):
        Ad = _dense(Ad); Bd = _dense(Bd)
        nx, nu = Bd.shape
        Nc = Np if Nc is None else Nc
        self.nx, self.nu, self.Np, self.Nc = nx, nu, Np, Nc
        self.Ad, self.Bd = Ad, Bd
        z = np.zeros
        self.Qx = z((nx, nx)) if Qx is None else _dense(Qx)          # mpc.py:150 (zeros, not eye)
        self.QxN = self.Qx if QxN is None else _dense(QxN)            # mpc.py:158
        self.Qu = z((nu, nu)) if Qu is None else _dense(Qu)
        self.QDu = z((nu, nu)) if QDu is None else _dense(QDu)
        self.x0 = z(nx) if x0 is None else np.asarray(x0, float).ravel()
        self.xref = z(nx) if xref is None else np.asarray(xref, float)
        self.uref = z(nu) if uref is None else np.asarray(uref, float).ravel()
        self.uminus1 = self.uref.copy() if uminus1 is None else np.asarray(uminus1, float).ravel()
        inf = np.inf
        self.xmin = -inf * np.ones(nx) if xmin is None else np.asarray(xmin, float).ravel()
        self.xmax = inf * np.ones(nx) if xmax is None else np.asarray(xmax, float).ravel()
        self.umin = -inf * np.ones(nu) if umin is None else np.asarray(umin, float).ravel()
        self.umax = inf * np.ones(nu) if umax is None else np.asarray(umax, float).ravel()
        self.Dumin = -inf * np.ones(nu) if Dumin is None else np.asarray(Dumin, float).ravel()
        self.Dumax = inf * np.ones(nu) if Dumax is None else np.asarray(Dumax, float).ravel()
        self.eps_feas = float(eps_feas)
        self.soft = soft
        self.NX = (Np + 1) * nx
        self.NU = Nc * nu
        self.n = self.NX + self.NU + (self.NX if soft else 0)
        self.m = self.NX + self.NX + self.NU + (Nc + 1) * nu
        self._build_static()
        self.q = self.linear_term(self.xref, self.uminus1)
        self.l, self.u = self.bounds(self.x0, self.uminus1)

    # ---- pieces that never change after setup (mpc.py:482-487,510-524,531,537-571) ----
    def _build_static(self):
        nx, nu, Np, Nc, NX, NU = self.nx, self.nu, self.Np, self.Nc, self.NX, self.NU
        n, m = self.n, self.m
        P = np.zeros((n, n))
        for k in range(Np):                                   # stage weights x_0..x_{Np-1}
            P[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = self.Qx
        P[Np * nx:NX, Np * nx:NX] = self.QxN                  # terminal weight
        self.P_X = P[:NX, :NX].copy()
        w = np.ones(Nc); w[Nc - 1] = Np - Nc + 1              # Qu multiplicity of the held input
        self.w_u = w
        T = 2 * np.eye(Nc) - np.eye(Nc, k=1) - np.eye(Nc, k=-1)
        T[Nc - 1, Nc - 1] = 1
        P[NX:NX + NU, NX:NX + NU] = np.kron(np.diag(w), self.Qu) + np.kron(T, self.QDu)
        if self.soft:
            P[NX + NU:, NX + NU:] = self.eps_feas * np.eye(NX)
        self.P = P
        A = np.zeros((m, n))
        for k in range(Np + 1):                               # dynamics rows
            A[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = -np.eye(nx)
            if k >= 1:
                A[k * nx:(k + 1) * nx, (k - 1) * nx:k * nx] = self.Ad
                j = min(k - 1, Nc - 1)
                A[k * nx:(k + 1) * nx, NX + j * nu:NX + (j + 1) * nu] = self.Bd
        r = NX
        A[r:r + NX, :NX] = np.eye(NX)                         # x (+eps) bounds
        if self.soft:
            A[r:r + NX, NX + NU:] = np.eye(NX)
        r += NX
        A[r:r + NU, NX:NX + NU] = np.eye(NU)                  # u bounds
        r += NU
        A[r:r + nu, NX:NX + nu] = np.eye(nu)                  # u_0 - u_{-1}
        r += nu
        A[r:r + NU, NX:NX + NU] = -np.eye(NU) + np.eye(NU, k=1)   # scalar-shift quirk
        self.A = A

    # ---- per-step vectors (mpc.py:404-452, 489-526, 551-580) ----
    def linear_term(self, xref, uminus1):
        nx, nu, Np, Nc = self.nx, self.nu, self.Np, self.Nc
        xref = np.asarray(xref, float)
        if xref.ndim == 2 and xref.shape[0] >= Np + 1:
            qX = -(xref.reshape(1, -1) @ self.P_X).ravel()    # needs exactly Np+1 rows, as the reference
        else:
            qX = -np.hstack([np.kron(np.ones(Np), self.Qx @ xref), self.QxN @ xref])
        qU = -np.kron(self.w_u, self.Qu @ self.uref)
        qU[:nu] += -(self.QDu @ np.asarray(uminus1, float).ravel())
        q = np.hstack([qX, qU])
        if self.soft:
            q = np.hstack([q, np.zeros(self.NX)])
        return q

    def bounds(self, x0, uminus1):
        nu, Np, Nc = self.nu, self.Np, self.Nc
        um1 = np.asarray(uminus1, float).ravel()
        leq = np.hstack([-np.asarray(x0, float).ravel(), np.zeros(self.Np * self.nx)])
        ldu = np.kron(np.ones(Nc + 1), self.Dumin); ldu[:nu] += um1[:nu]
        udu = np.kron(np.ones(Nc + 1), self.Dumax); udu[:nu] += um1[:nu]
        l = np.hstack([leq, np.kron(np.ones(Np + 1), self.xmin), np.kron(np.ones(Nc), self.umin), ldu])
        u = np.hstack([leq, np.kron(np.ones(Np + 1), self.xmax), np.kron(np.ones(Nc), self.umax), udu])
        return l, u

    def update(self, x0, uminus1=None, xref=None):
        """Restates ``update()`` + ``_update_QP_matrices_`` (mpc.py:338-364, 386-454)."""
        if uminus1 is not None:
            self.uminus1 = np.asarray(uminus1, float).ravel()
        if xref is not None:
            self.xref = np.asarray(xref, float)
        self.x0 = np.asarray(x0, float).ravel()
        self.q = self.linear_term(self.xref, self.uminus1)
        self.l, self.u = self.bounds(self.x0, self.uminus1)

    def constant_term(self, compute_jx=False):
        """J_CNST exactly as the reference accumulates it (mpc.py:412-442; quirk Q7)."""
        J = 0.0
        xref = self.xref
        if compute_jx:
            if xref.ndim == 2 and xref.shape[0] >= self.Np + 1:
                qX = -(xref.reshape(1, -1) @ self.P_X).ravel()
                J += -0.5 * qX @ xref.ravel()
            else:
                J += 0.5 * self.Np * (xref @ (self.QxN @ xref)) + 0.5 * xref @ (self.QxN @ xref)
        J += 0.5 * self.Np * (self.uref @ (self.Qu @ self.uref))
        J += 0.5 * self.uminus1 @ (self.QDu @ self.uminus1)
        return J

    def to_csc(self):
        """(P upper-triangular CSC, A CSC) for the C solver."""
        return sp.csc_matrix(np.triu(self.P)), sp.csc_matrix(self.A)

    def u0_slice(self):
        return slice(self.NX, self.NX + self.nu)                # mpc.py:302
