"""ctypes front-end of ``osqp_port.c`` with the ``osqp`` 0.6.x Python surface pyMPC uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``OSQP().setup(P, q, A, l, u, **settings)``,
``.update(q=, l=, u=)``, ``.solve() -> res`` with ``res.x, res.y, res.info.status,
res.info.status_val, res.info.obj_val, res.info.iter`` — the members read at
``/root/reference/pyMPC/mpc.py:266,301-327,369-372,454``.  Injecting this module as
``sys.modules['osqp']`` lets the UNMODIFIED reference ``MPCController`` run end to end in a
container that has the reference but no ``osqp`` wheel (tests/refharness.py).

Setup-time work done here in numpy (not timed on the per-step path): Ruiz equilibration
(OSQP paper §5.1, 10 passes, as OSQP's default ``scaling=10``), KKT assembly, and a greedy
minimum-degree ordering standing in for AMD.
"""
import ctypes
import os
import subprocess
import types

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libosqp_port.so")
_lib = None

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4

STATUS_STR = {1: "solved", 2: "solved inaccurate", -2: "maximum iterations reached",
              -3: "primal infeasible", -4: "dual infeasible", -10: "unsolved"}


class _Settings(ctypes.Structure):
    _fields_ = [("rho", ctypes.c_double), ("sigma", ctypes.c_double), ("alpha", ctypes.c_double),
                ("eps_abs", ctypes.c_double), ("eps_rel", ctypes.c_double),
                ("eps_prim_inf", ctypes.c_double), ("eps_dual_inf", ctypes.c_double),
                ("max_iter", ctypes.c_int), ("check_termination", ctypes.c_int),
                ("adaptive_rho", ctypes.c_int), ("adaptive_rho_interval", ctypes.c_int),
                ("warm_start", ctypes.c_int), ("adaptive_rho_tolerance", ctypes.c_double)]


def build(force=False):
    """Compile the C restatement with gcc (called by ``__graft_entry__.build()`` and on first use)."""
    src = os.path.join(_HERE, "osqp_port.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", src,
                               "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        P = ctypes.c_void_p
        L.osqp_port_setup.restype = P
        L.osqp_port_setup.argtypes = [ctypes.c_int, ctypes.c_int] + [P] * 9 + [P, P, ctypes.c_double] + [P] * 5 + [P]
        L.osqp_port_update.argtypes = [P, P, P, P]
        L.osqp_port_warm_start.argtypes = [P, P, P]
        L.osqp_port_solve.argtypes = [P]; L.osqp_port_solve.restype = ctypes.c_int
        L.osqp_port_x.argtypes = [P]; L.osqp_port_x.restype = ctypes.POINTER(ctypes.c_double)
        L.osqp_port_y.argtypes = [P]; L.osqp_port_y.restype = ctypes.POINTER(ctypes.c_double)
        L.osqp_port_info.argtypes = [P] + [P] * 7
        L.osqp_port_clone.argtypes = [P]; L.osqp_port_clone.restype = P
        L.osqp_port_free.argtypes = [P]
        L.osqp_port_nnzL.argtypes = [P]; L.osqp_port_nnzL.restype = ctypes.c_int
        L.osqp_port_mpc_step_batch.restype = ctypes.c_int
        L.osqp_port_mpc_step_batch.argtypes = [P] + [ctypes.c_int] * 6 + [P] * 11 + [ctypes.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


def ruiz_scale(P, q, A, iters=10):
    """Returns scaled (P, q, A) and (D, E, c).  P is a full symmetric sparse matrix here."""
    n, m = P.shape[0], A.shape[0]
    D = np.ones(n); E = np.ones(m); c = 1.0
    P = sp.csc_matrix(P, dtype=float).copy(); A = sp.csc_matrix(A, dtype=float).copy(); q = np.array(q, float)
    for _ in range(iters):
        colP = np.abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
        colA = np.abs(A).max(axis=0).toarray().ravel() if A.nnz else np.zeros(n)
        rowA = np.abs(A).max(axis=1).toarray().ravel() if A.nnz else np.zeros(m)
        Dt = 1.0 / np.sqrt(_limit(np.maximum(colP, colA)))
        Et = 1.0 / np.sqrt(_limit(rowA))
        P = sp.diags(Dt) @ P @ sp.diags(Dt)
        A = sp.diags(Et) @ A @ sp.diags(Dt)
        q = Dt * q
        D *= Dt; E *= Et
        colP = np.abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
        ct = float(_limit(np.array([colP.mean()]))[0])
        nq = float(_limit(np.array([np.max(np.abs(q))]))[0])
        ct = 1.0 / max(ct, nq)
        P = P * ct; q = q * ct; c *= ct
    return sp.csc_matrix(P), q, sp.csc_matrix(A), D, E, c


def min_degree_order(K):
    """Greedy minimum-degree elimination order of the symmetric pattern K (stand-in for AMD)."""
    K = sp.csr_matrix(K)
    N = K.shape[0]
    adj = [set(K.indices[K.indptr[i]:K.indptr[i + 1]]) - {i} for i in range(N)]
    alive = np.ones(N, bool)
    import heapq
    heap = [(len(adj[i]), i) for i in range(N)]
    heapq.heapify(heap)
    order = []
    while heap:
        d, i = heapq.heappop(heap)
        if not alive[i] or d != len(adj[i]):
            continue
        alive[i] = False
        order.append(i)
        nb = adj[i]
        for j in nb:
            adj[j].discard(i)
        for j in nb:
            new = nb - adj[j] - {j}
            if new:
                adj[j] |= new
            heapq.heappush(heap, (len(adj[j]), j))
        adj[i] = set()
    return np.array(order, dtype=np.int32)


class _Info:
    pass


class OSQP:
    """Mirror of the ``osqp.OSQP`` object pyMPC drives (mpc.py:241)."""

    def __init__(self):
        self._w = None

    def __del__(self):
        if getattr(self, "_w", None):
            lib().osqp_port_free(self._w)
            self._w = None

    def setup(self, P=None, q=None, A=None, l=None, u=None, rho=0.1, sigma=1e-6, alpha=1.6,
              eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000,
              check_termination=25, adaptive_rho=True, adaptive_rho_interval=50,
              adaptive_rho_tolerance=5.0, warm_start=True, scaling=10, verbose=False, polish=False, **_ignored):
        if polish:
            raise NotImplementedError("osqp_port restates OSQP as pyMPC calls it (polish off, mpc.py:266)")
        P = sp.csc_matrix(P, dtype=float); A = sp.csc_matrix(A, dtype=float)
        n, m = P.shape[0], A.shape[0]
        Pfull = sp.triu(P, 1).T + sp.triu(P)                   # accept upper or full input like osqp does
        Pfull = sp.csc_matrix(Pfull)
        q = np.asarray(q, float).ravel()
        l = np.maximum(np.asarray(l, float).ravel(), -OSQP_INFTY)
        u = np.minimum(np.asarray(u, float).ravel(), OSQP_INFTY)
        if scaling:
            Ps, qs, As, D, E, c = ruiz_scale(Pfull, q, A, scaling)
        else:
            Ps, qs, As, D, E, c = Pfull, q.copy(), A.copy(), np.ones(n), np.ones(m), 1.0
        ls, us = E * l, E * u
        Pu = sp.csc_matrix(sp.triu(Ps)); Pu.sort_indices()
        As = sp.csc_matrix(As); As.sort_indices()
        # KKT = [[P + sigma I, A'], [A, -1/rho]]; the -1/rho diagonal is (re)written by the C side
        marker = -np.arange(1, m + 1, dtype=float) * 1e-3 - 7.0   # unique tags to locate the diagonal slots
        K = sp.bmat([[Ps + sigma * sp.eye(n), As.T], [As, sp.diags(marker)]], format="csc")
        perm = min_degree_order(K)
        Kp_ = sp.csc_matrix(K[perm][:, perm])
        Ku = sp.csc_matrix(sp.triu(Kp_)); Ku.sort_indices()
        inv = np.empty(n + m, np.int64); inv[perm] = np.arange(n + m)
        rho_idx = np.empty(m, np.int32)
        for i in range(m):
            col = inv[n + i]
            seg = slice(Ku.indptr[col], Ku.indptr[col + 1])
            k = np.nonzero(Ku.indices[seg] == col)[0]
            rho_idx[i] = Ku.indptr[col] + k[0]
            assert Ku.data[rho_idx[i]] == marker[i]
        self._keep = dict(
            Pp=Pu.indptr.astype(np.int32), Pi=Pu.indices.astype(np.int32), Px=Pu.data.astype(float),
            Ap=As.indptr.astype(np.int32), Ai=As.indices.astype(np.int32), Ax=As.data.astype(float),
            q=qs.astype(float), l=ls.astype(float), u=us.astype(float), D=D.astype(float), E=E.astype(float),
            Kp=Ku.indptr.astype(np.int32), Ki=Ku.indices.astype(np.int32), Kx=Ku.data.astype(float),
            perm=perm.astype(np.int32), rho_idx=rho_idx)
        k = self._keep
        st = _Settings(rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, int(max_iter),
                       int(check_termination), int(bool(adaptive_rho)), int(adaptive_rho_interval),
                       int(bool(warm_start)), float(adaptive_rho_tolerance))
        self.n, self.m = n, m
        self._w = lib().osqp_port_setup(n, m, _ptr(k["Pp"]), _ptr(k["Pi"]), _ptr(k["Px"]), _ptr(k["q"]),
                                        _ptr(k["Ap"]), _ptr(k["Ai"]), _ptr(k["Ax"]), _ptr(k["l"]), _ptr(k["u"]),
                                        _ptr(k["D"]), _ptr(k["E"]), c, _ptr(k["Kp"]), _ptr(k["Ki"]), _ptr(k["Kx"]),
                                        _ptr(k["perm"]), _ptr(k["rho_idx"]), ctypes.byref(st))
        if not self._w:
            raise ValueError("osqp_port: KKT factorisation failed")
        return self

    def update(self, q=None, l=None, u=None, **other):
        if other:
            raise NotImplementedError("osqp_port.update supports q, l, u (all pyMPC uses, mpc.py:454)")
        qa = None if q is None else np.ascontiguousarray(q, float)
        la = None if l is None else np.ascontiguousarray(l, float)
        ua = None if u is None else np.ascontiguousarray(u, float)
        lib().osqp_port_update(self._w, None if qa is None else _ptr(qa), None if la is None else _ptr(la),
                               None if ua is None else _ptr(ua))

    def warm_start(self, x=None, y=None):
        xa = None if x is None else np.ascontiguousarray(x, float)
        ya = None if y is None else np.ascontiguousarray(y, float)
        lib().osqp_port_warm_start(self._w, None if xa is None else _ptr(xa), None if ya is None else _ptr(ya))

    def solve(self):
        L = lib()
        L.osqp_port_solve(self._w)
        it, st, nr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        obj, pri, dua, rho = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        L.osqp_port_info(self._w, *[ctypes.byref(v) for v in (it, st, nr, obj, pri, dua, rho)])
        res = types.SimpleNamespace()
        res.x = np.ctypeslib.as_array(L.osqp_port_x(self._w), (self.n,)).copy()
        res.y = np.ctypeslib.as_array(L.osqp_port_y(self._w), (self.m,)).copy()
        info = _Info()
        info.iter, info.status_val, info.rho_updates = it.value, st.value, nr.value
        info.status = STATUS_STR.get(st.value, "unsolved")
        info.obj_val, info.pri_res, info.dua_res, info.rho_estimate = obj.value, pri.value, dua.value, rho.value
        res.info = info
        return res

    def nnz_L(self):
        return lib().osqp_port_nnzL(self._w)

    def clone_handle(self):
        return lib().osqp_port_clone(self._w)


class BatchCPU:
    """B independent solver objects stepped with OpenMP: the CPU baseline of bench.py.

    Restates the per-step work of ``MPCController.update()+solve()+output()`` for a constant
    ``xref`` (/root/reference/pyMPC/mpc.py:338-364, 386-454, 301-302) with the Python overhead
    removed, i.e. the most favourable reading of the reference's CPU path.
    """

    def __init__(self, qp, batch, **settings):
        """``qp``: an ``oracle.qp_assembly.QPData`` (soft constraints on)."""
        self.qp, self.B = qp, batch
        Pu, Ac = qp.to_csc()
        self.proto = OSQP().setup(Pu, qp.q, Ac, qp.l, qp.u, **settings)
        L = lib()
        self.handles = (ctypes.c_void_p * batch)(*[L.osqp_port_clone(self.proto._w) for _ in range(batch)])
        nx, nu, NX, NU = qp.nx, qp.nu, qp.NX, qp.NU
        # q_X = qx_coef @ xref ; q_U[:nu] += qdu @ um1 ; everything else constant
        self.qx_coef = np.ascontiguousarray(-np.vstack([qp.Qx] * qp.Np + [qp.QxN]))
        self.qdu = np.ascontiguousarray(-qp.QDu)
        qb = np.zeros(qp.n); qb[NX:NX + NU] = -np.kron(qp.w_u, qp.Qu @ qp.uref)
        self.q_base = qb
        l0, u0 = qp.bounds(np.zeros(nx), np.zeros(nu))
        self.l_base, self.u_base = np.ascontiguousarray(l0), np.ascontiguousarray(u0)
        self.nrow_du0 = 2 * NX + NU

    def step(self, X0, Um1, Xref, nthreads=None):
        qp, B = self.qp, self.B
        X0 = np.ascontiguousarray(X0, float); Um1 = np.ascontiguousarray(Um1, float); Xref = np.ascontiguousarray(Xref, float)
        U0 = np.empty((B, qp.nu)); status = np.empty(B, np.int32); iters = np.empty(B, np.int32)
        nthreads = nthreads or os.cpu_count()
        lib().osqp_port_mpc_step_batch(self.handles, B, qp.nx, qp.nu, qp.NX, qp.NU, self.nrow_du0,
                                       _ptr(self.qx_coef), _ptr(self.qdu), _ptr(self.q_base), _ptr(self.l_base),
                                       _ptr(self.u_base), _ptr(X0), _ptr(Um1), _ptr(Xref), _ptr(U0), _ptr(status),
                                       _ptr(iters), int(nthreads))
        return U0, status, iters

    def close(self):
        L = lib()
        for h in self.handles:
            L.osqp_port_free(h)
        self.handles = ()
