"""Generates the committed golden fixtures (run HERE, where /root/reference exists).

For every case the QP is assembled by the UNMODIFIED reference (pyMPC.mpc.MPCController with a
stub ``osqp`` module, tests/refharness.py), solved by ``oracle.kkt.solve_exact`` and certified by
the solver-independent KKT residuals (< 1e-9) on the reference-assembled (P, q, A, l, u).
Closed loops use the linear plant x+ = Ad x + Bd u (README.md:59-77 pattern).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from refharness import load_reference_controller  # noqa: E402
from oracle.kkt import solve_exact, kkt_residuals  # noqa: E402
from pympc_b200.workloads import point_mass, pendulum, mimo, pendulum_random  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MPC = load_reference_controller()


def exact(K, warm=None):
    P, A = K.P.toarray(), K.A.toarray()
    z, y, r = solve_exact(P, K.q, A, K.l, K.u, warm=warm)
    assert max(r.values()) < 1e-9, r
    return z, y, r


def first_solve(name, cfg):
    K = MPC(**cfg); K.setup(solve=False)
    z, y, r = exact(K)
    NX = (K.Np + 1) * K.nx; NU = K.Nc * K.nu
    obj = 0.5 * z @ (K.P @ z) + K.q @ z
    np.savez_compressed(os.path.join(OUT, f"{name}_first.npz"), z=z, y=y, q=K.q, l=K.l, u=K.u,
                        P_data=K.P.toarray()[np.nonzero(K.P.toarray())], A_sum=np.array([K.A.toarray().sum(), np.abs(K.A.toarray()).sum()]),
                        u_seq=z[NX:NX + NU], x_seq=z[:NX], eps_seq=z[NX + NU:], obj=obj, J_CNST=K.J_CNST,
                        kkt=np.array([r['stat'], r['prim'], r['comp']]))
    print(name, "first u0", z[NX:NX + K.nu], r)


def closed_loop(name, cfg, nsteps, xref_fn=None):
    K = MPC(**cfg); K.setup(solve=False)
    Ad, Bd = np.asarray(cfg['Ad']), np.asarray(cfg['Bd'])
    NX = (K.Np + 1) * K.nx
    x = np.array(cfg['x0'], float); um1 = np.array(cfg['uminus1'], float)
    xs, us, warm = [], [], None
    for t in range(nsteps):
        K.update(x, um1, xref=(xref_fn(t) if xref_fn else None), solve=False)
        z, y, r = exact(K, warm)
        warm = (z, y)
        u0 = z[NX:NX + K.nu].copy()
        xs.append(x.copy()); us.append(u0)
        x = Ad @ x + Bd @ u0; um1 = u0
    np.savez_compressed(os.path.join(OUT, f"{name}_loop.npz"), x=np.array(xs), u=np.array(us))
    print(name, "loop u[:8]", np.array(us)[:8].ravel())


def random_batch(B=24, nsteps=4):
    cfg = pendulum()
    X0, Xref = pendulum_random(B, seed=0)
    Ad, Bd = cfg['Ad'], cfg['Bd']
    U = np.zeros((nsteps, B, 1)); X = np.zeros((nsteps, B, 4))
    for b in range(B):
        c = dict(cfg); c['x0'] = X0[b]; c['xref'] = Xref[b]
        K = MPC(**c); K.setup(solve=False)
        x = X0[b].copy(); um1 = np.zeros(1); warm = None
        for t in range(nsteps):
            K.update(x, um1, solve=False)
            z, y, r = exact(K, warm); warm = (z, y)
            u0 = z[84:85].copy(); U[t, b] = u0; X[t, b] = x
            x = Ad @ x + Bd @ u0; um1 = u0
    np.savez_compressed(os.path.join(OUT, "pend_rand.npz"), X0=X0, Xref=Xref, U=U, X=X)
    print("pend_rand", U[0, :4].ravel())


def variants():
    out = {}
    # (a) Nc < Np, time-varying xref, nonzero uref, Qu > 0  (mpc.py:618-692 demo shape)
    c = point_mass(); c['Np'] = 25; c['Nc'] = 10; c['uref'] = np.array([0.1])
    c['xmin'] = np.array([-10.0, -10.0]); c['xmax'] = np.array([7.0, 10.0])
    Xref = np.kron(np.ones((26, 1)), c['xref']); Xref[:, 0] = np.linspace(5.0, 7.0, 26)
    c['xref'] = Xref
    K = MPC(**c); K.setup(solve=False); z, y, r = exact(K)
    out['a_z'] = z; out['a_xref'] = Xref
    # (b) small MIMO with Nc < Np (channel-mixing delta-u quirk live)
    c = mimo(); c['Np'] = 12; c['Nc'] = 5; c['x0'] = np.array([0.3, -0.2, 0.1, 0.0, -0.4, 0.2, 0.0, 0.1])
    c['umin'] = -0.5 * np.ones(4); c['umax'] = 0.5 * np.ones(4); c['Qu'] = 0.1 * np.eye(4)
    K = MPC(**c); K.setup(solve=False); z, y, r = exact(K)
    out['b_z'] = z
    # (c) pendulum starting outside the soft position bound (slacks strongly active)
    c = pendulum(); c['x0'] = np.array([0.45, 0.3, -0.05, 0.1])
    K = MPC(**c); K.setup(solve=False); z, y, r = exact(K)
    out['c_z'] = z
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **out)
    print("variants ok")


if __name__ == "__main__":
    first_solve("pm", point_mass()); first_solve("pend", pendulum()); first_solve("mimo", mimo())
    closed_loop("pm", point_mass(), 30)
    closed_loop("pend", pendulum(), 40)
    closed_loop("mimo", mimo(), 12)
    random_batch()
    variants()
