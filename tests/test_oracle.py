"""CPU tests of the oracle: assembly vs the reference / golden fixtures, exact solver and the C
restatement of OSQP vs the KKT-certified goldens."""
import numpy as np
import pytest

from conftest import golden
from oracle.qp_assembly import QPData
from oracle.kkt import solve_exact, kkt_residuals
from pympc_b200.workloads import point_mass, pendulum, mimo, WORKLOADS
from refharness import reference_available, load_reference_controller

CASES = {"pm": point_mass, "pend": pendulum, "mimo": mimo}


@pytest.mark.parametrize("name", list(CASES))
def test_assembly_matches_golden_vectors(name):
    """q, l, u and fingerprints of P, A stored from the reference's own assembly."""
    g = golden(f"{name}_first.npz")
    Q = QPData(**CASES[name]())
    assert np.array_equal(Q.q, g["q"])
    for a, b in ((Q.l, g["l"]), (Q.u, g["u"])):
        assert np.array_equal(np.isinf(a), np.isinf(b))
        assert np.array_equal(np.where(np.isinf(a), 0, a), np.where(np.isinf(b), 0, b))
    assert np.array_equal(Q.P[np.nonzero(Q.P)], g["P_data"])
    assert np.allclose([Q.A.sum(), np.abs(Q.A).sum()], g["A_sum"], rtol=0, atol=1e-12)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("variant", ["pm", "pend", "mimo", "pm_Nc_tv", "mimo_Nc_uref", "pend_inf"])
def test_assembly_matches_reference_live(variant):
    MPC = load_reference_controller()
    import scipy.sparse as sp
    if variant in CASES:
        cfg = CASES[variant]()
    elif variant == "pm_Nc_tv":
        cfg = point_mass(); cfg["Np"] = 25; cfg["Nc"] = 10; cfg["xref"] = np.kron(np.ones((26, 1)), cfg["xref"])
    elif variant == "mimo_Nc_uref":
        cfg = mimo(); cfg["Np"] = 12; cfg["Nc"] = 5; cfg["uref"] = np.array([0.1, -0.2, 0.3, 0.0]); cfg["Qu"] = np.diag([1., 2, 3, 4])
    else:
        cfg = pendulum(); cfg["Qx"] = sp.diags([0.3, 0, 1.0, 0]); cfg["xmax"] = np.array([np.inf, 1, 2, 3]); cfg["Dumin"] = np.array([-np.inf])
    K = MPC(**cfg); K.setup(solve=False)
    Q = QPData(**cfg)
    fin = lambda v: np.where(np.isinf(v), 0, v)
    assert np.array_equal(K.P.toarray(), Q.P) and np.array_equal(K.A.toarray(), Q.A)
    assert np.array_equal(K.q, Q.q) and np.array_equal(fin(K.l), fin(Q.l)) and np.array_equal(fin(K.u), fin(Q.u))
    rng = np.random.default_rng(3)
    for t in range(3):
        x = rng.normal(size=Q.nx); um1 = rng.normal(size=Q.nu)
        xr = cfg["xref"] if t == 0 else rng.normal(size=np.shape(cfg["xref"]))
        K.update(x, um1, xref=xr, solve=False); Q.update(x, um1, xr)
        assert np.array_equal(K.q, Q.q) and np.array_equal(fin(K.l), fin(Q.l)) and np.array_equal(fin(K.u), fin(Q.u))
        assert abs(K.J_CNST - Q.constant_term()) < 1e-12


@pytest.mark.parametrize("name", ["pm", "pend"])
def test_exact_solver_reproduces_golden(name):
    g = golden(f"{name}_first.npz")
    Q = QPData(**CASES[name]())
    z, y, r = solve_exact(Q.P, Q.q, Q.A, Q.l, Q.u)
    assert max(r.values()) < 1e-9
    assert np.max(np.abs(z[Q.NX:Q.NX + Q.NU] - g["u_seq"])) < 1e-8
    rg = kkt_residuals(Q.P, Q.q, Q.A, Q.l, Q.u, g["z"], g["y"])
    assert max(rg.values()) < 1e-9          # the stored golden is itself a KKT point of the oracle-assembled QP


def test_point_mass_analytic_known_answer():
    """rate- then magnitude-limited ramp 0.2, 0.4, ..., 1.2, 1.2."""
    g = golden("pm_loop.npz")
    expect = np.minimum(0.2 * (np.arange(8) + 1), 1.2)
    assert np.max(np.abs(g["u"][:8, 0] - expect)) < 1e-9


@pytest.mark.parametrize("name", ["pm", "pend", "mimo"])
def test_osqp_port_default_and_tight(name, osqp_port_lib):
    """The C restatement: at the reference's eps=1e-3 it stops within a few 1e-2 of the optimum (as OSQP does);
    at eps=1e-9 it agrees with the certified optimum to 1e-6."""
    g = golden(f"{name}_first.npz")
    Q = QPData(**CASES[name]()); Pu, Ac = Q.to_csc()
    s = osqp_port_lib.OSQP().setup(Pu, Q.q, Ac, Q.l, Q.u)
    r = s.solve()
    assert r.info.status == "solved" and r.info.iter % 25 == 0
    assert np.max(np.abs(r.x[Q.u0_slice()] - g["u_seq"][:Q.nu])) < 5e-2
    s2 = osqp_port_lib.OSQP().setup(Pu, Q.q, Ac, Q.l, Q.u, eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)
    r2 = s2.solve()
    assert r2.info.status == "solved"
    assert np.max(np.abs(r2.x[Q.u0_slice()] - g["u_seq"][:Q.nu])) < 1e-6
    assert abs(r2.info.obj_val - float(g["obj"])) < 1e-6 * (1 + abs(float(g["obj"])))


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")
def test_unmodified_reference_runs_on_the_port(osqp_port_lib):
    """Reference MPCController end to end with osqp_port injected as `osqp` (tight tolerance) vs golden loop."""
    import sys
    saved = sys.modules.get("osqp")
    sys.modules["osqp"] = osqp_port_lib
    try:
        for m in [k for k in sys.modules if k.startswith("pyMPC")]:
            del sys.modules[m]
        sys.path.insert(0, "/root/reference")
        from pyMPC.mpc import MPCController
        cfg = pendulum()
        K = MPCController(**cfg, eps_abs=1e-9, eps_rel=1e-9)
        K.setup(solve=False)
        # the reference does not forward max_iter; tight tolerance needs more than OSQP's 4000 default
        K.prob = osqp_port_lib.OSQP()
        K.prob.setup(K.P, K.q, K.A, K.l, K.u, warm_start=True, eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)
        K.solve()
        g = golden("pend_loop.npz")
        x = np.array(cfg["x0"]); um1 = np.array(cfg["uminus1"])
        for t in range(5):
            K.update(x, um1); u = np.array(K.output())
            assert np.max(np.abs(u - g["u"][t])) < 1e-6
            x = cfg["Ad"] @ x + cfg["Bd"] @ u; um1 = u
    finally:
        for m in [k for k in sys.modules if k.startswith("pyMPC")]:
            del sys.modules[m]
        if saved is not None:
            sys.modules["osqp"] = saved
        else:
            sys.modules.pop("osqp", None)


def test_batch_cpu_driver_matches_single(osqp_port_lib):
    Q = QPData(**pendulum())
    B = 6
    bc = osqp_port_lib.BatchCPU(Q, B, eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)
    g = golden("pend_loop.npz")
    X0 = np.tile(g["x"][0], (B, 1)); Um1 = np.zeros((B, 1)); Xref = np.tile(pendulum()["xref"], (B, 1))
    U, st, it = bc.step(X0, Um1, Xref, nthreads=2)
    assert np.all(st == 1) and np.max(np.abs(U - g["u"][0])) < 1e-6
    bc.close()


@pytest.mark.parametrize("name", ["pm", "pend", "mimo"])
def test_goldens_vs_independent_ldp_solver(name):
    """the goldens (oracle/kkt.py: ADMM -> active set on the reference-form QP) against an algorithmically independent exact
    solver: condensed QP with explicit slack as a least-distance problem solved by ONE Lawson-Hanson NNLS (oracle/ldp.py)"""
    from oracle.ldp import solve_mpc
    cfg = {"pm": point_mass, "pend": pendulum, "mimo": mimo}[name](); g = golden(f"{name}_first.npz")
    assert np.max(np.abs(solve_mpc(QPData(**cfg)) - g["u_seq"])) < 1e-9
    if name == "mimo":
        return
    gl = golden(f"{name}_loop.npz"); x = np.array(cfg["x0"], float); u = np.array(cfg["uminus1"], float)
    for t in range(10):
        c = dict(cfg); c["x0"] = x; c["uminus1"] = u
        assert abs(solve_mpc(QPData(**c))[0] - gl["u"][t][0]) < 1e-9, t
        u = gl["u"][t]; x = cfg["Ad"] @ x + cfg["Bd"] @ u
