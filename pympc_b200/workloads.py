"""Workload definitions (BASELINE.json ``configs``; parameters from SURVEY.md §8d).

Each function returns the keyword arguments of ``MPCController(...)`` for one MPC instance.
Sources of the numbers (reference file:line):
  point_mass   /root/reference/examples/example_point_mass.py:11-71
  pendulum     /root/reference/examples/example_inverted_pendulum.py:10-69
  mimo         shape of /root/reference/test_scripts/cvx_mpc_reference_governor_du_mimo.py:15-62,
               SISO blocks in canonical form per test_scripts/reference_governor/system_dynamics.py:9-24
"""
import numpy as np


def point_mass():
    Ts, M, b = 0.2, 2.0, 0.3
    Ac = np.array([[0.0, 1.0], [0.0, -b / M]])
    Bc = np.array([[0.0], [1.0 / M]])
    return dict(Ad=np.eye(2) + Ac * Ts, Bd=Bc * Ts, Np=20,
                x0=np.array([0.1, 0.2]), xref=np.array([7.0, 0.0]), uminus1=np.array([0.0]),
                Qx=np.diag([0.5, 0.1]), QxN=np.diag([0.5, 0.1]), Qu=2.0 * np.eye(1), QDu=10.0 * np.eye(1),
                xmin=np.array([-100.0, -100.0]), xmax=np.array([100.0, 100.0]),
                umin=np.array([-1.2]), umax=np.array([1.2]),
                Dumin=np.array([-0.2]), Dumax=np.array([0.2]))


def pendulum():
    M, m, b, ftheta, l, g, Ts = 0.5, 0.2, 0.1, 0.1, 0.3, 9.81, 50e-3
    Ac = np.array([[0, 1, 0, 0],
                   [0, -b / M, -(g * m) / M, (ftheta * m) / M],
                   [0, 0, 0, 1],
                   [0, b / (M * l), (M * g + g * m) / (M * l), -(M * ftheta + ftheta * m) / (M * l)]])
    Bc = np.array([[0.0], [1.0 / M], [0.0], [-1 / (M * l)]])
    return dict(Ad=np.eye(4) + Ac * Ts, Bd=Bc * Ts, Np=20,
                x0=np.array([0.0, 0.0, 15 * 2 * np.pi / 360, 0.0]),
                xref=np.array([0.3, 0.0, 0.0, 0.0]), uminus1=np.array([0.0]),
                Qx=np.diag([0.3, 0.0, 1.0, 0.0]), QxN=np.diag([0.3, 0.0, 1.0, 0.0]),
                Qu=0.0 * np.eye(1), QDu=0.01 * np.eye(1),
                xmin=np.array([-1.0, -100.0, -100.0, -100.0]), xmax=np.array([0.3, 100.0, 100.0, 100.0]),
                umin=np.array([-20.0]), umax=np.array([20.0]),
                Dumin=np.array([-5.0]), Dumax=np.array([5.0]), eps_feas=1e3)


def pendulum_random(batch, seed=0):
    """Per-instance x0 / xref of BASELINE config 3: returns (X0[B,4], Xref[B,4])."""
    rng = np.random.default_rng(seed)
    X0 = np.stack([rng.uniform(-0.5, 0.25, batch), rng.uniform(-0.5, 0.5, batch),
                   rng.uniform(-15, 15, batch) * np.pi / 180, rng.uniform(-0.5, 0.5, batch)], axis=1)
    Xref = np.zeros((batch, 4))
    Xref[:, 0] = rng.uniform(-0.5, 0.3, batch)
    return X0, Xref


def mimo():
    r, ws = 0.9, [0.2, 0.4, 0.2, 0.4]
    nb = len(ws)
    Ad = np.zeros((2 * nb, 2 * nb)); Bd = np.zeros((2 * nb, nb)); C = np.zeros((nb, 2 * nb))
    xref = np.zeros(2 * nb)
    for i, w in enumerate(ws):
        kap = 1 - 2 * r * np.cos(w) + r * r
        Ad[2 * i:2 * i + 2, 2 * i:2 * i + 2] = [[2 * r * np.cos(w), -r * r], [1.0, 0.0]]
        Bd[2 * i, i] = 1.0
        C[i, 2 * i + 1] = kap
        xref[2 * i:2 * i + 2] = 1.0 / kap                      # steady state with C xref = 1
    Qx = C.T @ (20.0 * np.eye(nb)) @ C
    return dict(Ad=Ad, Bd=Bd, Np=40, x0=np.zeros(2 * nb), xref=xref, uminus1=np.zeros(nb),
                Qx=Qx, QxN=Qx, Qu=np.zeros((nb, nb)), QDu=0.5 * np.eye(nb),
                xmin=-100.0 * np.ones(2 * nb), xmax=100.0 * np.ones(2 * nb),
                umin=-1000.0 * np.ones(nb), umax=1000.0 * np.ones(nb),
                Dumin=-0.2 * np.ones(nb), Dumax=0.2 * np.ones(nb))


WORKLOADS = {"point_mass": point_mass, "pendulum": pendulum, "mimo": mimo}
