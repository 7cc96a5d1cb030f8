"""CPU baseline of SURVEY 8(d), all variants, on this box's host cores (the oracle's C restatement of OSQP; the real
`osqp` wheel is not available offline):
  (i) single core / (ii) all cores, solver only, eps = 1e-3 (the reference default) and a tight eps = 1e-6,
  (iii) one solver object per instance driven from Python like `MPCController.update()` does it (numpy re-assembly of
        q, l, u per step + ctypes calls), single core — what a user of the reference pays per controller.
    python tools/cpu_baseline_report.py [--sample 4096] [--steps 3]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle.qp_assembly import QPData
from oracle import osqp_port

ap = argparse.ArgumentParser()
ap.add_argument("--sample", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
out = []
for workload in ("identical", "random"):
    for eps in (1e-3, 1e-6):
        for threads in (1, None):
            b = a.sample if threads is None else max(64, a.sample // 16)
            r = bench.cpu_arm(a.steps, 2, b, workload, threads=threads, eps_abs=eps, eps_rel=eps)
            out.append({"variant": "solver-only", "workload": workload, "eps": eps, "threads": r["cores"], "solves_per_s": r["value"], "sample": r["sample"]})
            print(json.dumps(out[-1]), flush=True)

# (iii) Python-driven, one object per instance
cfg, X0, Xref = bench.pendulum_batch(64, "random")
objs = []
for i in range(64):
    Q = QPData(**dict(cfg, x0=X0[i], xref=Xref[i])); Pu, Ac = Q.to_csc()
    m = osqp_port.OSQP(); m.setup(P=Pu, q=Q.q, A=Ac, l=Q.l, u=Q.u, verbose=False, eps_abs=1e-3, eps_rel=1e-3)
    m.solve(); objs.append((Q, m))
X = X0.copy(); U = np.zeros((64, 1)); tt = 0.0; n = 0
for t in range(2 + a.steps):
    t0 = time.perf_counter()
    for i, (Q, m) in enumerate(objs):
        Q.update(X[i], U[i], Xref[i])                          # numpy restatement of _update_QP_matrices_ (mpc.py:386-454)
        m.update(q=Q.q, l=Q.l, u=Q.u)
        r = m.solve()
        U[i] = r.x[Q.u0_slice()]
    dt = time.perf_counter() - t0
    if t >= 2:
        tt += dt; n += 64
    X = X @ cfg["Ad"].T + U @ cfg["Bd"].T
out.append({"variant": "python-driven (update + solve per controller object)", "workload": "random", "eps": 1e-3, "threads": 1, "solves_per_s": n / tt,
            "sample": f"64 controller objects x {a.steps} closed-loop steps"})
print(json.dumps(out[-1]), flush=True)
