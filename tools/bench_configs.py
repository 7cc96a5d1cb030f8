"""Side measurements for the BASELINE configs that are not the bench line (configs 3 and 4): closed-loop steps through
the public API (pinned host buffers), reporting solves/s and the device time of the ADMM / polish kernels.
    python tools/bench_configs.py [pend_random|mimo|pend_identical] [--batch B] [--steps K] [--opt k=v ...]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pympc_b200 import MPCController
from pympc_b200.workloads import pendulum, pendulum_random, mimo

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="mimo")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
opts = {}
for kv in a.opt:
    k, v = kv.split("="); opts[k] = float(v) if "." in v else int(v)
if a.workload == "mimo":
    cfg = mimo(); B = a.batch or 16384
    rng = np.random.default_rng(4); X0 = 0.3 * rng.standard_normal((B, 8)); Xref = np.tile(cfg["xref"], (B, 1))
elif a.workload == "pend_per_instance":
    # SURVEY 8f-3 at BASELINE size: one (Ad, Bd, Qx, umax) per instance, pendulum shape
    cfg = pendulum(); B = a.batch or 65536; rng = np.random.default_rng(13)
    X0, Xref = pendulum_random(B, 3)
    cfg = dict(cfg, Ad=cfg["Ad"][None] + 0.01 * rng.standard_normal((B, 4, 4)) * (cfg["Ad"] != 0), Bd=cfg["Bd"][None] * (1 + 0.1 * rng.standard_normal((B, 1, 1))),
               Qx=np.diag([0.3, 0, 1.0, 0])[None] * (1 + 0.3 * rng.random((B, 1, 1))), umax=15.0 + 10 * rng.random((B, 1)))
    cfg["QxN"] = cfg["Qx"]; cfg["umin"] = -cfg["umax"]
else:
    cfg = pendulum(); B = a.batch or 65536
    if a.workload == "pend_random":
        X0, Xref = pendulum_random(B, 0)
    else:
        X0 = np.tile(cfg["x0"], (B, 1)); Xref = np.tile(cfg["xref"], (B, 1))
nu = cfg["Bd"].shape[-1]
keys = [k for k in ("Qx", "QxN", "Qu", "QDu", "xmin", "xmax", "umin", "umax", "Dumin", "Dumax", "eps_feas") if k in cfg]
K = MPCController(cfg["Ad"], cfg["Bd"], Np=cfg["Np"], x0=X0, xref=Xref, uminus1=np.zeros(nu), batch=B, **{k: cfg[k] for k in keys}, **opts)
t0 = time.perf_counter(); K.setup(); cold = time.perf_counter() - t0
st0 = K.stats()
Xh = K.pinned_buffer("x0"); Uh = K.pinned_buffer("uminus1"); Xh[...] = X0; Uh[...] = K.output()
plant = (lambda X, U: np.einsum("bij,bj->bi", cfg["Ad"], X) + np.einsum("bij,bj->bi", cfg["Bd"], U)) if np.ndim(cfg["Ad"]) == 3 else (lambda X, U: X @ cfg["Ad"].T + U @ cfg["Bd"].T)
Xh[...] = plant(Xh, Uh)
tt = 0.0; acc = {"ms_admm": 0.0, "ms_polish": 0.0, "rounds": 0, "admm_iters": 0, "unsolved": 0, "polish_steps": 0}
for t in range(a.warmup + a.steps):
    t0 = time.perf_counter(); K.update(Xh, Uh); U = K.output(); dt = time.perf_counter() - t0
    s = K.stats()
    if t >= a.warmup:
        tt += dt
        for k in acc: acc[k] += s[k]
    Uh[...] = U; Xh[...] = plant(Xh, U)
print(json.dumps({"workload": a.workload, "batch": B, "opts": opts, "solves_per_s_e2e": B * a.steps / tt, "ms_per_step": 1e3 * tt / a.steps,
                  "ms_admm_per_step": acc["ms_admm"] / a.steps, "ms_polish_per_step": acc["ms_polish"] / a.steps,
                  "mean_rounds": acc["rounds"] / a.steps, "admm_iters_per_solve": acc["admm_iters"] / (B * a.steps),
                  "polish_steps_per_solve": acc["polish_steps"] / (B * a.steps), "unsolved": acc["unsolved"],
                  "cold_setup_s": cold, "cold_rounds": st0["rounds"], "cold_unsolved": st0["unsolved"]}))
