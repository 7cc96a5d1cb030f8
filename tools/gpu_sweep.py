#!/usr/bin/env python
"""Kernel-variant / solver-option sweep on one GPU (development tool, not the bench).

    python tools/gpu_sweep.py [--steps K] [--warmup W] [--batch B] spec [spec ...]

spec = name[:lib=path][:workload=identical|random][:opt=value ...]  e.g.  base:first_iters=3  v2:lib=pympc_b200/libbmpc_v2.so
Each spec runs in its own process (the library is loaded once per process): device-resident closed loop of pendulum
instances exactly like bench.py's timed region (update -> solve -> output(commit), CUDA events on the launching stream, L2
flushed between steps) and prints one JSON line with ms/step, the ADMM / polish kernel times and solver statistics.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_one(spec, steps, warmup, B):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from bench import pendulum_batch
    from pympc_b200 import MPCController
    workload = spec["opts"].pop("workload", "identical")
    cfgp, X0, Xref = pendulum_batch(B, workload)
    keys = ("Qx", "QxN", "Qu", "QDu", "xmin", "xmax", "umin", "umax", "Dumin", "Dumax", "eps_feas")
    opts = {k: int(v) for k, v in spec["opts"].items()}
    K = MPCController(cfgp["Ad"], cfgp["Bd"], Np=20, x0=X0, xref=Xref, uminus1=np.zeros(1), batch=B, device=0,
                      **{k: cfgp[k] for k in keys}, **opts)
    K.setup(solve=False)
    K.solve(); K.output()
    cold = K.stats()
    dev = torch.device("cuda", 0)
    L, h = K._L, K.handle
    stream = torch.cuda.Stream(dev); torch.cuda.synchronize(dev); torch.cuda.set_stream(stream)   # handle 0 would mean the library's own stream
    L.bmpc_set_stream(h, stream.cuda_stream)
    Ad = torch.tensor(cfgp["Ad"], device=dev); Bd = torch.tensor(cfgp["Bd"], device=dev)
    Xd = torch.tensor(X0, device=dev); Xn = torch.empty_like(Xd)
    Uloc = torch.zeros(B, 1, dtype=torch.float64, device=dev)
    L.bmpc_bind_output(h, Uloc.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0; acc = dict(admm_iters=0, ms_admm=0.0, ms_polish=0.0, launches=0, polish_steps=0, unsolved=0); rounds = []; per = []
    for t in range(warmup + steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        L.bmpc_update(h, Xd.data_ptr(), Uloc.data_ptr(), None, 1, int(os.environ.get('SWEEP_ONDEV', '1')))
        assert L.bmpc_solve(h) == 0
        L.bmpc_output(h, None, None, 1, 1)
        e1.record(stream)
        torch.matmul(Xd, Ad.T, out=Xn); Xn.addmm_(Uloc, Bd.T)
        st = K.stats()
        torch.cuda.synchronize(dev)
        Xd, Xn = Xn, Xd
        if t >= warmup:
            ms = e0.elapsed_time(e1); tot += ms; per.append(ms)
            for k in acc:
                acc[k] += st[k]
            rounds.append(st["rounds"])
    K.close()
    per = np.array(per)
    return {"name": spec["name"], "workload": workload, "opts": opts, "B": B, "steps": steps, "ms_per_step": tot / steps,
            "ms_median": float(np.median(per)), "ms_max": float(per.max()), "solves_per_s": B * steps / (tot * 1e-3),
            "ms_admm": acc["ms_admm"] / steps, "ms_polish": acc["ms_polish"] / steps, "admm_iters_per_solve": acc["admm_iters"] / (B * steps),
            "polish_steps_per_solve": acc["polish_steps"] / (B * steps), "launches_per_step": acc["launches"] / steps,
            "mean_rounds": float(np.mean(rounds)), "unsolved": acc["unsolved"], "cold": {k: cold[k] for k in ("rounds", "ms_admm", "ms_polish", "unsolved")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--child", default=None)
    ap.add_argument("specs", nargs="*")
    a = ap.parse_args()
    if a.child:
        spec = json.loads(a.child)
        print("SWEEP " + json.dumps(run_one(spec, a.steps, a.warmup, a.batch)), flush=True)
        return
    for s in a.specs:
        parts = s.split(":"); spec = {"name": parts[0], "opts": {}}; env = dict(os.environ)
        for p in parts[1:]:
            k, v = p.split("=", 1)
            if k == "lib":
                env["BMPC_LIB"] = os.path.join(ROOT, v)
            else:
                spec["opts"][k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(a.steps), "--warmup", str(a.warmup), "--batch", str(a.batch),
                            "--child", json.dumps(spec)], env=env, capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("SWEEP ")]
        print(lines[-1][6:] if lines else json.dumps({"name": spec["name"], "error": (r.stderr or r.stdout)[-600:]}), flush=True)


if __name__ == "__main__":
    main()
