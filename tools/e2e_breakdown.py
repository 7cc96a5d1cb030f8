#!/usr/bin/env python
"""Where the end-to-end step (pinned host buffers -> u* on the host) spends its time: wall clock of update() / solve part / output()
and of the raw C calls, 65 536 pendulum instances (development tool)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import pendulum_batch, make_controller

B = 65536
cfg, X0, Xref = pendulum_batch(B, "identical")
K = make_controller(cfg, X0, Xref, B, 0)
Xh = K.pinned_buffer("x0"); Uh = K.pinned_buffer("uminus1"); Xh[...] = X0; Uh[...] = 0
K.setup(); K.output()
L, h = K._L, K.handle
from pympc_b200._lib import ptr
acc = np.zeros(6); n = 0
for t in range(60):
    t0 = time.perf_counter(); K.update(Xh, Uh); t1 = time.perf_counter(); t2 = t1; U = K.output(); t3 = time.perf_counter()
    Uh[...] = U; Xh[...] = Xh @ cfg["Ad"].T + U @ cfg["Bd"].T
    if t >= 10:
        acc[:3] += (t1 - t0, t2 - t1, t3 - t2); n += 1
print("python API  : update+solve %.1f us  (-) %.1f us  output %.1f us  total %.1f us" % (*(1e6 * acc[:3] / n), 1e6 * acc[:3].sum() / n))
K.close()
