#!/usr/bin/env python
"""cProfile of the end-to-end step through the public API (development tool): where update() / output() spend host time."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import pendulum_batch, make_controller

B = 65536
cfg, X0, Xref = pendulum_batch(B, "identical")
K = make_controller(cfg, X0, Xref, B, 0)
Xh = K.pinned_buffer("x0"); Uh = K.pinned_buffer("uminus1"); Xh[...] = X0; Uh[...] = 0
K.setup(); K.output()
for t in range(20):
    K.update(Xh, Uh); U = K.output()
pr = cProfile.Profile()
N = 300
t0 = time.perf_counter()
pr.enable()
for t in range(N):
    K.update(Xh, Uh); U = K.output()
pr.disable()
print("per step (profiled) %.1f us" % (1e6 * (time.perf_counter() - t0) / N))
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(14)
K.close()
