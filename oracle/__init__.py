"""CPU oracle for the pyMPC hot path (TEST INFRASTRUCTURE — not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product (``pympc_b200``) never imports it and has no CPU fallback.

Contents
--------
qp_assembly   numpy restatement of the reference QP assembly / per-step update
              (``/root/reference/pyMPC/mpc.py:386-615``), pinned against the unmodified
              reference imported with a stub ``osqp`` module (tests/golden/make_golden.py).
osqp_port     C restatement of the OSQP ADMM algorithm (third-party dependency of the
              reference, ``setup.py:11`` unpinned, 0.6.x-era API; NOT present under
              /root/reference and not installable here) + ctypes wrapper.
kkt           solver-independent KKT certificate and an exact dense active-set solver used
              to generate the committed golden vectors.

Parity status: the reference's own tests pin no numbers at the solver boundary and the
``osqp`` package is unavailable, so the *solver* half of this oracle is "parity unpinned"
against real OSQP output; it is pinned instead by (i) the reference's unmodified assembly
code (exact), (ii) solver-independent KKT certificates < 1e-9 on the reference-assembled
(P, q, A, l, u), (iii) the analytic known answer for the point-mass example.
"""
