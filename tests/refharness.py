"""Loads the UNMODIFIED reference (``/root/reference/pyMPC/mpc.py``) with a stub ``osqp`` module.

Only usable where /root/reference exists (this container, not the GPU box): used by
``tests/golden/make_golden.py`` to generate committed fixtures and by tests that skip otherwise.
The stub records what pyMPC hands to the solver."""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pyMPC", "mpc.py"))


def load_reference_controller():
    if "osqp" not in sys.modules:
        stub = types.ModuleType("osqp")

        class _OSQP:
            def setup(self, *a, **k):
                self.setup_args = (a, k)

            def update(self, **k):
                self.update_args = k

            def solve(self):
                raise RuntimeError("osqp is not installed; the reference cannot solve here")

        stub.OSQP = _OSQP
        stub.__stub__ = True
        sys.modules["osqp"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from pyMPC.mpc import MPCController
    return MPCController
