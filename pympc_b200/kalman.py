"""Batched ``LinearStateEstimator`` on the GPU — the step on the input side of the MPC path.

Mirrors ``/root/reference/pyMPC/kalman.py:109-152`` (``__init__``, ``out_y``, ``predict``, ``update``, ``sim``) for a
batch of B estimators that share (A, B, C, D, L): ``predict(u)`` does ``x <- A x + B u; y <- C x`` (kalman.py:126-129)
and ``update(y_meas)`` does ``x <- x + L (y_meas - y)`` (kalman.py:131-133, ``y`` is not refreshed — as in the
reference).  The state stays on the device; ``device_state()`` feeds ``MPCController.update_from_device``.
The Kalman *design* helpers (``kalman_design*``: DARE through the ``control`` package, kalman.py:24-106) are out of
scope: they run once, on the host, and produce the gain ``L`` this class takes.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import BmpcError, ptr


class LinearStateEstimator:
    def __init__(self, x0, A, B, C, D, L, batch=None, device=0):
        self._L = _lib.load()
        A = np.ascontiguousarray(A, float); B = np.ascontiguousarray(B, float).reshape(A.shape[0], -1)
        C = np.ascontiguousarray(C, float).reshape(-1, A.shape[0]); Lg = np.ascontiguousarray(L, float).reshape(A.shape[0], -1)
        self.A, self.B, self.C, self.D, self.L = A, B, C, np.asarray(D, float), Lg
        self.nx, self.nu, self.ny = A.shape[0], B.shape[1], C.shape[0]
        self.batch = batch
        self._B = 1 if batch is None else int(batch)
        x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, float).reshape(-1, self.nx), (self._B, self.nx)))
        h = ctypes.c_void_p()
        rc = self._L.bmpc_est_create(self.nx, self.nu, self.ny, self._B, int(device), ptr(A), ptr(B), ptr(C), ptr(Lg), ptr(x0),
                                     ctypes.byref(h))
        if rc < 0:
            raise BmpcError(f"bmpc_est_create failed ({rc}): {self._L.bmpc_last_error(None).decode()}")
        self._h = h

    def _shape(self, a, n):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, float).reshape(-1, n), (self._B, n)))
        return a

    def _out(self, a):
        return a if self.batch is not None else a[0]

    @property
    def x(self):
        x = np.empty((self._B, self.nx)); self._L.bmpc_est_get(self._h, ptr(x), None); return self._out(x)

    @property
    def y(self):
        y = np.empty((self._B, self.ny)); self._L.bmpc_est_get(self._h, None, ptr(y)); return self._out(y)

    def out_y(self, u):
        return self.y

    def predict(self, u):
        rc = self._L.bmpc_est_predict(self._h, ptr(self._shape(u, self.nu)), 0)
        if rc < 0:
            raise BmpcError("bmpc_est_predict failed")
        return self.x

    def update(self, y_meas):
        rc = self._L.bmpc_est_update(self._h, ptr(self._shape(y_meas, self.ny)), 0)
        if rc < 0:
            raise BmpcError("bmpc_est_update failed")
        return self.x

    def predict_device(self, u_ptr):
        """u already on the device (e.g. the MPC output buffer): no host traffic."""
        return self._L.bmpc_est_predict(self._h, u_ptr, 1)

    def update_device(self, y_ptr):
        return self._L.bmpc_est_update(self._h, y_ptr, 1)

    def attach(self, controller):
        """Chain this estimator with an MPCController on the device (estimate -> update_from_device -> u0 -> predict_device)
        without host reads in between: both run on the controller's stream, and the controller's deferred solve is retired before
        its output buffer is consumed.  `None` detaches."""
        rc = self._L.bmpc_est_attach(self._h, controller.handle if controller is not None else None)
        if rc < 0:
            raise BmpcError("bmpc_est_attach failed (controller on another device?)")

    def device_state(self):
        """Device pointer (int) of x [B, nx], usable as bmpc_update(x0=..., on_device=1)."""
        return self._L.bmpc_est_state_ptr(self._h)

    def sim(self, u_seq, x=None):
        """Open-loop output prediction on the host, as the reference does (kalman.py:136-152); B = 1 only."""
        x = np.array(self.x if x is None else x, float).reshape(-1)
        u_seq = np.asarray(u_seq, float).reshape(len(u_seq), -1)
        assert u_seq.shape[1] == self.nu
        y = np.zeros((u_seq.shape[0], self.ny))
        for i in range(u_seq.shape[0]):
            y[i] = self.C @ x + self.D.reshape(self.ny, -1) @ u_seq[i]
            x = self.A @ x + self.B @ u_seq[i]
        return y

    def close(self):
        if getattr(self, "_h", None):
            self._L.bmpc_est_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
