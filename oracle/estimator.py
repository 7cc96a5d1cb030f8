"""numpy restatement of ``LinearStateEstimator`` (/root/reference/pyMPC/kalman.py:109-134).  TEST INFRASTRUCTURE.
(The reference module itself does not import here: it needs the ``control`` package and uses ``scipy.size``,
removed from current scipy — SURVEY.md §2 row 4.)"""
import numpy as np


class LinearStateEstimator:
    def __init__(self, x0, A, B, C, D, L):
        self.x = np.copy(x0); self.y = C @ self.x                      # kalman.py:112-113
        self.A, self.B, self.C, self.D, self.L = A, B, C, D, L

    def predict(self, u):
        self.x = self.A @ self.x + self.B @ u                           # kalman.py:127
        self.y = self.C @ self.x                                        # kalman.py:128
        return self.x

    def update(self, y_meas):
        self.x = self.x + self.L @ (y_meas - self.y)                    # kalman.py:132
        return self.x
