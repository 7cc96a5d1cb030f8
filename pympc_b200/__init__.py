"""pympc_b200 — batched linear MPC (pyMPC's MPCController) on B200: Python host over a C-ABI CUDA library."""
from .mpc import MPCController  # noqa: F401
from ._lib import BmpcError  # noqa: F401

__all__ = ["MPCController", "BmpcError"]
