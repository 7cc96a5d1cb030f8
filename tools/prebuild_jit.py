"""Build ahead of time the on-demand fast-path libraries the GPU tests ask for (pympc_b200/_jit/, git-ignored but shipped to the
GPU box with the snapshot), so that `pytest -m gpu` does not spend its first minutes in nvcc.  The tests build them themselves
when they are missing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pympc_b200 import build  # noqa: E402

if __name__ == "__main__":
    print(build.jit_shape(3, 1, 12, 9))                               # test_any_single_input_shape_gets_the_fast_path
    A = np.ones((3, 3)); A[2, 0] = 0.0                                  # test_multi_input_fast_path_built_on_demand_for_a_dense_system
    print(build.jit_multi_input_shape(3, 2, 10, 6, A, np.ones((3, 2))))
