"""In-tree build of libbmpc.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "bmpc.cu")
SHAPES = os.path.join(_HERE, "csrc", "tpi_shapes.inc")
import glob
# every header / table the translation unit includes: an edit to any of them makes the library stale
DEPS = sorted(set([SRC, SHAPES, os.path.join(_HERE, "..", "include", "bmpc.h")] + glob.glob(os.path.join(_HERE, "csrc", "*.cuh")) +
                  glob.glob(os.path.join(_HERE, "csrc", "*.inc"))))
LIB = os.path.join(_HERE, "libbmpc.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(p) for p in DEPS)
    if stale:
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", LIB]
        subprocess.check_call(cmd)
    return LIB


def add_shape(nx, nu, Np, Nc=None):
    """Register one more compile-time fast-path shape (nu must be 1, Nc <= Np, default Nc = Np) and rebuild."""
    Nc = Np if Nc is None else Nc
    if nu != 1 or Np * nx > 128 or Np >= 32 or not (1 <= Nc <= Np):
        raise ValueError("fast-path shapes need nu == 1, Np*nx <= 128, Np < 32, 1 <= Nc <= Np")
    line = f"BMPC_TPI_SHAPE({nx}, {nu}, {Np}, {Nc})"
    txt = open(SHAPES).read()
    if line not in txt:
        open(SHAPES, "a").write(line + "\n")
    return build(force=True)


def jit_shape(nx, nu, Np, Nc=None, verbose=False):
    """Build (once; cached under pympc_b200/_jit/) a copy of the library whose fast-path table holds exactly the shape
    (nx, 1, Np, Nc) and return its path: lets a controller of ANY single-input shape run on the thread-per-instance kernels
    without editing csrc/tpi_shapes.inc.  Raises on shapes the fast path cannot hold or when nvcc is not available."""
    Nc = Np if Nc is None else Nc
    if nu != 1 or Np * nx > 128 or Np >= 32 or not (1 <= Nc <= Np):
        raise ValueError("fast-path shapes need nu == 1, Np*nx <= 128, Np < 32, 1 <= Nc <= Np")
    jdir = os.path.join(_HERE, "_jit"); os.makedirs(jdir, exist_ok=True)
    tag = f"{nx}_{nu}_{Np}_{Nc}"
    inc = os.path.join(jdir, f"shape_{tag}.inc"); lib = os.path.join(jdir, f"libbmpc_{tag}.so")
    if not os.path.exists(inc):
        open(inc, "w").write(f"BMPC_TPI_SHAPE({nx}, {nu}, {Np}, {Nc})\n")
    deps = [p for p in DEPS if p != SHAPES]
    if not os.path.exists(lib) or any(os.path.getmtime(lib) < os.path.getmtime(p) for p in deps):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f'-DBMPC_TPI_SHAPES_FILE="{inc}"', SRC, "-o", lib + ".tmp"]
        subprocess.check_call(cmd)
        os.replace(lib + ".tmp", lib)
    return lib


def jit_multi_input_shape(nx, nu, Np, Nc=None, Ad=None, Bd=None, verbose=False):
    """Build (once; cached under pympc_b200/_jit/) a copy of the library whose multi-input fast-path table holds the shape
    (nx, nu, Np, Nc) — with the sparsity pattern of (Ad, Bd) fixed at compile time when they are given and small enough for the
    pattern masks (the sweeps then skip the structural zeros) — and return its path."""
    import numpy as np
    Nc = Np if Nc is None else Nc
    if nu < 1 or 2 * nx + 10 * nu > 64 or not (1 <= Nc <= Np):
        raise ValueError("multi-input fast-path shapes need 2 nx + 10 nu <= 64 and 1 <= Nc <= Np")
    line = f"BMPC_TPM_SHAPE({nx}, {nu}, {Np}, {Nc})"; tag = f"m{nx}_{nu}_{Np}_{Nc}"
    if Ad is not None and Bd is not None and nx * nx <= 64 and nx * nu <= 32:
        A = np.asarray(Ad, float).reshape(nx, nx); Bm = np.asarray(Bd, float).reshape(nx, nu)
        am = sum(1 << i for i, v in enumerate(A.ravel()) if v != 0.0); bm = sum(1 << i for i, v in enumerate(Bm.ravel()) if v != 0.0)
        if am != (1 << (nx * nx)) - 1 or bm != (1 << (nx * nu)) - 1:
            line = f"BMPC_TPM_SPARSE_SHAPE({nx}, {nu}, {Np}, {Nc}, 0x{am:x}ull, 0x{bm:x}u)\n" + line      # the dense entry stays behind it
            tag += f"_{am:x}_{bm:x}"
    jdir = os.path.join(_HERE, "_jit"); os.makedirs(jdir, exist_ok=True)
    inc = os.path.join(jdir, f"shape_{tag}.inc"); lib = os.path.join(jdir, f"libbmpc_{tag}.so")
    if not os.path.exists(inc):
        open(inc, "w").write(line + "\n")
    deps = [p for p in DEPS if not p.endswith("tpm_shapes.inc")]
    if not os.path.exists(lib) or any(os.path.getmtime(lib) < os.path.getmtime(p) for p in deps):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f'-DBMPC_TPM_SHAPES_FILE="{inc}"', SRC, "-o", lib + ".tmp"]
        subprocess.check_call(cmd)
        os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 3 and sys.argv[1] == "--add-shape":          # nx,nu,Np[,Nc]
        print(add_shape(*[int(v) for v in sys.argv[2].split(",")]))
    else:
        print(build(force=True, verbose=True))
