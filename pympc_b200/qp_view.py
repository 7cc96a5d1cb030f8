"""Reference-form view of the QP (API completeness, NOT on the hot path).

The reference exposes ``K.P, K.q, K.A, K.l, K.u`` after ``setup()`` / ``update()`` (mpc.py:597-606, 404-452).  The GPU
path never materialises them (it condenses on the device), so ``MPCController`` builds them lazily on the host, for
instance 0 of a batch, with scipy.sparse in the reference's variable / row order:
z = [x_0..x_Np | u_0..u_{Nc-1} | eps_0..eps_Np],  rows = [dynamics | x(+eps) bounds | u bounds | delta-u rows].
"""
import numpy as np
import scipy.sparse as sp


def assemble(Ad, Bd, Np, Nc, Qx, QxN, Qu, QDu, xmin, xmax, umin, umax, Dumin, Dumax, eps_feas, uref, soft_on=True):
    nx, nu = Bd.shape
    NX, NU = (Np + 1) * nx, Nc * nu
    w = np.ones(Nc); w[-1] = Np - Nc + 1
    T = 2 * np.eye(Nc) - np.eye(Nc, k=1) - np.eye(Nc, k=-1); T[-1, -1] = 1
    blocks = [sp.block_diag([sp.kron(sp.eye(Np), Qx), QxN]), sp.kron(sp.diags(w), Qu) + sp.kron(T, QDu)]
    if soft_on:
        blocks.append(eps_feas * sp.eye(NX))
    P = sp.block_diag(blocks, format="csc")
    n = P.shape[0]
    Ax = sp.kron(sp.eye(Np + 1), -sp.eye(nx)) + sp.kron(sp.eye(Np + 1, k=-1), Ad)
    hold = np.zeros((Np + 1, Nc))
    for k in range(1, Np + 1):
        hold[k, min(k - 1, Nc - 1)] = 1.0
    rows = [sp.hstack([Ax, sp.kron(hold, Bd)] + ([sp.csc_matrix((NX, NX))] if soft_on else [])),
            sp.hstack([sp.eye(NX), sp.csc_matrix((NX, NU))] + ([sp.eye(NX)] if soft_on else [])),
            sp.hstack([sp.csc_matrix((NU, NX)), sp.eye(NU)] + ([sp.csc_matrix((NU, NX))] if soft_on else []))]
    Dblk = sp.vstack([sp.hstack([sp.eye(nu), sp.csc_matrix((nu, NU - nu))]), -sp.eye(NU) + sp.eye(NU, k=1)])
    rows.append(sp.hstack([sp.csc_matrix(((Nc + 1) * nu, NX)), Dblk] + ([sp.csc_matrix(((Nc + 1) * nu, NX))] if soft_on else [])))
    A = sp.vstack(rows).tocsc()
    return P, A, n, w


def vectors(Np, Nc, nx, nu, Qx, QxN, Qu, QDu, w, xmin, xmax, umin, umax, Dumin, Dumax, uref, x0, um1, xref, soft_on=True):
    NX = (Np + 1) * nx
    xref = np.asarray(xref, float)
    if xref.ndim == 2:
        qX = -np.concatenate([Qx @ xref[k] for k in range(Np)] + [QxN @ xref[Np]])
    else:
        qX = -np.concatenate([np.tile(Qx @ xref, Np), QxN @ xref])
    qU = -np.kron(w, Qu @ uref); qU[:nu] -= QDu @ um1
    q = np.concatenate([qX, qU] + ([np.zeros(NX)] if soft_on else []))
    eq = np.concatenate([-x0, np.zeros(Np * nx)])
    ldu = np.tile(Dumin, Nc + 1); udu = np.tile(Dumax, Nc + 1); ldu[:nu] += um1; udu[:nu] += um1
    l = np.concatenate([eq, np.tile(xmin, Np + 1), np.tile(umin, Nc), ldu])
    u = np.concatenate([eq, np.tile(xmax, Np + 1), np.tile(umax, Nc), udu])
    return q, l, u
