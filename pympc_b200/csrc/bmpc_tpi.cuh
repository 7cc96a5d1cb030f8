// bmpc_tpi.cuh — thread-per-instance (TPI) fast path for SMALL MPC problems with compile-time shape
// (nx, nu, Np, Nc), e.g. the inverted pendulum (4,1,20,20) and the point mass (2,1,20,20).
//
// Why: the first (team-per-instance) kernels spent ~2700 warp-instructions per ADMM iteration on index
// arithmetic and dense mat-vecs through shared memory (profiles/ncu_r1_summary.txt: fp64 pipe 14 % busy).
// Here ONE THREAD owns one instance, so a warp advances 32 instances per instruction, nothing is exchanged
// between lanes, and the block-Toeplitz structure of the prediction matrix is used as what it is — the
// linear dynamics:   A x̃  = forward simulation  x_{k+1} = Ad x_k + Bd u_k          (Np*(nx^2+nx*nu) FMAs)
//                    A' w  = adjoint sweep       λ_k = w_k + Ad' λ_{k+1}, r_j = Bd' λ_{j+1}
// instead of two dense (Np*nx x Nc*nu) mat-vecs.  All shared matrices (Ad, Bd, K^-1, H^-1, G*) travel in the
// kernel-parameter constant bank (uniform operands, no load instructions); the iterate v lives in shared
// memory as v[row][lane] (conflict-free), x / r / x̃ in registers.  Every loop has compile-time bounds.
//
// Rows handled here: predicted states k = 1..Np (NS = Np*nx rows; the k = 0 block never touches U),
// inputs (NU), the reference's delta-u rows (ND).  The generic layout [B, mc] with mc = NS + nx + NU + ND is
// kept in global memory so the team kernels can take over any instance this path does not finish.
#pragma once
#include "bmpc_core.cuh"

template <int NXc, int NUc, int NPc, int NCc>
struct TpiShape {
    static constexpr int nx = NXc, nu = NUc, Np = NPc, Nc = NCc;
    static constexpr int NS = NPc * NXc;
    static constexpr int NU = NCc * NUc;
    static constexpr int ND = (NCc + 1) * NUc;
    static constexpr int MT = NS + NU + ND;
    static constexpr int NX = (NPc + 1) * NXc;
    static constexpr int mc = NX + NU + ND;
    static constexpr int RMAX = 24;                         // working-set capacity of the TPI polish
    static constexpr int T0 = 0, R0 = RMAX, S0 = 2 * RMAX;  // polish workspace rows: t/mu | R | S packed lower
    static constexpr int U0B = S0 + RMAX * (RMAX + 1) / 2;   // U0 = -H^-1 g
    static constexpr int UB = U0B + NCc * NUc;               // candidate U
    static constexpr int WS = UB + NCc * NUc;
    static constexpr int PROWS = (WS > MT + S0 ? WS : MT + S0);   // v* is staged at rows [S0, S0 + MT) (S is dead by then)
};

template <class S>
struct TpiCommon {
    double Ad[S::nx * S::nx], Bd[S::nx * S::nu];
    double Gx0[S::NU * S::nx], Gref[S::NU * S::nx], g0[S::NU], QDu[S::nu * S::nu];
    double xmin[S::nx], xmax[S::nx], c1x[S::nx], c2x[S::nx], rhox[S::nx];
    double umin[S::nu], umax[S::nu], rhou[S::nu];
    double dmin[S::nu], dmax[S::nu], rhod[S::nu];
    double sigma, alpha, inv_rho_e;      // inv_rho_e = 0 -> hard state rows
};

template <class S>
struct TpiAdmmParams {
    TpiCommon<S> c;
    double Kinv[S::NU * S::NU];
    double Gcc[S::NU * S::nx];           // B' R_x Acal : folds the affine offset of the state rows into g
};

template <class S>
struct TpiPolishParams {
    TpiCommon<S> c;
    double Hinv[S::NU * S::NU];
    const double* M;                      // [mc, mc]  A H^-1 A'   (generic row indexing)
    const double* AHinv;                  // [mc, NU]
};

// per-thread strided accessor: element i of this thread's private column
struct TpiAcc {
    double* p; int stride;
    BMPC_HD double& operator()(int i) const { return p[i * stride]; }
};

template <class S>
BMPC_HD double tpi_prox_x(const TpiCommon<S>& c, int a, double v) {
    const double lo = c.xmin[a], hi = c.xmax[a];
    return v > hi ? c.c1x[a] * v + c.c2x[a] * hi : (v < lo ? c.c1x[a] * v + c.c2x[a] * lo : v);
}
BMPC_HD double tpi_clamp(double v, double lo, double hi) { return v > hi ? hi : (v < lo ? lo : v); }

template <class S>
BMPC_HD void tpi_dbounds(const TpiCommon<S>& c, const double* um1, int rr, double& lo, double& hi) {
    lo = c.dmin[rr % S::nu]; hi = c.dmax[rr % S::nu];
    if (rr < S::nu) { lo += um1[rr]; hi += um1[rr]; }
}

// true linear term g (NU) of the condensed QP for this instance (constant xref)
template <class S>
BMPC_HD void tpi_linear_term(const TpiCommon<S>& c, const double* x0, const double* um1, const double* xref, double* g) {
#pragma unroll
    for (int a = 0; a < S::NU; a++) {
        double acc = c.g0[a];
#pragma unroll
        for (int q = 0; q < S::nx; q++) acc += c.Gx0[a * S::nx + q] * x0[q] + c.Gref[a * S::nx + q] * xref[q];
        if (a < S::nu) {
#pragma unroll
            for (int q = 0; q < S::nu; q++) acc -= c.QDu[a * S::nu + q] * um1[q];
        }
        g[a] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// niter ADMM iterations.  V: this thread's iterate v (MT rows).  x: NU registers (in/out).
template <class S>
BMPC_HD void tpi_admm(const TpiAdmmParams<S>& P, TpiAcc V, const double* x0, const double* um1, const double* xref,
                      double* x, int niter, bool cold) {
    constexpr int nx = S::nx, nu = S::nu, Np = S::Np, Nc = S::Nc, NS = S::NS, NU = S::NU, ND = S::ND;
    const TpiCommon<S>& c = P.c;
    double gp[NU];
    tpi_linear_term<S>(c, x0, um1, xref, gp);
#pragma unroll
    for (int a = 0; a < NU; a++) {
#pragma unroll
        for (int q = 0; q < nx; q++) gp[a] += P.Gcc[a * nx + q] * x0[q];
    }
    if (cold) {
        // x = 0, v = A x + cc : free response on the state rows, zeros elsewhere
        double xk[nx];
#pragma unroll
        for (int q = 0; q < nx; q++) xk[q] = x0[q];
#pragma unroll
        for (int k = 1; k <= Np; k++) {
            double xn[nx];
#pragma unroll
            for (int a = 0; a < nx; a++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < nx; q++) acc += c.Ad[a * nx + q] * xk[q];
                xn[a] = acc;
            }
#pragma unroll
            for (int a = 0; a < nx; a++) { xk[a] = xn[a]; V((k - 1) * nx + a) = xn[a]; }
        }
#pragma unroll
        for (int i = NS; i < S::MT; i++) V(i) = 0.0;
#pragma unroll
        for (int a = 0; a < NU; a++) x[a] = 0.0;
    }
#pragma unroll 1
    for (int it = 0; it < niter; it++) {
        double r[NU];
#pragma unroll
        for (int a = 0; a < NU; a++) r[a] = c.sigma * x[a] - gp[a];
        // input rows
#pragma unroll
        for (int a = 0; a < NU; a++) {
            double v = V(NS + a);
            double z = tpi_clamp(v, c.umin[a % nu], c.umax[a % nu]);
            r[a] += c.rhou[a % nu] * (2.0 * z - v);
        }
        // delta-u rows (reference quirk: scalar shift)
#pragma unroll
        for (int rr = 0; rr < ND; rr++) {
            double lo, hi; tpi_dbounds<S>(c, um1, rr, lo, hi);
            double v = V(NS + NU + rr);
            double w = c.rhod[rr % nu] * (2.0 * tpi_clamp(v, lo, hi) - v);
            if (rr < nu) r[rr] += w;
            else { r[rr - nu] -= w; if (rr - nu + 1 < NU) r[rr - nu + 1] += w; }
        }
        // state rows: adjoint sweep  lam_k = w_k + Ad' lam_{k+1} ; r_j += Bd' lam_{j+1}
        double lam[nx];
#pragma unroll
        for (int q = 0; q < nx; q++) lam[q] = 0.0;
#pragma unroll
        for (int k = Np; k >= 1; k--) {
            double ln[nx];
#pragma unroll
            for (int a = 0; a < nx; a++) {
                double v = V((k - 1) * nx + a);
                double acc = c.rhox[a] * (2.0 * tpi_prox_x<S>(c, a, v) - v);
#pragma unroll
                for (int q = 0; q < nx; q++) acc += c.Ad[q * nx + a] * lam[q];
                ln[a] = acc;
            }
#pragma unroll
            for (int a = 0; a < nx; a++) lam[a] = ln[a];
            const int j = (k - 1 < Nc - 1) ? (k - 1) : (Nc - 1);
#pragma unroll
            for (int b = 0; b < nu; b++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < nx; q++) acc += c.Bd[q * nu + b] * lam[q];
                r[j * nu + b] += acc;
            }
        }
        // xt = Kinv r
        double xt[NU];
#pragma unroll
        for (int a = 0; a < NU; a++) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < NU; b++) acc += P.Kinv[a * NU + b] * r[b];
            xt[a] = acc;
        }
        // forward simulation: zt on the state rows, v += alpha (zt - z)
        double xk[nx];
#pragma unroll
        for (int q = 0; q < nx; q++) xk[q] = x0[q];
#pragma unroll
        for (int k = 1; k <= Np; k++) {
            const int j = (k - 1 < Nc - 1) ? (k - 1) : (Nc - 1);
            double xn[nx];
#pragma unroll
            for (int a = 0; a < nx; a++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < nx; q++) acc += c.Ad[a * nx + q] * xk[q];
#pragma unroll
                for (int b = 0; b < nu; b++) acc += c.Bd[a * nu + b] * xt[j * nu + b];
                xn[a] = acc;
            }
#pragma unroll
            for (int a = 0; a < nx; a++) {
                xk[a] = xn[a];
                double v = V((k - 1) * nx + a);
                V((k - 1) * nx + a) = v + c.alpha * (xn[a] - tpi_prox_x<S>(c, a, v));
            }
        }
#pragma unroll
        for (int a = 0; a < NU; a++) {
            double v = V(NS + a);
            V(NS + a) = v + c.alpha * (xt[a] - tpi_clamp(v, c.umin[a % nu], c.umax[a % nu]));
        }
#pragma unroll
        for (int rr = 0; rr < ND; rr++) {
            double lo, hi; tpi_dbounds<S>(c, um1, rr, lo, hi);
            double v = V(NS + NU + rr);
            double zt = rr < nu ? xt[rr] : (-xt[rr - nu] + (rr - nu + 1 < NU ? xt[rr - nu + 1] : 0.0));
            V(NS + NU + rr) = v + c.alpha * (zt - tpi_clamp(v, lo, hi));
        }
#pragma unroll
        for (int a = 0; a < NU; a++) x[a] += c.alpha * (xt[a] - x[a]);
    }
}

// ------------------------------------------------------------------------------------------------
// TPI polish: same primal-dual active-set / KKT-verified scheme as bmpc_polish (bmpc_core.cuh), one thread
// per instance.  Sets are two 128-bit masks; the candidate's rows come from a forward simulation; the
// Schur matrix S (<= RMAX x RMAX, packed lower) lives in this thread's shared-memory column.
struct TpiMask {
    unsigned long long w[2];
    BMPC_HD TpiMask() { w[0] = 0ull; w[1] = 0ull; }
    // explicit selects (no dynamic indexing of w[]) so the masks stay in registers
    BMPC_HD bool get(int i) const { return (((i < 64) ? w[0] : w[1]) >> (i & 63)) & 1ull; }
    BMPC_HD void set(int i, bool b) {
        const unsigned long long bit = (b ? 1ull : 0ull) << (i & 63);
        if (i < 64) w[0] |= bit; else w[1] |= bit;
    }
};
BMPC_HD int tpi_popc(unsigned long long v) {
#ifdef BMPC_HOSTEMU
    return __builtin_popcountll(v);
#else
    return __popcll(v);
#endif
}
// number of set bits strictly below position i
BMPC_HD int tpi_rank(const TpiMask& a, int i) {
    if (i < 64) return tpi_popc(a.w[0] & ((1ull << i) - 1ull));
    return tpi_popc(a.w[0]) + tpi_popc(a.w[1] & ((1ull << (i - 64)) - 1ull));
}

// ---- row iterators (rolled over the horizon so the code stays small; inner nx / nu loops are unrolled) ----
// f(i, lo, hi, rho) for every TPI row i
template <class S, class F>
BMPC_HD void tpi_for_rows(const TpiCommon<S>& c, const double* um1, F f) {
    constexpr int nx = S::nx, nu = S::nu, Np = S::Np, NS = S::NS, NU = S::NU, ND = S::ND;
#pragma unroll 1
    for (int k = 0; k < Np; k++) {
#pragma unroll
        for (int a = 0; a < nx; a++) f(k * nx + a, c.xmin[a], c.xmax[a], c.rhox[a]);
    }
#pragma unroll 1
    for (int j = 0; j < NU / nu; j++) {
#pragma unroll
        for (int b = 0; b < nu; b++) f(NS + j * nu + b, c.umin[b], c.umax[b], c.rhou[b]);
    }
#pragma unroll
    for (int b = 0; b < nu; b++) f(NS + NU + b, c.dmin[b] + um1[b], c.dmax[b] + um1[b], c.rhod[b]);
#pragma unroll 1
    for (int j = 1; j < ND / nu; j++) {
#pragma unroll
        for (int b = 0; b < nu; b++) f(NS + NU + j * nu + b, c.dmin[b], c.dmax[b], c.rhod[b]);
    }
}

// f(i, lo, hi, rho, value) with value = (A U + cc)_i, U read from the workspace rows [ub, ub+NU)
template <class S, class F>
BMPC_HD void tpi_rows_of(const TpiCommon<S>& c, const double* x0, const double* um1, TpiAcc W, int ub, F f) {
    constexpr int nx = S::nx, nu = S::nu, Np = S::Np, Nc = S::Nc, NS = S::NS, NU = S::NU, ND = S::ND;
    double xk[nx];
#pragma unroll
    for (int q = 0; q < nx; q++) xk[q] = x0[q];
#pragma unroll 1
    for (int k = 1; k <= Np; k++) {
        const int j = (k - 1 < Nc - 1) ? (k - 1) : (Nc - 1);
        double uj[nu];
#pragma unroll
        for (int b = 0; b < nu; b++) uj[b] = W(ub + j * nu + b);
        double xn[nx];
#pragma unroll
        for (int a = 0; a < nx; a++) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q < nx; q += 2) { a0 += c.Ad[a * nx + q] * xk[q]; if (q + 1 < nx) a1 += c.Ad[a * nx + q + 1] * xk[q + 1]; }
#pragma unroll
            for (int b = 0; b < nu; b++) a1 += c.Bd[a * nu + b] * uj[b];
            xn[a] = a0 + a1;
        }
#pragma unroll
        for (int a = 0; a < nx; a++) { xk[a] = xn[a]; f((k - 1) * nx + a, c.xmin[a], c.xmax[a], c.rhox[a], xn[a]); }
    }
#pragma unroll 1
    for (int j = 0; j < NU / nu; j++) {
#pragma unroll
        for (int b = 0; b < nu; b++) f(NS + j * nu + b, c.umin[b], c.umax[b], c.rhou[b], W(ub + j * nu + b));
    }
    // delta-u rows: first nu rows = u_0 (bounds shifted by u_-1), then -U[s] + U[s+1] on the scalar stacking
#pragma unroll
    for (int b = 0; b < nu; b++) f(NS + NU + b, c.dmin[b] + um1[b], c.dmax[b] + um1[b], c.rhod[b], W(ub + b));
#pragma unroll 1
    for (int j = 1; j < ND / nu; j++) {
#pragma unroll
        for (int b = 0; b < nu; b++) {
            const int s2 = (j - 1) * nu + b;
            const double val = -W(ub + s2) + (s2 + 1 < NU ? W(ub + s2 + 1) : 0.0);
            f(NS + NU + j * nu + b, c.dmin[b], c.dmax[b], c.rhod[b], val);
        }
    }
}

// W: workspace accessor (rows: [T0,..) t/mu, [R0,..) R, [S0,..) S packed lower with 1/L_jj on the diagonal,
// [U0B,..) U0, [UB,..) candidate U).  up/dn: in = initial sets (from v), out = verified working set.
// Returns steps used (>0) when KKT-verified (U in rows UB.., multipliers by rank in rows T0..),
// 0 if not verified within max_steps, -1 if the working set outgrew RMAX.
template <class S>
BMPC_HD int tpi_polish(const TpiPolishParams<S>& P, TpiAcc W, const double* x0, const double* um1, const double* g,
                       TpiMask& up, TpiMask& dn, int max_steps) {
    constexpr int NS = S::NS, NU = S::NU, RMAX = S::RMAX, nx = S::nx, R0 = S::R0, S0 = S::S0, U0B = S::U0B, UB = S::UB;
    const TpiCommon<S>& c = P.c;
    const bool soft_on = c.inv_rho_e > 0.0;
#pragma unroll 1
    for (int a = 0; a < NU; a++) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int b = 0; b < NU; b += 4) {
            a0 += P.Hinv[a * NU + b] * g[b];
            if (b + 1 < NU) a1 += P.Hinv[a * NU + b + 1] * g[b + 1];
            if (b + 2 < NU) a2 += P.Hinv[a * NU + b + 2] * g[b + 2];
            if (b + 3 < NU) a3 += P.Hinv[a * NU + b + 3] * g[b + 3];
        }
        W(U0B + a) = -((a0 + a1) + (a2 + a3));
    }
#pragma unroll 1
    for (int step = 0; step < max_steps; step++) {
        TpiMask act; act.w[0] = up.w[0] | dn.w[0]; act.w[1] = up.w[1] | dn.w[1];
        const int r = tpi_popc(act.w[0]) + tpi_popc(act.w[1]);
        if (r > RMAX) return -1;
        // residual of the working rows at U0, and their generic row indices
        tpi_rows_of<S>(c, x0, um1, W, U0B, [&](int i, double lo, double hi, double, double val) {
            if (act.get(i)) {
                const int k = tpi_rank(act, i);
                W(k) = val - (up.get(i) ? hi : lo);
                W(R0 + k) = (double)(i + nx);
            }
        });
        // S = M[R,R] + diag (packed lower), then Cholesky in place with 1/L_jj on the diagonal
        for (int p = 0; p < r; p++) {
            const int Rp = (int)W(R0 + p);
            const double* Mrow = P.M + (size_t)Rp * S::mc;
            for (int q = 0; q <= p; q++) {
                double val = Mrow[(int)W(R0 + q)];
                if (p == q) val += (soft_on && Rp < S::NX) ? c.inv_rho_e : 1e-13 * (1.0 + fabs(val));
                W(S0 + p * (p + 1) / 2 + q) = val;
            }
        }
        for (int i = 0; i < r; i++) {
            const int bi = S0 + i * (i + 1) / 2;
            for (int j = 0; j <= i; j++) {
                const int bj = S0 + j * (j + 1) / 2;
                double s0 = W(bi + j), s1 = 0.0;
                int k = 0;
                for (; k + 1 < j; k += 2) { s0 -= W(bi + k) * W(bj + k); s1 -= W(bi + k + 1) * W(bj + k + 1); }
                if (k < j) s0 -= W(bi + k) * W(bj + k);
                double s = s0 + s1;
                if (j < i) W(bi + j) = s * W(bj + j);
                else { if (!(s > 1e-300)) s = 1e-300; W(bi + i) = 1.0 / sqrt(s); }
            }
        }
        for (int i = 0; i < r; i++) {                         // L y = t
            const int bi = S0 + i * (i + 1) / 2;
            double s0 = W(i), s1 = 0.0;
            int k = 0;
            for (; k + 1 < i; k += 2) { s0 -= W(bi + k) * W(k); s1 -= W(bi + k + 1) * W(k + 1); }
            if (k < i) s0 -= W(bi + k) * W(k);
            W(i) = (s0 + s1) * W(bi + i);
        }
        double mumax = 0.0;
        for (int i = r - 1; i >= 0; i--) {                    // L' mu = y
            double s = W(i);
            for (int k = i + 1; k < r; k++) s -= W(S0 + k * (k + 1) / 2 + i) * W(k);
            s *= W(S0 + i * (i + 1) / 2 + i);
            W(i) = s; mumax = fmax(mumax, fabs(s));
        }
        // candidate U = U0 - (A Hinv)[R,:]' mu
        {
            double Ur[NU];
#pragma unroll
            for (int a = 0; a < NU; a++) Ur[a] = W(U0B + a);
            for (int p = 0; p < r; p++) {
                const double* row = P.AHinv + (size_t)((int)W(R0 + p)) * NU; const double mu = W(p);
#pragma unroll
                for (int a = 0; a < NU; a++) Ur[a] -= row[a] * mu;
            }
#pragma unroll
            for (int a = 0; a < NU; a++) W(UB + a) = Ur[a];
        }
        // KKT verification + next sets
        bool ok = true;
        const double mutol = 1e-9 * (1.0 + mumax);
        TpiMask nup, ndn;
        tpi_rows_of<S>(c, x0, um1, W, UB, [&](int i, double lo, double hi, double, double zi) {
            const bool su = up.get(i), sd = dn.get(i);
            bool nu_, nd_;
            if (soft_on && i < NS) {
                nu_ = zi > hi + 1e-11 * (1.0 + fabs(hi)); nd_ = (!nu_) && zi < lo - 1e-11 * (1.0 + fabs(lo));
                if (nu_ != su || nd_ != sd) {
                    const bool hi_side = (su || nu_) && !(sd || nd_), lo_side = (sd || nd_) && !(su || nu_);
                    const double gap = hi_side ? fabs(zi - hi) : (lo_side ? fabs(zi - lo) : 1e300);
                    const double bnd = hi_side ? hi : lo;
                    if (!(gap <= 1e-11 * (1.0 + fabs(bnd)))) ok = false;
                }
            } else {
                const double mu = (su || sd) ? W(tpi_rank(act, i)) : 0.0;
                const bool vu = zi > hi + 1e-9 * (1.0 + fabs(hi)), vd = zi < lo - 1e-9 * (1.0 + fabs(lo));
                if (vu || vd || (su && mu < -mutol) || (sd && mu > mutol)) ok = false;
                nu_ = vu || (!vd && su && mu > 0.0);
                nd_ = (!nu_) && (vd || (sd && mu < 0.0));
            }
            nup.set(i, nu_); ndn.set(i, nd_);
        });
        if (ok) return step + 1;      // `up`/`dn` still describe the verified working set; W(0..r) = mu by rank
        up = nup; dn = ndn;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host side: fill the parameter blocks from a host copy of the condensed system block
template <class S>
inline void tpi_fill_common(const double* sys, const BmpcSysOff& o, TpiCommon<S>& c) {
    constexpr int nx = S::nx, nu = S::nu, NU = S::NU, NX = S::NX;
    for (int i = 0; i < nx * nx; i++) c.Ad[i] = sys[o.Ad + i];
    for (int i = 0; i < nx * nu; i++) c.Bd[i] = sys[o.Bd + i];
    for (int i = 0; i < NU * nx; i++) { c.Gx0[i] = sys[o.Gx0 + i]; c.Gref[i] = sys[o.Gref + i]; }
    for (int i = 0; i < NU; i++) c.g0[i] = sys[o.g0 + i];
    for (int i = 0; i < nu * nu; i++) c.QDu[i] = sys[o.QDu + i];
    const double rho_e = sys[o.scal + BMPC_S_RHOE];
    c.inv_rho_e = rho_e > 0.0 ? 1.0 / rho_e : 0.0;
    c.sigma = sys[o.scal + BMPC_S_SIGMA]; c.alpha = sys[o.scal + BMPC_S_ALPHA];
    for (int a = 0; a < nx; a++) {
        c.xmin[a] = sys[o.lo0 + nx + a]; c.xmax[a] = sys[o.hi0 + nx + a];
        const double rho = sys[o.rho + nx + a]; c.rhox[a] = rho;
        if (rho_e > 0.0) { c.c1x[a] = rho / (rho + rho_e); c.c2x[a] = rho_e / (rho + rho_e); }
        else { c.c1x[a] = 0.0; c.c2x[a] = 1.0; }
    }
    for (int b = 0; b < nu; b++) {
        c.umin[b] = sys[o.lo0 + NX + b]; c.umax[b] = sys[o.hi0 + NX + b]; c.rhou[b] = sys[o.rho + NX + b];
        c.dmin[b] = sys[o.lo0 + NX + NU + b]; c.dmax[b] = sys[o.hi0 + NX + NU + b]; c.rhod[b] = sys[o.rho + NX + NU + b];
    }
}
template <class S>
inline void tpi_fill_admm(const double* sys, const BmpcSysOff& o, TpiAdmmParams<S>& P) {
    constexpr int nx = S::nx, NU = S::NU, NX = S::NX;
    tpi_fill_common<S>(sys, o, P.c);
    for (int i = 0; i < NU * NU; i++) P.Kinv[i] = sys[o.Kinv + i];
    for (int a = 0; a < NU; a++)
        for (int q = 0; q < nx; q++) {
            double acc = 0.0;
            for (int i = 0; i < NX; i++) acc += sys[o.Bcal + i * NU + a] * sys[o.rho + i] * sys[o.Acal + i * nx + q];
            P.Gcc[a * nx + q] = acc;
        }
}
template <class S>
inline void tpi_fill_polish(const double* sys, const BmpcSysOff& o, const double* dev_sys, TpiPolishParams<S>& P) {
    tpi_fill_common<S>(sys, o, P.c);
    for (int i = 0; i < S::NU * S::NU; i++) P.Hinv[i] = sys[o.Hinv + i];
    P.M = dev_sys + o.M; P.AHinv = dev_sys + o.AHinv;
}
