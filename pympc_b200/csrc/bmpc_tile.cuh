// bmpc_tile.cuh — K4 for large shapes (MIMO nx=8, nu=4, Np=40: NU=160, mc=652): one CTA iterates a TILE of T instances
// together.  Same iteration and same residual bookkeeping as bmpc_admm_rows (bmpc_core.cuh), reorganised so that the
// shared matrices are read once per tile instead of once per instance:
//   * vectors of the T instances are interleaved in shared memory, element-major: vec[e*T + t]; a thread that owns one
//     output row reads its matrix entry once and the T right-hand-side values as one broadcast, T FMAs per entry;
//   * the prediction matrix Bcal is block-Toeplitz (block (k,j) = Ad^(k-1-j) Bd, plus the "held input" block column
//     when Nc < Np, mpc.py:537-544): its Np (+ Np-Nc+1) distinct nx*nu blocks live in shared memory (10 KB for MIMO
//     instead of 2 x 420 KB streamed from L2 per instance-iteration);
//   * K^-1 (symmetric, NU x NU, one copy per adaptive-rho level) is the only matrix streamed from L2: NU^2*8 bytes per
//     TILE-iteration; instances of a tile that sit on different levels are served level by level.
// Written against the team abstraction (tid, n, sync) so that tests/hostemu runs the same code on the host.
#pragma once
#include "bmpc_core.cuh"

// strides of one nx*nu block in the two shared-memory copies of the Toeplitz generator:
//   copy 1 [q][c] (rows of Bcal: consecutive threads = consecutive state rows c), copy 2 [c][q] (columns: consecutive inputs q)
BMPC_HOSTDEV int bmpc_tile_stride1(const BmpcDims& d) { int s = d.nx * d.nu; return s + ((8 - s % 16) + 16) % 16; }
BMPC_HOSTDEV int bmpc_tile_stride2(const BmpcDims& d) { int s = d.nx * d.nu; return s + ((d.nu % 16 - s % 16) + 16) % 16; }
BMPC_HOSTDEV int bmpc_tile_nblocks(const BmpcDims& d) { return d.Np + (d.Np - d.Nc + 1); }
BMPC_HOSTDEV size_t bmpc_tile_phi_doubles(const BmpcDims& d) {
    size_t n = (size_t)bmpc_tile_nblocks(d) * (bmpc_tile_stride1(d) + bmpc_tile_stride2(d)) + 3 * (size_t)d.mc;   // + row bounds and rho
    return (n + 1) & ~(size_t)1;                       // even: the tile vectors behind it stay 16-byte aligned
}
// per-instance doubles: g, x, xt, r (NU each), cc (NX), v, w (mc each), um1 (nu), res (4)
BMPC_HOSTDEV size_t bmpc_tile_inst_doubles(const BmpcDims& d) { return 4 * (size_t)d.NU + d.NX + 2 * (size_t)d.mc + d.nu + 4; }
BMPC_HOSTDEV size_t bmpc_tile_smem_doubles(const BmpcDims& d, int T) {
    return bmpc_tile_phi_doubles(d) + (size_t)T * bmpc_tile_inst_doubles(d) + 2 * (size_t)T /* lvl, new lvl, instance ids as ints */ + 2;
}

// index of block (k, j) of Bcal (state stage k in 1..Np, input block j < min(k, Nc)) in the generator tables
BMPC_HD int bmpc_tile_block(const BmpcDims& d, int k, int j) { return (j < d.Nc - 1) ? (k - 1 - j) : (d.Np + k - d.Nc); }

template <class Team>
BMPC_HD void bmpc_tile_load_rows(Team& t, const BmpcDims& d, const BmpcSysOff& o, const double* sys, double* lo, double* hi, double* rho) {
    for (int i = t.tid; i < d.mc; i += t.n) { lo[i] = sys[o.lo0 + i]; hi[i] = sys[o.hi0 + i]; rho[i] = sys[o.rho + i]; }
}

template <class Team>
BMPC_HD void bmpc_tile_load_phi(Team& t, const BmpcDims& d, const double* Bcal, double* phi1, double* phi2) {
    const int nb = bmpc_tile_nblocks(d), s1 = bmpc_tile_stride1(d), s2 = bmpc_tile_stride2(d), bs = d.nx * d.nu;
    for (int idx = t.tid; idx < nb * bs; idx += t.n) {
        const int b = idx / bs, e = idx % bs, c = e / d.nu, q = e % d.nu;
        int k, j;
        if (b < d.Np) { k = b + 1; j = 0; } else { k = d.Nc + (b - d.Np); j = d.Nc - 1; }
        const double val = Bcal[(size_t)(k * d.nx + c) * d.NU + j * d.nu + q];
        phi1[b * s1 + q * d.nx + c] = val;
        phi2[b * s2 + c * d.nu + q] = val;
    }
}

// N consecutive doubles of a tile vector (16-byte aligned by construction) -> registers, as 128-bit shared-memory loads
#ifdef BMPC_HOSTEMU
template <int N> BMPC_HD void bmpc_ldv(const double* p, double* o) { for (int i = 0; i < N; i++) o[i] = p[i]; }
#else
template <int N> BMPC_HD void bmpc_ldv(const double* p, double* o) {
    static_assert(N % 2 == 0, "tile vectors are read in pairs");
#pragma unroll
    for (int i = 0; i < N; i += 2) { const double2 t2 = *reinterpret_cast<const double2*>(p + i); o[i] = t2.x; o[i + 1] = t2.y; }
}
#endif

#ifdef BMPC_HOSTEMU
#define BMPC_TILE_MAX(slot, val) do { if ((val) > *(slot)) *(slot) = (val); } while (0)
#else
// non-negative doubles order like their bit patterns
#define BMPC_TILE_MAX(slot, val) atomicMax((unsigned long long*)(slot), (unsigned long long)__double_as_longlong(val))
#endif

// nearest ladder level suggested by OSQP's rule (see bmpc_adapt_level)
BMPC_HD int bmpc_next_level(const double* res, int lvl) {
    const double rp = res[0] / fmax(res[2], 1e-12), rd = res[1] / fmax(res[3], 1e-12);
    int nl = lvl;
    if (rp > 0.0 && rd > 0.0) {
        const double steps = log10(rp / rd);
        int mv = (int)(steps >= 0.0 ? steps + 0.5 : steps - 0.5);
        if (mv >= 1 || mv <= -1) nl = lvl + mv;
        if (nl < 0) nl = 0;
        if (nl > BMPC_NLEV - 1) nl = BMPC_NLEV - 1;
    }
    return nl;
}

// Shared-memory view of one tile
template <int T>
struct BmpcTile {
    double *phi1, *phi2;                       // Toeplitz generator, two layouts
    double *lo, *hi, *rho;                     // [mc] row bounds (without the u_-1 shift) and base rho of every row
    double *g, *cc, *x, *v, *w, *xt, *r;       // [elem][T]
    double *um1;                               // [T][nu]
    double *res;                               // [T][4]   rp, rd, np, nd of the last iteration
    double *fac;                               // [T]      rho factor of each instance's ladder level
    int *lvl, *nlvl, *inst;                    // [T]
    BMPC_HD void carve(double* base, const BmpcDims& d) {
        const int nb = bmpc_tile_nblocks(d);
        phi1 = base; phi2 = phi1 + (size_t)nb * bmpc_tile_stride1(d);
        lo = phi2 + (size_t)nb * bmpc_tile_stride2(d); hi = lo + d.mc; rho = hi + d.mc;
        g = base + bmpc_tile_phi_doubles(d); cc = g + (size_t)d.NU * T; x = cc + (size_t)d.NX * T;
        v = x + (size_t)d.NU * T; w = v + (size_t)d.mc * T; xt = w + (size_t)d.mc * T; r = xt + (size_t)d.NU * T;
        um1 = r + (size_t)d.NU * T; res = um1 + (size_t)d.nu * T;
        fac = res + 4 * T;
        lvl = (int*)(fac + T); nlvl = lvl + T; inst = nlvl + T;
    }
};

// `niter` ADMM iterations for the T instances of a tile.  NS = how many threads share one output row (each takes
// TG = T/NS instances).  On return S.res holds the residuals of the last iteration (as bmpc_admm_rows reports them).
// Three barriers per iteration:
//   B   r  = sigma x - g + A' w                (x first takes the relaxation step of the previous iteration)
//   C   xt = Kinv[level] r
//   D   zt = A xt + cc row by row, in registers; v += alpha (zt - prox(v)); then straight away the next iteration's
//       w = rho (2 prox(v) - v - cc) for the same element (no separate pass, no extra barrier)
// NXC, NUC: compile-time copies of d.nx, d.nu (0 = use the run-time values): the nx / nu inner loops unroll, so their
// shared-memory loads issue back to back instead of one load -> FMA round trip per step.
template <int T, int NS, int NXC, int NUC, class Team>
BMPC_HD void bmpc_admm_tile(Team& t, const BmpcDims& d, const BmpcSysOff& o, const double* sys, BmpcTile<T>& S, int niter) {
    constexpr int TG = T / NS;
    const double *H = sys + o.H, *lo0 = S.lo, *hi0 = S.hi, *rhov = S.rho, *scal = sys + o.scal;
    const double sigma = scal[BMPC_S_SIGMA], alpha = scal[BMPC_S_ALPHA], rho_e = scal[BMPC_S_RHOE];
    const bool soft_on = rho_e > 0.0;
    const int nx = NXC ? NXC : d.nx, nu = NUC ? NUC : d.nu, Np = d.Np, Nc = d.Nc, NX = d.NX, NU = d.NU, mc = d.mc;
    const int s1 = bmpc_tile_stride1(d), s2 = bmpc_tile_stride2(d);
    double *g = S.g, *cc = S.cc, *x = S.x, *v = S.v, *w = S.w, *xt = S.xt, *r = S.r;

    // (A' w)_a for TG instances: Toeplitz part over the state rows + input row + the reference's delta-u rows
    auto ATw = [&](const double* ww, int a, int sg, double* acc) {
        const int j = a / nu, q = a % nu;
        for (int e = 0; e < TG; e++) acc[e] = 0.0;
        for (int k = Np; k > j; k--) {                     // descending: every lane of a warp reads the same w rows (broadcast)
            const double* blk = S.phi2 + bmpc_tile_block(d, k, j) * s2 + q;
            const double* wk = ww + (k * nx) * T + sg * TG;
#pragma unroll
            for (int c = 0; c < nx; c++) {
                const double m = blk[c * nu]; double wv[TG];
                bmpc_ldv<TG>(wk + c * T, wv);
#pragma unroll
                for (int e = 0; e < TG; e++) acc[e] = fma(m, wv[e], acc[e]);
            }
        }
        const double* wu = ww + (NX + a) * T + sg * TG;
        const double* wd = ww + (NX + NU) * T + sg * TG;
#pragma unroll
        for (int e = 0; e < TG; e++) {
            double sm = wu[e] - wd[(nu + a) * T + e];
            if (a < nu) sm += wd[a * T + e];
            if (a >= 1) sm += wd[(nu + a - 1) * T + e];
            acc[e] += sm;
        }
    };
    // out_a = sum_b Msym[b*NU + a] in[b]   (symmetric matrix streamed from global memory, coalesced over a; the next 8
    // entries are requested before the current 8 are consumed, so the L2 latency overlaps the FMAs)
    auto symv = [&](const double* Msym, const double* in, int a, int sg, double* acc) {
        for (int e = 0; e < TG; e++) acc[e] = 0.0;
        const double* pm = Msym + a;
        const double* iv = in + sg * TG;
        const int ngrp = NU / 8;
        double m0[8], m1[8];
        auto fetch = [&](double* mm) {
#pragma unroll
            for (int u = 0; u < 8; u++) mm[u] = pm[u * NU];
            pm += 8 * NU;
        };
        auto use = [&](const double* mm) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                double rv[TG];
                bmpc_ldv<TG>(iv + u * T, rv);
#pragma unroll
                for (int e = 0; e < TG; e++) acc[e] = fma(mm[u], rv[e], acc[e]);
            }
            iv += 8 * T;
        };
        // ping-pong between two register sets: the next group is in flight while the current one is consumed
        if (ngrp > 0) fetch(m0);
        int gi = 0;
        for (; gi + 2 <= ngrp; gi += 2) {
            fetch(m1); use(m0);
            if (gi + 2 < ngrp) fetch(m0);
            use(m1);
        }
        if (gi < ngrp) use(m0);
        for (int b = ngrp * 8; b < NU; b++, pm += NU, iv += T) {
            const double m0 = pm[0]; double rv[TG];
            bmpc_ldv<TG>(iv, rv);
#pragma unroll
            for (int e = 0; e < TG; e++) acc[e] = fma(m0, rv[e], acc[e]);
        }
    };
    auto row_prox = [&](int i, int e, double vi, double& rho) {
        double lo = lo0[i], hi = hi0[i];
        const int rr = i - (NX + NU);
        if (rr >= 0 && rr < nu) { const double um = S.um1[e * nu + rr]; lo += um; hi += um; }
        rho = S.fac[e] * rhov[i];
        return bmpc_prox(vi, lo, hi, soft_on && i < NX, rho, rho_e);
    };
    // one row element: relaxation step of v with this iteration's zt, then either the next iteration's w or (last
    // iteration) the residual bookkeeping
    auto row_step = [&](int i, int e, int p, double zt, bool last) {
        const double vi = v[p]; double rho;
        const double z = row_prox(i, e, vi, rho);
        const double dz = zt - z, vn = vi + alpha * dz;
        v[p] = vn;
        if (last) {
            BMPC_TILE_MAX(S.res + e * 4 + 0, fabs(dz));
            BMPC_TILE_MAX(S.res + e * 4 + 2, fmax(fabs(zt), fabs(z)));
            w[p] = rho * dz;
        } else {
            const double zn = row_prox(i, e, vn, rho);
            w[p] = rho * (2.0 * zn - vn - (i < NX ? cc[p] : 0.0));
        }
    };

    for (int e = t.tid; e < 4 * T; e += t.n) S.res[e] = 0.0;
    for (int e = t.tid; e < T; e += t.n) S.fac[e] = bmpc_level_factor(S.lvl[e]);
    t.sync();
    // levels present in this tile
    unsigned lmask = 0;
    for (int e = 0; e < T; e++) lmask |= 1u << S.lvl[e];
    // w of the first iteration
    for (int idx = t.tid; idx < mc * T; idx += t.n) {
        const int i = idx / T, e = idx % T;
        const double vi = v[idx]; double rho;
        const double z = row_prox(i, e, vi, rho);
        w[idx] = rho * (2.0 * z - vi - (i < NX ? cc[idx] : 0.0));
    }
    t.sync();

    for (int it = 0; it < niter; it++) {
        const bool last = (it == niter - 1);
        // B: r = sigma x - g + A' w
        for (int wk = t.tid; wk < NU * NS; wk += t.n) {
            const int a = wk / NS, sg = wk % NS; double acc[TG];
            ATw(w, a, sg, acc);
#pragma unroll
            for (int e = 0; e < TG; e++) {
                const int p = a * T + sg * TG + e;
                double xa = x[p];
                if (it > 0) { xa += alpha * (xt[p] - xa); x[p] = xa; }
                r[p] = sigma * xa - g[p] + acc[e];
            }
        }
        t.sync();
        // C: xt = Kinv[level] r, level by level
        for (int L = 0; L < BMPC_NLEV; L++) {
            if (!((lmask >> L) & 1u)) continue;
            const double* Kinv = sys + o.KinvL + (size_t)L * NU * NU;
            for (int wk = t.tid; wk < NU * NS; wk += t.n) {
                const int a = wk / NS, sg = wk % NS; double acc[TG];
                symv(Kinv, r, a, sg, acc);
#pragma unroll
                for (int e = 0; e < TG; e++) if (S.lvl[sg * TG + e] == L) xt[a * T + sg * TG + e] = acc[e];
            }
        }
        t.sync();
        // D: state rows: zt = cc + sum_{j<min(k,Nc)} block(k,j)[c,:] xt_j in registers, then the row step
        for (int wk = t.tid; wk < NX * NS; wk += t.n) {
            const int i = wk / NS, sg = wk % NS, k = i / nx, c = i % nx;
            double acc[TG];
            bmpc_ldv<TG>(cc + i * T + sg * TG, acc);
            const int jend = k < Nc ? k : Nc;
            for (int j = 0; j < jend; j++) {
                const double* blk = S.phi1 + bmpc_tile_block(d, k, j) * s1 + c;
                const double* xj = xt + (j * nu) * T + sg * TG;
#pragma unroll
                for (int q = 0; q < nu; q++) {
                    const double m = blk[q * nx]; double xv[TG];
                    bmpc_ldv<TG>(xj + q * T, xv);
#pragma unroll
                    for (int e = 0; e < TG; e++) acc[e] = fma(m, xv[e], acc[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < TG; e++) row_step(i, sg * TG + e, i * T + sg * TG + e, acc[e], last);
        }
        //    input rows and the reference's delta-u rows
        for (int idx = NX * T + t.tid; idx < mc * T; idx += t.n) {
            const int i = idx / T, e = idx % T;
            double zt;
            if (i < NX + NU) zt = xt[(i - NX) * T + e];
            else {
                const int rr = i - NX - NU;
                if (rr < nu) zt = xt[rr * T + e];
                else { const int q2 = rr - nu; zt = -xt[q2 * T + e] + (q2 + 1 < NU ? xt[(q2 + 1) * T + e] : 0.0); }
            }
            row_step(i, e, idx, zt, last);
        }
        t.sync();
    }
    // residuals of the last iteration: r_dual = ||sigma (xt - x) + A' rho (zt - z)||, normalisers ||g||, ||H xt||, ||A' y||
    for (int wk = t.tid; wk < NU * NS; wk += t.n) {
        const int a = wk / NS, sg = wk % NS; double acc[TG], hx[TG];
        ATw(w, a, sg, acc);
        symv(H, xt, a, sg, hx);
#pragma unroll
        for (int e = 0; e < TG; e++) {
            const int p = a * T + sg * TG + e;
            BMPC_TILE_MAX(S.res + (sg * TG + e) * 4 + 1, fabs(sigma * (xt[p] - x[p]) + acc[e]));
            BMPC_TILE_MAX(S.res + (sg * TG + e) * 4 + 3, fmax(fabs(hx[e]), fabs(g[p])));
        }
    }
    t.sync();
    for (int idx = t.tid; idx < mc * T; idx += t.n) {
        const int i = idx / T, e = idx % T;
        const double vi = v[idx]; double rho;
        const double z = row_prox(i, e, vi, rho);
        w[idx] = rho * (vi - z);
    }
    t.sync();
    for (int wk = t.tid; wk < NU * NS; wk += t.n) {
        const int a = wk / NS, sg = wk % NS; double acc[TG];
        ATw(w, a, sg, acc);
#pragma unroll
        for (int e = 0; e < TG; e++) BMPC_TILE_MAX(S.res + (sg * TG + e) * 4 + 3, fabs(acc[e]));
    }
    for (int idx = t.tid; idx < NU * T; idx += t.n) x[idx] += alpha * (xt[idx] - x[idx]);
    t.sync();
}

// adaptive-rho move for every instance of the tile (v rescaled so that (z, y) are unchanged); S.nlvl = new levels
template <int T, class Team>
BMPC_HD void bmpc_tile_adapt(Team& t, const BmpcDims& d, const BmpcSysOff& o, const double* sys, BmpcTile<T>& S) {
    const double *lo0 = S.lo, *hi0 = S.hi, *rhov = S.rho;
    const double rho_e = sys[o.scal + BMPC_S_RHOE];
    const bool soft_on = rho_e > 0.0;
    for (int e = t.tid; e < T; e += t.n) S.nlvl[e] = bmpc_next_level(S.res + e * 4, S.lvl[e]);
    t.sync();
    for (int idx = t.tid; idx < d.mc * T; idx += t.n) {
        const int i = idx / T, e = idx % T;
        if (S.nlvl[e] == S.lvl[e]) continue;
        const double fo = bmpc_level_factor(S.lvl[e]), ratio = fo / bmpc_level_factor(S.nlvl[e]);
        double lo, hi; bmpc_row_bounds(d, lo0, hi0, S.um1 + e * d.nu, i, lo, hi);
        const double vi = S.v[idx];
        const double z = bmpc_prox(vi, lo, hi, soft_on && i < d.NX, fo * rhov[i], rho_e);
        S.v[idx] = z + (vi - z) * ratio;
    }
    t.sync();
}

// K3 for a tile: g and cc of every instance (same formulas as bmpc_prep)
template <int T, class Team>
BMPC_HD void bmpc_tile_prep(Team& t, const BmpcDims& d, const BmpcSysOff& o, const double* sys, BmpcTile<T>& S,
                            const double* x0 /*[T][nx]*/, const double* xref_all /*[B][xl]*/, int xref_mode) {
    const double *Gx0 = sys + o.Gx0, *Gref = sys + o.Gref, *GrefFull = sys + o.GrefFull, *g0 = sys + o.g0, *QDu = sys + o.QDu;
    const double* Acal = sys + o.Acal;
    for (int idx = t.tid; idx < d.NU * T; idx += t.n) {
        const int a = idx / T, e = idx % T;
        const double* xe = x0 + e * d.nx; const double* xr = xref_all + (size_t)S.inst[e] * (xref_mode ? d.NX : d.nx);
        double acc = g0[a];
        for (int c = 0; c < d.nx; c++) acc += Gx0[a * d.nx + c] * xe[c];
        if (xref_mode == 0) { for (int c = 0; c < d.nx; c++) acc += Gref[a * d.nx + c] * xr[c]; }
        else { for (int i = 0; i < d.NX; i++) acc += GrefFull[(size_t)a * d.NX + i] * xr[i]; }
        if (a < d.nu) { for (int q = 0; q < d.nu; q++) acc -= QDu[a * d.nu + q] * S.um1[e * d.nu + q]; }
        S.g[idx] = acc;
    }
    for (int idx = t.tid; idx < d.NX * T; idx += t.n) {
        const int i = idx / T, e = idx % T;
        const double* xe = x0 + e * d.nx; double acc = 0.0;
        for (int c = 0; c < d.nx; c++) acc += Acal[i * d.nx + c] * xe[c];
        S.cc[idx] = acc;
    }
    t.sync();
}
