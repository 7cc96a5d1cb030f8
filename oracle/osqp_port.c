/*
 * osqp_port.c — CPU restatement of the OSQP ADMM solver, as pyMPC uses it.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): the checker and the CPU baseline, never the
 * product.  The reference delegates its whole solve to the third-party `osqp` package
 * (/root/reference/setup.py:11, unpinned; call sites /root/reference/pyMPC/mpc.py:241,266,369,454),
 * whose source is NOT under /root/reference and which cannot be installed here.  This file
 * restates the published algorithm — Stellato, Banjac, Goulart, Bemporad, Boyd, "OSQP: an
 * operator splitting solver for quadratic programs", Math. Prog. Comp. 2020: Algorithm 1
 * (ADMM step), §3.4 (termination), §3.4 infeasibility certificates, §5.1 (Ruiz scaling, done by
 * the Python wrapper), §5.2 (adaptive rho) — with OSQP 0.6.x default settings.
 * "Parity unpinned" against real OSQP output: OSQP's default adaptive-rho interval is chosen
 * from wall-clock timings and is not reproducible; here it is a fixed iteration count.
 *
 * The linear system  [[P+sigma I, A'],[A, -diag(1/rho)]]  is factored L D L' with an
 * elimination-tree, up-looking sparse factorisation (the textbook algorithm QDLDL also
 * implements; Davis, "Direct Methods for Sparse Linear Systems", ch. 4).  The fill-reducing
 * permutation is computed by the Python wrapper.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC osqp_port.c -o libosqp_port.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OP_SOLVED 1
#define OP_MAX_ITER (-2)
#define OP_PRIMAL_INFEASIBLE (-3)
#define OP_DUAL_INFEASIBLE (-4)
#define OP_UNSOLVED (-10)
#define OP_INFTY 1e30
#define OP_RHO_MIN 1e-6
#define OP_RHO_MAX 1e6
#define OP_RHO_TOL 1e-4
#define OP_RHO_EQ_FACTOR 1e3
#define OP_MIN_SCALING 1e-4

typedef struct {
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int max_iter, check_termination, adaptive_rho, adaptive_rho_interval, warm_start;
    double adaptive_rho_tolerance;
} op_settings;

typedef struct {
    int n, m, N;
    /* scaled problem data (CSC; P upper triangle) */
    int *Pp, *Pi; double *Px;
    int *Ap, *Ai; double *Ax;
    double *q, *l, *u;
    double *D, *E, c; /* scaling: x = D xbar, z = E^-1 zbar, y = E ybar / c */
    /* permuted upper-triangular KKT */
    int *Kp, *Ki; double *Kx; int *perm; int *rho_idx;
    /* LDL' */
    int *etree, *Lnz, *Lp, *Li; double *Lx, *Dg, *Dinv;
    int *iwork; unsigned char *bwork; double *fwork;
    /* iterates (scaled) */
    double *x, *z, *y, *xprev, *zprev, *sol, *rhs, *dx, *dy, *Axv, *Pxv, *Atyv, *tmpn, *tmpm;
    double *rho_vec, *rho_inv;
    op_settings s;
    /* info */
    int iter, status, rho_updates; double obj_val, pri_res, dua_res;
    /* unscaled solution */
    double *xout, *yout;
} op_work;

/* ---------- sparse helpers ---------- */
static void csc_mv(int ncol, const int *p, const int *i, const double *x, const double *v, double *out, int nrow) {
    memset(out, 0, sizeof(double) * nrow);
    for (int c = 0; c < ncol; c++) { double vc = v[c]; for (int k = p[c]; k < p[c + 1]; k++) out[i[k]] += x[k] * vc; }
}
static void csc_tmv(int ncol, const int *p, const int *i, const double *x, const double *v, double *out) {
    for (int c = 0; c < ncol; c++) { double s = 0; for (int k = p[c]; k < p[c + 1]; k++) s += x[k] * v[i[k]]; out[c] = s; }
}
static void sym_upper_mv(int n, const int *p, const int *i, const double *x, const double *v, double *out) {
    memset(out, 0, sizeof(double) * n);
    for (int c = 0; c < n; c++) for (int k = p[c]; k < p[c + 1]; k++) {
        int r = i[k]; out[r] += x[k] * v[c]; if (r != c) out[c] += x[k] * v[r];
    }
}
static double inf_norm(const double *v, int n) { double m = 0; for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > m) m = a; } return m; }
static double inf_norm_scaled(const double *s, const double *v, int n, int inv) {
    double m = 0; for (int i = 0; i < n; i++) { double a = fabs(inv ? v[i] / s[i] : v[i] * s[i]); if (a > m) m = a; } return m;
}

/* ---------- L D L' : elimination tree, symbolic counts, up-looking numeric ---------- */
static int ldl_etree(int n, const int *Ap, const int *Ai, int *work, int *Lnz, int *etree) {
    for (int i = 0; i < n; i++) { work[i] = 0; Lnz[i] = 0; etree[i] = -1; if (Ap[i] == Ap[i + 1]) return -1; }
    for (int j = 0; j < n; j++) {
        work[j] = j;
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p]; if (i > j) return -1;
            while (work[i] != j) { if (etree[i] == -1) etree[i] = j; Lnz[i]++; work[i] = j; i = etree[i]; }
        }
    }
    int sum = 0; for (int i = 0; i < n; i++) sum += Lnz[i];
    return sum;
}
static int ldl_factor(op_work *w) {
    const int n = w->N; const int *Ap = w->Kp, *Ai = w->Ki; const double *Ax = w->Kx;
    int *Lp = w->Lp, *Li = w->Li; double *Lx = w->Lx, *D = w->Dg, *Dinv = w->Dinv;
    const int *Lnz = w->Lnz, *etree = w->etree;
    unsigned char *marked = w->bwork; int *ybuf = w->iwork, *ebuf = ybuf + n, *nextcol = ebuf + n;
    double *yv = w->fwork;
    Lp[0] = 0;
    for (int i = 0; i < n; i++) { Lp[i + 1] = Lp[i] + Lnz[i]; marked[i] = 0; yv[i] = 0; nextcol[i] = Lp[i]; D[i] = 0; }
    for (int k = 0; k < n; k++) {
        int nny = 0;
        for (int p = Ap[k]; p < Ap[k + 1]; p++) {
            int b = Ai[p];
            if (b == k) { D[k] = Ax[p]; continue; }
            yv[b] = Ax[p];
            int nb = b;
            if (!marked[nb]) {
                marked[nb] = 1; ebuf[0] = nb; int ne = 1; nb = etree[b];
                while (nb != -1 && nb < k) { if (marked[nb]) break; marked[nb] = 1; ebuf[ne++] = nb; nb = etree[nb]; }
                while (ne) ybuf[nny++] = ebuf[--ne];
            }
        }
        for (int i = nny - 1; i >= 0; i--) {
            int c = ybuf[i]; int tmp = nextcol[c]; double yc = yv[c];
            for (int j = Lp[c]; j < tmp; j++) yv[Li[j]] -= Lx[j] * yc;
            Li[tmp] = k; Lx[tmp] = yc * Dinv[c]; D[k] -= yc * Lx[tmp]; nextcol[c]++;
            yv[c] = 0; marked[c] = 0;
        }
        if (D[k] == 0.0) return -1;
        Dinv[k] = 1.0 / D[k];
    }
    return 0;
}
static void ldl_solve(const op_work *w, double *x) {
    const int n = w->N; const int *Lp = w->Lp, *Li = w->Li; const double *Lx = w->Lx, *Dinv = w->Dinv;
    for (int i = 0; i < n; i++) { double xi = x[i]; for (int j = Lp[i]; j < Lp[i + 1]; j++) x[Li[j]] -= Lx[j] * xi; }
    for (int i = 0; i < n; i++) x[i] *= Dinv[i];
    for (int i = n - 1; i >= 0; i--) { double xi = x[i]; for (int j = Lp[i]; j < Lp[i + 1]; j++) xi -= Lx[j] * x[Li[j]]; x[i] = xi; }
}

/* ---------- rho handling ---------- */
static void set_rho_vec(op_work *w) {
    for (int i = 0; i < w->m; i++) {
        double r;
        if (w->l[i] < -OP_INFTY * OP_MIN_SCALING && w->u[i] > OP_INFTY * OP_MIN_SCALING) r = OP_RHO_MIN;
        else if (w->u[i] - w->l[i] < OP_RHO_TOL) r = OP_RHO_EQ_FACTOR * w->s.rho;
        else r = w->s.rho;
        w->rho_vec[i] = r; w->rho_inv[i] = 1.0 / r; w->Kx[w->rho_idx[i]] = -1.0 / r;
    }
}

/* ---------- public API ---------- */
op_work *osqp_port_setup(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                         const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                         const double *D, const double *E, double c,
                         const int *Kp, const int *Ki, const double *Kx, const int *perm, const int *rho_idx,
                         const op_settings *s) {
    op_work *w = (op_work *)calloc(1, sizeof(op_work));
    w->n = n; w->m = m; w->N = n + m; w->s = *s; w->c = c;
#define DUP(dst, src, cnt, T) do { dst = (T *)malloc(sizeof(T) * (cnt)); memcpy(dst, src, sizeof(T) * (cnt)); } while (0)
    DUP(w->Pp, Pp, n + 1, int); DUP(w->Pi, Pi, Pp[n], int); DUP(w->Px, Px, Pp[n], double);
    DUP(w->Ap, Ap, n + 1, int); DUP(w->Ai, Ai, Ap[n], int); DUP(w->Ax, Ax, Ap[n], double);
    DUP(w->q, q, n, double); DUP(w->l, l, m, double); DUP(w->u, u, m, double);
    DUP(w->D, D, n, double); DUP(w->E, E, m, double);
    int N = w->N;
    DUP(w->Kp, Kp, N + 1, int); DUP(w->Ki, Ki, Kp[N], int); DUP(w->Kx, Kx, Kp[N], double);
    DUP(w->perm, perm, N, int); DUP(w->rho_idx, rho_idx, m, int);
#define ALD(cnt) (double *)calloc((cnt), sizeof(double))
    w->x = ALD(n); w->z = ALD(m); w->y = ALD(m); w->xprev = ALD(n); w->zprev = ALD(m);
    w->sol = ALD(N); w->rhs = ALD(N); w->dx = ALD(n); w->dy = ALD(m);
    w->Axv = ALD(m); w->Pxv = ALD(n); w->Atyv = ALD(n); w->tmpn = ALD(n); w->tmpm = ALD(m);
    w->rho_vec = ALD(m); w->rho_inv = ALD(m); w->xout = ALD(n); w->yout = ALD(m);
    w->etree = (int *)malloc(sizeof(int) * N); w->Lnz = (int *)malloc(sizeof(int) * N);
    w->iwork = (int *)malloc(sizeof(int) * 3 * N); w->bwork = (unsigned char *)malloc(N); w->fwork = ALD(N);
    w->Lp = (int *)malloc(sizeof(int) * (N + 1)); w->Dg = ALD(N); w->Dinv = ALD(N);
    set_rho_vec(w);
    int nnzL = ldl_etree(N, w->Kp, w->Ki, w->iwork, w->Lnz, w->etree);
    if (nnzL < 0) { free(w); return NULL; }
    w->Li = (int *)malloc(sizeof(int) * (nnzL + 1)); w->Lx = ALD(nnzL + 1);
    if (ldl_factor(w) != 0) { free(w); return NULL; }
    w->status = OP_UNSOLVED;
    return w;
}

int osqp_port_nnzL(const op_work *w) { return w->Lp[w->N]; }

/* vectors are UNSCALED on input (like OSQP.update); NULL = keep */
void osqp_port_update(op_work *w, const double *q, const double *l, const double *u) {
    if (q) for (int i = 0; i < w->n; i++) w->q[i] = w->c * w->D[i] * q[i];
    if (l) for (int i = 0; i < w->m; i++) { double v = l[i] < -OP_INFTY ? -OP_INFTY : l[i]; w->l[i] = w->E[i] * v; }
    if (u) for (int i = 0; i < w->m; i++) { double v = u[i] > OP_INFTY ? OP_INFTY : u[i]; w->u[i] = w->E[i] * v; }
    if (l || u) { /* equality pattern could change: refresh rho (OSQP does the same) */
        int changed = 0;
        for (int i = 0; i < w->m; i++) {
            double r;
            if (w->l[i] < -OP_INFTY * OP_MIN_SCALING && w->u[i] > OP_INFTY * OP_MIN_SCALING) r = OP_RHO_MIN;
            else if (w->u[i] - w->l[i] < OP_RHO_TOL) r = OP_RHO_EQ_FACTOR * w->s.rho; else r = w->s.rho;
            if (r != w->rho_vec[i]) changed = 1;
        }
        if (changed) { set_rho_vec(w); ldl_factor(w); }
    }
}

void osqp_port_warm_start(op_work *w, const double *x, const double *y) {
    if (x) { for (int i = 0; i < w->n; i++) w->x[i] = x[i] / w->D[i]; csc_mv(w->n, w->Ap, w->Ai, w->Ax, w->x, w->z, w->m); }
    if (y) for (int i = 0; i < w->m; i++) w->y[i] = w->c * y[i] / w->E[i];
}

static void compute_residual_terms(op_work *w) {
    csc_mv(w->n, w->Ap, w->Ai, w->Ax, w->x, w->Axv, w->m);
    sym_upper_mv(w->n, w->Pp, w->Pi, w->Px, w->x, w->Pxv);
    csc_tmv(w->n, w->Ap, w->Ai, w->Ax, w->y, w->Atyv);
}

static int check_termination(op_work *w, int approx_unused) {
    (void)approx_unused;
    const int n = w->n, m = w->m; const op_settings *s = &w->s;
    compute_residual_terms(w);
    for (int i = 0; i < m; i++) w->tmpm[i] = w->Axv[i] - w->z[i];
    double pri = inf_norm_scaled(w->E, w->tmpm, m, 1);
    for (int i = 0; i < n; i++) w->tmpn[i] = w->Pxv[i] + w->q[i] + w->Atyv[i];
    double dua = inf_norm_scaled(w->D, w->tmpn, n, 1) / w->c;
    w->pri_res = pri; w->dua_res = dua;
    double np_ = fmax(inf_norm_scaled(w->E, w->Axv, m, 1), inf_norm_scaled(w->E, w->z, m, 1));
    double nd_ = fmax(fmax(inf_norm_scaled(w->D, w->Pxv, n, 1), inf_norm_scaled(w->D, w->Atyv, n, 1)),
                      inf_norm_scaled(w->D, w->q, n, 1)) / w->c;
    double eps_p = s->eps_abs + s->eps_rel * np_, eps_d = s->eps_abs + s->eps_rel * nd_;
    if (pri <= eps_p && dua <= eps_d) return OP_SOLVED;
    /* primal infeasibility certificate on delta_y (OSQP paper §3.4): project delta_y on the polar of the
       recession cone of [l,u], then test  A' dy ~ 0  and  u'dy+ + l'dy- < 0 */
    for (int i = 0; i < m; i++) {
        double d = w->dy[i];
        if (w->u[i] >= OP_INFTY * OP_MIN_SCALING && d > 0) d = 0;
        if (w->l[i] <= -OP_INFTY * OP_MIN_SCALING && d < 0) d = 0;
        w->tmpm[i] = d;
    }
    double ndy = inf_norm_scaled(w->E, w->tmpm, m, 0);
    if (ndy > 1e-30) {
        double supp = 0;
        for (int i = 0; i < m; i++) { double d = w->tmpm[i]; supp += d > 0 ? w->u[i] * d : (d < 0 ? w->l[i] * d : 0.0); }
        if (supp < -s->eps_prim_inf * ndy) {
            csc_tmv(n, w->Ap, w->Ai, w->Ax, w->tmpm, w->tmpn);
            if (inf_norm_scaled(w->D, w->tmpn, n, 1) < s->eps_prim_inf * ndy) return OP_PRIMAL_INFEASIBLE;
        }
    }
    /* dual infeasibility certificate on delta_x */
    double ndx = inf_norm_scaled(w->D, w->dx, n, 0);
    if (ndx > 1e-30) {
        double qdx = 0; for (int i = 0; i < n; i++) qdx += w->q[i] * w->dx[i];
        if (qdx / w->c < -s->eps_dual_inf * ndx) {
            sym_upper_mv(n, w->Pp, w->Pi, w->Px, w->dx, w->tmpn);
            if (inf_norm_scaled(w->D, w->tmpn, n, 1) / w->c < s->eps_dual_inf * ndx) {
                csc_mv(n, w->Ap, w->Ai, w->Ax, w->dx, w->tmpm, m);
                int ok = 1;
                for (int i = 0; i < m && ok; i++) {
                    double v = w->tmpm[i] / w->E[i], t = s->eps_dual_inf * ndx;
                    if (w->u[i] < OP_INFTY * OP_MIN_SCALING && v > t) ok = 0;
                    if (w->l[i] > -OP_INFTY * OP_MIN_SCALING && v < -t) ok = 0;
                }
                if (ok) return OP_DUAL_INFEASIBLE;
            }
        }
    }
    return 0;
}

static int adapt_rho(op_work *w) {
    const int n = w->n, m = w->m;
    compute_residual_terms(w);
    for (int i = 0; i < m; i++) w->tmpm[i] = w->Axv[i] - w->z[i];
    for (int i = 0; i < n; i++) w->tmpn[i] = w->Pxv[i] + w->q[i] + w->Atyv[i];
    double pri = inf_norm(w->tmpm, m), dua = inf_norm(w->tmpn, n);
    double np_ = fmax(inf_norm(w->Axv, m), inf_norm(w->z, m));
    double nd_ = fmax(fmax(inf_norm(w->Pxv, n), inf_norm(w->Atyv, n)), inf_norm(w->q, n));
    pri /= (np_ + 1e-10); dua /= (nd_ + 1e-10);
    double rn = w->s.rho * sqrt(pri / (dua + 1e-10));
    rn = fmin(fmax(rn, OP_RHO_MIN), OP_RHO_MAX);
    if (rn > w->s.rho * w->s.adaptive_rho_tolerance || rn < w->s.rho / w->s.adaptive_rho_tolerance) {
        w->s.rho = rn; set_rho_vec(w); ldl_factor(w); w->rho_updates++; return 1;
    }
    return 0;
}

int osqp_port_solve(op_work *w) {
    const int n = w->n, m = w->m, N = w->N; const op_settings *s = &w->s;
    if (!s->warm_start) { memset(w->x, 0, sizeof(double) * n); memset(w->z, 0, sizeof(double) * m); memset(w->y, 0, sizeof(double) * m); }
    int status = OP_MAX_ITER, it;
    for (it = 1; it <= s->max_iter; it++) {
        memcpy(w->xprev, w->x, sizeof(double) * n); memcpy(w->zprev, w->z, sizeof(double) * m);
        /* rhs = [sigma x - q ; z - y/rho], permuted */
        for (int i = 0; i < n; i++) w->rhs[i] = s->sigma * w->xprev[i] - w->q[i];
        for (int i = 0; i < m; i++) w->rhs[n + i] = w->zprev[i] - w->rho_inv[i] * w->y[i];
        for (int i = 0; i < N; i++) w->sol[i] = w->rhs[w->perm[i]];
        ldl_solve(w, w->sol);
        for (int i = 0; i < N; i++) w->rhs[w->perm[i]] = w->sol[i];   /* rhs now = [xtilde ; nu] */
        for (int i = 0; i < n; i++) {
            double xt = w->rhs[i];
            w->x[i] = s->alpha * xt + (1 - s->alpha) * w->xprev[i];
            w->dx[i] = w->x[i] - w->xprev[i];
        }
        for (int i = 0; i < m; i++) {
            double zt = w->zprev[i] + w->rho_inv[i] * (w->rhs[n + i] - w->y[i]);
            double zr = s->alpha * zt + (1 - s->alpha) * w->zprev[i];
            double zn = zr + w->rho_inv[i] * w->y[i];
            zn = fmin(fmax(zn, w->l[i]), w->u[i]);
            double dy = w->rho_vec[i] * (zr - zn);
            w->z[i] = zn; w->dy[i] = dy; w->y[i] += dy;
        }
        if (s->check_termination && it % s->check_termination == 0) {
            int st = check_termination(w, 0);
            if (st) { status = st; break; }
        }
        if (s->adaptive_rho && s->adaptive_rho_interval && it % s->adaptive_rho_interval == 0) adapt_rho(w);
    }
    if (it > s->max_iter) { it = s->max_iter; int st = check_termination(w, 0); if (st) status = st; }
    w->iter = it; w->status = status;
    for (int i = 0; i < n; i++) w->xout[i] = w->D[i] * w->x[i];
    for (int i = 0; i < m; i++) w->yout[i] = w->E[i] * w->y[i] / w->c;
    sym_upper_mv(n, w->Pp, w->Pi, w->Px, w->x, w->Pxv);
    double obj = 0; for (int i = 0; i < n; i++) obj += w->x[i] * (0.5 * w->Pxv[i] + w->q[i]);
    w->obj_val = obj / w->c;
    return status;
}

const double *osqp_port_x(const op_work *w) { return w->xout; }
const double *osqp_port_y(const op_work *w) { return w->yout; }
void osqp_port_info(const op_work *w, int *iter, int *status, int *rho_updates, double *obj, double *pri, double *dua, double *rho) {
    *iter = w->iter; *status = w->status; *rho_updates = w->rho_updates; *obj = w->obj_val; *pri = w->pri_res; *dua = w->dua_res; *rho = w->s.rho;
}

op_work *osqp_port_clone(const op_work *src) {
    /* deep copy so that every MPC instance owns an independent solver (one controller object per instance) */
    op_work *w = (op_work *)malloc(sizeof(op_work)); *w = *src;
    int n = src->n, m = src->m, N = src->N, nnzL = src->Lp[N];
#define CL(f, cnt, T) do { w->f = (T *)malloc(sizeof(T) * (cnt)); memcpy(w->f, src->f, sizeof(T) * (cnt)); } while (0)
    CL(Pp, n + 1, int); CL(Pi, src->Pp[n], int); CL(Px, src->Pp[n], double);
    CL(Ap, n + 1, int); CL(Ai, src->Ap[n], int); CL(Ax, src->Ap[n], double);
    CL(q, n, double); CL(l, m, double); CL(u, m, double); CL(D, n, double); CL(E, m, double);
    CL(Kp, N + 1, int); CL(Ki, src->Kp[N], int); CL(Kx, src->Kp[N], double); CL(perm, N, int); CL(rho_idx, m, int);
    CL(etree, N, int); CL(Lnz, N, int); CL(Lp, N + 1, int); CL(Li, nnzL + 1, int); CL(Lx, nnzL + 1, double);
    CL(Dg, N, double); CL(Dinv, N, double); CL(iwork, 3 * N, int); CL(bwork, N, unsigned char); CL(fwork, N, double);
    CL(x, n, double); CL(z, m, double); CL(y, m, double); CL(xprev, n, double); CL(zprev, m, double);
    CL(sol, N, double); CL(rhs, N, double); CL(dx, n, double); CL(dy, m, double);
    CL(Axv, m, double); CL(Pxv, n, double); CL(Atyv, n, double); CL(tmpn, n, double); CL(tmpm, m, double);
    CL(rho_vec, m, double); CL(rho_inv, m, double); CL(xout, n, double); CL(yout, m, double);
    return w;
}

void osqp_port_free(op_work *w) {
    if (!w) return;
    void *ptrs[] = {w->Pp, w->Pi, w->Px, w->Ap, w->Ai, w->Ax, w->q, w->l, w->u, w->D, w->E, w->Kp, w->Ki, w->Kx, w->perm,
                    w->rho_idx, w->etree, w->Lnz, w->Lp, w->Li, w->Lx, w->Dg, w->Dinv, w->iwork, w->bwork, w->fwork, w->x, w->z,
                    w->y, w->xprev, w->zprev, w->sol, w->rhs, w->dx, w->dy, w->Axv, w->Pxv, w->Atyv, w->tmpn, w->tmpm,
                    w->rho_vec, w->rho_inv, w->xout, w->yout};
    for (unsigned i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
    free(w);
}

/*
 * Batched MPC step on the host cores: what a user of the reference does for B controllers, minus
 * the Python overhead — per instance rewrite (q, l, u) exactly as _update_QP_matrices_
 * (/root/reference/pyMPC/mpc.py:404-452) does for a constant xref, OSQP.update, OSQP.solve, slice
 * u_0 (mpc.py:302).  qx_coef is (NX x nx) with q_X = qx_coef @ xref; qdu is (nu x nu) = -QDu;
 * q_base holds the u_ref part of q.  Threads: OpenMP over instances (one solver object each).
 */
int osqp_port_mpc_step_batch(op_work **ws, int B, int nx, int nu, int NX, int NU, int nrow_du0,
                             const double *qx_coef, const double *qdu, const double *q_base,
                             const double *l_base, const double *u_base,
                             const double *X0, const double *Um1, const double *Xref,
                             double *U0, int *status, int *iters, int nthreads) {
    int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : bad)
#endif
    for (int b = 0; b < B; b++) {
        op_work *w = ws[b]; int n = w->n, m = w->m;
        double *q = (double *)malloc(sizeof(double) * (n + 2 * m)), *l = q + n, *u = l + m;
        memcpy(q, q_base, sizeof(double) * n); memcpy(l, l_base, sizeof(double) * m); memcpy(u, u_base, sizeof(double) * m);
        const double *x0 = X0 + (size_t)b * nx, *um1 = Um1 + (size_t)b * nu, *xr = Xref + (size_t)b * nx;
        for (int i = 0; i < NX; i++) { double s = 0; for (int j = 0; j < nx; j++) s += qx_coef[i * nx + j] * xr[j]; q[i] += s; }
        for (int i = 0; i < nu; i++) { double s = 0; for (int j = 0; j < nu; j++) s += qdu[i * nu + j] * um1[j]; q[NX + i] += s; }
        for (int i = 0; i < nx; i++) { l[i] = -x0[i]; u[i] = -x0[i]; }
        for (int i = 0; i < nu; i++) { l[nrow_du0 + i] += um1[i]; u[nrow_du0 + i] += um1[i]; }
        osqp_port_update(w, q, l, u);
        int st = osqp_port_solve(w);
        status[b] = st; iters[b] = w->iter;
        for (int i = 0; i < nu; i++) U0[(size_t)b * nu + i] = w->xout[NX + i];
        if (st != OP_SOLVED) bad++;
        free(q);
    }
    (void)NU;
    return bad;
}
