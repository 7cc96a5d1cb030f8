// bmpc.cu — CUDA kernels (sm_100a) and the C ABI of libbmpc.so (see include/bmpc.h).
//
// Kernel map:
//   k_condense   K1/K2  one CTA per system: prediction matrices, H, K, inverses, dual operators
//   k_admm       K3+K4  one team (warp or CTA) per instance: per-step prep + ADMM iterations, state in smem
//   k_polish     K5+K6  one team per instance: active-set polish, KKT verification, output epilogue
//   k_finalize          status / fallback output for instances the polish never verified
//   k_sequences         optional x/u/eps sequences and objective value (output() info, mpc.py:307-328)
// Host side: handle management and the round loop  ADMM(first_iters) -> polish -> [ADMM(chunk) -> polish]*.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/bmpc.h"
#include "bmpc_core.cuh"
#include "bmpc_tpi.cuh"
#include "bmpc_tile.cuh"
#include "bmpc_tpm.cuh"

// One block of round counters = BMPC_CNT ints (the handle keeps two, used alternately): [0] instances still unfinished (-> next list),
// [1] active-set refinements, [2] overflow of the small polish tier, [3] certified infeasible, [4..7] queue control of k_tpi_pol
// (cursor, head, and one 64-bit word: low = reserved queue tail, high = finished chunks), [8] instances that left as "solved,
// unpolished" on tight ADMM residuals.
// [10..11] / [12..13]: ~(earliest start) and latest end of k_tpi_pol in ns of the GPU's global timer (64-bit each): the kernel's
// own duration, for the roofline of bench.py (CUDA events around a launch also count launch gaps).
// [14]: warps of k_tpi_pol that have left; the last one copies the block into the handle's MAPPED pinned host buffer and raises
// [15] = launch epoch there: the host spins on that word instead of paying a D2H copy + stream synchronisation per solve.
enum { TPI_Q_CURSOR = 4, TPI_Q_HEAD = 5, TPI_Q_TAIL = 6, BMPC_CNT_TIGHT = 8, BMPC_CNT_T0 = 10, BMPC_CNT_T1 = 12, BMPC_CNT_EXIT = 14, BMPC_CNT_EPOCH = 15, BMPC_CNT = 16 };
__device__ __forceinline__ unsigned long long bmpc_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ------------------------------------------------------------------------------------------------
// teams
struct WarpTeam {
    int tid;
    static constexpr int n = 32;
    __device__ __forceinline__ WarpTeam() : tid(threadIdx.x & 31) {}
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ bool all(bool p) { return __all_sync(0xffffffffu, p); }
    __device__ __forceinline__ double max(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
        return v;
    }
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        return v;
    }
    __device__ __forceinline__ int excl_scan(int f, int& total) {
        unsigned b = __ballot_sync(0xffffffffu, f);
        total = __popc(b);
        return __popc(b & ((1u << tid) - 1u));
    }
    __device__ __forceinline__ int warp() const { return 0; }
    __device__ __forceinline__ int nwarps() const { return 1; }
    __device__ __forceinline__ int lane() const { return tid; }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    __device__ __forceinline__ double wsum(double v) { return sum(v); }
};

struct BlockTeam {
    int tid, n;
    double* sd;   // 32 doubles of scratch
    int* si;      // 32 ints of scratch
    __device__ __forceinline__ BlockTeam(double* sd_, int* si_) : tid(threadIdx.x), n(blockDim.x), sd(sd_), si(si_) {}
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool all(bool p) { return __syncthreads_and(p) != 0; }
    __device__ __forceinline__ double max(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
        if ((tid & 31) == 0) sd[tid >> 5] = v;
        __syncthreads();
        double r = sd[0];
        for (int w = 1; w < (n + 31) / 32; w++) r = fmax(r, sd[w]);
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if ((tid & 31) == 0) sd[tid >> 5] = v;
        __syncthreads();
        double r = 0.0;
        for (int w = 0; w < (n + 31) / 32; w++) r += sd[w];
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ int excl_scan(int f, int& total) {
        unsigned b = __ballot_sync(0xffffffffu, f);
        if ((tid & 31) == 0) si[tid >> 5] = __popc(b);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < (n + 31) / 32; w++) { int c = si[w]; if (w < (tid >> 5)) off += c; tot += c; }
        __syncthreads();
        total = tot;
        return off + __popc(b & ((1u << (tid & 31)) - 1u));
    }
    __device__ __forceinline__ int warp() const { return tid >> 5; }
    __device__ __forceinline__ int nwarps() const { return n >> 5; }
    __device__ __forceinline__ int lane() const { return tid & 31; }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    __device__ __forceinline__ double wsum(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// per-instance device arrays
struct BmpcInst {
    const double* x0;     // [B,nx]
    const double* um1;    // [B,nu]
    const double* xref;   // [B,nx] or [B,NX]
    const double* um1_solved;  // [B,nu] copy of um1 as seen by the last solve (output() may overwrite um1)
    double* g;            // [B,NU]   scratch (per-step linear term)
    double* cc;           // [B,NX]   scratch (free response Acal x0)
    double* xw;           // [B,NU]   ADMM x (warm start)
    double* vw;           // [B,mc]   ADMM v (warm start)
    double* Ua;           // [B,NU]   last ADMM xt (fallback solution)
    double* Us;           // [B,NU]   solution
    double* res;          // [B,4]    residuals of the last ADMM iteration
    double* u0;           // [B,nu]   output
    int32_t* status;      // [B]
    int32_t* iters;       // [B]
    int32_t* psteps;      // [B]
    int32_t* lvl;         // [B] adaptive-rho ladder level (reset to the base level at every solve)
    double* u0_shadow;    // second copy of the output = the NEXT solve's u_-1 if output() commits it (a pointer swap instead of a copy)
    double* u0_peer[8];   // extra copies of the output slice in peer GPUs' buffers (NVLink peer stores), see bmpc_bind_output_peers
    int n_peer;
    size_t sys_stride;    // 0: all instances share one system block; o.total: instance i uses block i (per-instance Ad, Bd, ...)
};

// K6: the solver epilogue stores u* into this rank's output slice and, when peers are bound, straight into every peer
// GPU's copy of the gathered buffer over NVLink (the all-gather is fused into the kernel: no collective launch)
__device__ __forceinline__ void bmpc_publish_u0(const BmpcInst& I, double* u0_out, size_t idx, double val) {
    u0_out[idx] = val;
    I.u0_shadow[idx] = val;
    for (int p = 0; p < I.n_peer; p++) I.u0_peer[p][idx] = val;
}

// K6, arrival: after the epilogue's peer stores, every rank raises its flag (= the step counter) in every peer's flag array
// and waits until all peers have raised theirs here: when the kernel ends, the gathered buffer of this step is complete on
// this rank.  One tiny launch instead of a collective; the flag stores are ordered behind the data stores of the solver
// kernels by the stream and a system-scope fence.
struct BmpcPeerFlags { long long* p[8]; };
__global__ void k_gather_arrive(long long* mine, BmpcPeerFlags P, int n_peer, int rank, int world, long long epoch) {
    const int t = threadIdx.x;
    if (t < n_peer) { __threadfence_system(); *(volatile long long*)(P.p[t] + rank) = epoch; }
    if (t < world && t != rank) {
        const long long t0 = clock64();
        while (*(volatile long long*)(mine + t) < epoch) {
            if (clock64() - t0 > 20000000000ll) __trap();            // ~10 s: a peer died; fail loudly instead of hanging the GPU
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// smem (doubles) per instance for the two kernels — keep in sync with the carve-up below
__host__ __device__ static inline size_t admm_smem_doubles(const BmpcDims& d) { return 4 * (size_t)d.NU + d.NX + 2 * (size_t)d.mc + d.nu + 4 + 4; }
__host__ __device__ static inline size_t polish_smem_doubles(const BmpcDims& d, int rmax) {
    return 3 * (size_t)d.NU + d.NX + 4 * (size_t)d.mc + (size_t)rmax * (rmax + 1) / 2 + rmax + d.nu + 2 + (d.mc + rmax + 3) / 2;
}

__global__ void k_condense(BmpcDims d, BmpcSysOff o, double* sys, double rho, double sigma, double alpha, double eps_feas,
                           int soft_on) {
    __shared__ double sd[32];
    __shared__ int si[32];
    BlockTeam t(sd, si);
    bmpc_condense(t, d, o, sys + (size_t)blockIdx.x * o.total, rho, sigma, alpha, eps_feas, soft_on);
}

template <bool WARP>
__global__ void k_admm(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, const int32_t* __restrict__ list,
                       int count, int niter, int do_prep, int cold, int xref_mode, int adapt) {
    extern __shared__ double smem[];
    __shared__ double sd[32];
    __shared__ int si[32];
    int slot, idx;
    if (WARP) { slot = threadIdx.x >> 5; idx = blockIdx.x * (blockDim.x >> 5) + slot; }
    else { slot = 0; idx = blockIdx.x; }
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    sys += (size_t)inst * I.sys_stride;
    double* base = smem + (size_t)slot * admm_smem_doubles(d);
    double *g = base, *cc = g + d.NU, *x = cc + d.NX, *v = x + d.NU, *w = v + d.mc, *xt = w + d.mc, *r = xt + d.NU,
           *um1 = r + d.NU, *res = um1 + d.nu + (d.nu & 1);
    auto run = [&](auto& t) {
        for (int q = t.tid; q < d.nu; q += t.n) um1[q] = I.um1[(size_t)inst * d.nu + q];
        t.sync();
        if (do_prep) {
            const int xl = xref_mode ? d.NX : d.nx;
            bmpc_prep(t, d, o, sys, I.x0 + (size_t)inst * d.nx, um1, I.xref + (size_t)inst * xl, xref_mode, g, cc);
            for (int a = t.tid; a < d.NU; a += t.n) I.g[(size_t)inst * d.NU + a] = g[a];
            for (int i = t.tid; i < d.NX; i += t.n) I.cc[(size_t)inst * d.NX + i] = cc[i];
        } else {
            for (int a = t.tid; a < d.NU; a += t.n) g[a] = I.g[(size_t)inst * d.NU + a];
            for (int i = t.tid; i < d.NX; i += t.n) cc[i] = I.cc[(size_t)inst * d.NX + i];
        }
        if (cold) {
            for (int a = t.tid; a < d.NU; a += t.n) x[a] = 0.0;
            t.sync();
            for (int i = t.tid; i < d.mc; i += t.n) v[i] = (i < d.NX) ? cc[i] : 0.0;
        } else {
            for (int a = t.tid; a < d.NU; a += t.n) x[a] = I.xw[(size_t)inst * d.NU + a];
            for (int i = t.tid; i < d.mc; i += t.n) v[i] = I.vw[(size_t)inst * d.mc + i];
        }
        t.sync();
        int lvl = I.lvl[inst];
        bmpc_admm(t, d, o, sys, um1, g, cc, x, v, w, xt, r, niter, res, lvl);
        if (adapt) {
            const int nl = bmpc_adapt_level(t, d, o, sys, um1, v, res, lvl);
            if (t.tid == 0 && nl != lvl) I.lvl[inst] = nl;
        }
        for (int a = t.tid; a < d.NU; a += t.n) { I.xw[(size_t)inst * d.NU + a] = x[a]; I.Ua[(size_t)inst * d.NU + a] = xt[a]; }
        for (int i = t.tid; i < d.mc; i += t.n) I.vw[(size_t)inst * d.mc + i] = v[i];
        if (t.tid < 4) I.res[(size_t)inst * 4 + t.tid] = res[t.tid];
        if (t.tid == 0) I.iters[inst] += niter;
    };
    if (WARP) { WarpTeam t; run(t); }
    else { BlockTeam t(sd, si); run(t); }
}


// K3+K4 for large shapes: one CTA per TILE of T instances (bmpc_tile.cuh).  Same contract as k_admm.
template <int T, int NS, int NXC, int NUC>
__global__ void __launch_bounds__(NS >= 4 ? 768 : 512) k_admm_tile(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, const int32_t* __restrict__ list,
                            int count, int niter, int do_prep, int cold, int xref_mode, int adapt) {
    extern __shared__ double smem[];
    __shared__ double sd[32];
    __shared__ int si[32];
    BlockTeam t(sd, si);
    BmpcTile<T> S; S.carve(smem, d);
    const int first = blockIdx.x * T;
    const int nact = (count - first) < T ? (count - first) : T;       // a partial tile repeats its last instance (no write-back)
    if (t.tid < T) {
        const int e = t.tid < nact ? t.tid : nact - 1;
        const int inst = list ? list[first + e] : first + e;
        S.inst[t.tid] = inst; S.lvl[t.tid] = I.lvl[inst];
    }
    t.sync();
    bmpc_tile_load_phi(t, d, sys + o.Bcal, S.phi1, S.phi2);
    bmpc_tile_load_rows(t, d, o, sys, S.lo, S.hi, S.rho);
    for (int idx = t.tid; idx < d.nu * T; idx += t.n) { const int e = idx / d.nu, q = idx % d.nu; S.um1[idx] = I.um1[(size_t)S.inst[e] * d.nu + q]; }
    if (do_prep) {
        double* x0 = S.r;                                              // staging: r is free until the first iteration
        for (int idx = t.tid; idx < d.nx * T; idx += t.n) { const int e = idx / d.nx, c = idx % d.nx; x0[idx] = I.x0[(size_t)S.inst[e] * d.nx + c]; }
        t.sync();
        bmpc_tile_prep(t, d, o, sys, S, x0, I.xref, xref_mode);
        for (int idx = t.tid; idx < d.NU * nact; idx += t.n) { const int e = idx / d.NU, a = idx % d.NU; I.g[(size_t)S.inst[e] * d.NU + a] = S.g[(size_t)a * T + e]; }
        for (int idx = t.tid; idx < d.NX * nact; idx += t.n) { const int e = idx / d.NX, i = idx % d.NX; I.cc[(size_t)S.inst[e] * d.NX + i] = S.cc[(size_t)i * T + e]; }
    } else {
        for (int idx = t.tid; idx < d.NU * T; idx += t.n) { const int e = idx / d.NU, a = idx % d.NU; S.g[(size_t)a * T + e] = I.g[(size_t)S.inst[e] * d.NU + a]; }
        for (int idx = t.tid; idx < d.NX * T; idx += t.n) { const int e = idx / d.NX, i = idx % d.NX; S.cc[(size_t)i * T + e] = I.cc[(size_t)S.inst[e] * d.NX + i]; }
    }
    t.sync();
    if (cold) {
        for (int idx = t.tid; idx < d.NU * T; idx += t.n) S.x[idx] = 0.0;
        for (int idx = t.tid; idx < d.mc * T; idx += t.n) S.v[idx] = (idx / T < d.NX) ? S.cc[idx] : 0.0;
    } else {
        for (int idx = t.tid; idx < d.NU * T; idx += t.n) { const int e = idx / d.NU, a = idx % d.NU; S.x[(size_t)a * T + e] = I.xw[(size_t)S.inst[e] * d.NU + a]; }
        for (int idx = t.tid; idx < d.mc * T; idx += t.n) { const int e = idx / d.mc, i = idx % d.mc; S.v[(size_t)i * T + e] = I.vw[(size_t)S.inst[e] * d.mc + i]; }
    }
    t.sync();
    bmpc_admm_tile<T, NS, NXC, NUC>(t, d, o, sys, S, niter);
    if (adapt) {
        bmpc_tile_adapt(t, d, o, sys, S);
        if (t.tid < nact && S.nlvl[t.tid] != S.lvl[t.tid]) I.lvl[S.inst[t.tid]] = S.nlvl[t.tid];
    }
    for (int idx = t.tid; idx < d.NU * nact; idx += t.n) {
        const int e = idx / d.NU, a = idx % d.NU; const size_t gidx = (size_t)S.inst[e] * d.NU + a;
        I.xw[gidx] = S.x[(size_t)a * T + e]; I.Ua[gidx] = S.xt[(size_t)a * T + e];
    }
    for (int idx = t.tid; idx < d.mc * nact; idx += t.n) { const int e = idx / d.mc, i = idx % d.mc; I.vw[(size_t)S.inst[e] * d.mc + i] = S.v[(size_t)i * T + e]; }
    for (int idx = t.tid; idx < 4 * nact; idx += t.n) I.res[(size_t)S.inst[idx / 4] * 4 + idx % 4] = S.res[idx];
    if (t.tid < nact) I.iters[S.inst[t.tid]] += niter;
}

template <bool WARP>
__global__ void k_polish(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, const int32_t* __restrict__ list,
                         int count, int rmax, int max_steps, int32_t* next_list, int32_t* next_count, double* u0_out,
                         const int32_t* dev_count, int32_t* ovf_list, int cand_warm) {
    extern __shared__ double smem[];
    __shared__ double sd[32];
    __shared__ int si[32];
    int slot, idx;
    if (WARP) { slot = threadIdx.x >> 5; idx = blockIdx.x * (blockDim.x >> 5) + slot; }
    else { slot = 0; idx = blockIdx.x; }
    if (dev_count) count = *dev_count;                 // second tier: the list was filled by the launch before this one
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    sys += (size_t)inst * I.sys_stride;
    double* base = smem + (size_t)slot * polish_smem_doubles(d, rmax);
    double *g = base, *cc = g + d.NU, *v = cc + d.NX, *W0 = v + d.mc, *zz = W0 + d.mc, *murow = zz + d.mc,
           *S = murow + d.mc, *tt = S + (size_t)rmax * (rmax + 1) / 2, *U0 = tt + rmax, *U = U0 + d.NU, *um1 = U + d.NU;
    int* st = (int*)(um1 + d.nu + (d.nu & 1));
    int* R = st + d.mc;
    auto run = [&](auto& t) {
        for (int q = t.tid; q < d.nu; q += t.n) um1[q] = I.um1[(size_t)inst * d.nu + q];
        for (int a = t.tid; a < d.NU; a += t.n) g[a] = I.g[(size_t)inst * d.NU + a];
        for (int i = t.tid; i < d.NX; i += t.n) cc[i] = I.cc[(size_t)inst * d.NX + i];
        for (int i = t.tid; i < d.mc; i += t.n) v[i] = I.vw[(size_t)inst * d.mc + i];
        t.sync();
        int ps = bmpc_polish(t, d, o, sys, um1, g, cc, v, W0, zz, murow, st, S, tt, R, U0, U, rmax, max_steps);
        if (ps > 0) {
            const double* rhov = sys + o.rho;
            for (int a = t.tid; a < d.NU; a += t.n) {
                double ua = U[a];
                I.Us[(size_t)inst * d.NU + a] = ua; I.xw[(size_t)inst * d.NU + a] = ua;
                if (a < d.nu) bmpc_publish_u0(I, u0_out, (size_t)inst * d.nu + a, ua);
            }
            // exact ADMM fixed point of this problem = warm start of the next one: v* = z* + y*/rho
            for (int i = t.tid; i < d.mc; i += t.n) I.vw[(size_t)inst * d.mc + i] = zz[i] + murow[i] / rhov[i];
            if (t.tid == 0) { I.status[inst] = BMPC_SOLVED; I.psteps[inst] += ps; atomicAdd(next_count + 1, ps); }
        } else if (ps == -1 && ovf_list) {
            // working set larger than this tier's capacity: hand the instance to the large-capacity launch that follows
            if (t.tid == 0) { int pos = atomicAdd(next_count + 2, 1); ovf_list[pos] = inst; }
        } else {
            // not verified.  (a) ADMM residuals far below any tolerance: the iterate is the answer to ~1e-8 although the active-set
            // iteration cannot certify it (degenerate vertex): "solved" like OSQP would say, no more rounds.  (b) otherwise, if the
            // last candidate is sane (hard rows feasible to 1e-2), its rows and multipliers become the ADMM state of the next
            // round: a strongly violated soft row carries the multiplier eps_feas * d, which ADMM alone builds in eps_feas d / rho steps
            const double* resg = I.res + (size_t)inst * 4;
            const int itsg = I.iters[inst];
            const bool tight = bmpc_residuals_tight(resg, itsg);
            if (tight) {
                for (int a = t.tid; a < d.NU; a += t.n) {
                    const double ua = I.Ua[(size_t)inst * d.NU + a];
                    I.Us[(size_t)inst * d.NU + a] = ua;
                    if (a < d.nu) bmpc_publish_u0(I, u0_out, (size_t)inst * d.nu + a, ua);
                }
            } else if (ps == 0 && cand_warm && bmpc_admm_stalled(resg, itsg) && bmpc_candidate_usable(t, d, o, sys, um1, zz, murow)) {
                const int lvl = I.lvl[inst];
                bmpc_warm_from_candidate(t, d, o, sys, zz, murow, U, I.xw + (size_t)inst * d.NU, I.vw + (size_t)inst * d.mc, lvl);
            }
            if (t.tid == 0) {
                int used = (ps < 0 ? 1 : max_steps);
                I.psteps[inst] += used; atomicAdd(next_count + 1, used);
                if (tight) { I.status[inst] = BMPC_SOLVED_UNPOLISHED; atomicAdd(next_count + BMPC_CNT_TIGHT, 1); }
                else if (I.status[inst] != BMPC_PRIMAL_INFEASIBLE) { int pos = atomicAdd(next_count, 1); next_list[pos] = inst; }   // a certified instance is finished
            }
        }
    };
    if (WARP) { WarpTeam t; run(t); }
    else { BlockTeam t(sd, si); run(t); }
}

// instances still unverified after the last round: OSQP's criterion decides between "solved" and "max iter";
// the output falls back to u_failure = uref when not solved (mpc.py:230,303-304).  pure_admm: no polish was run.
__global__ void k_finalize(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, const int32_t* __restrict__ list,
                           int count, double eps_abs, double eps_rel, double* u0_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    sys += (size_t)inst * I.sys_stride;
    const double* res = I.res + (size_t)inst * 4;
    bool conv = res[0] <= eps_abs + eps_rel * res[2] && res[1] <= eps_abs + eps_rel * res[3];
    if (I.status[inst] == BMPC_PRIMAL_INFEASIBLE) return;               // certified: status and u_failure already written
    I.status[inst] = conv ? BMPC_SOLVED_UNPOLISHED : BMPC_MAX_ITER;
    for (int a = 0; a < d.NU; a++) I.Us[(size_t)inst * d.NU + a] = I.Ua[(size_t)inst * d.NU + a];
    for (int q = 0; q < d.nu; q++) bmpc_publish_u0(I, u0_out, (size_t)inst * d.nu + q, conv ? I.Ua[(size_t)inst * d.NU + q] : sys[o.uref + q]);
}


// ---- primal infeasibility (OSQP paper 3.4), checked on straggler rounds only: the change of the multipliers y over one
// ADMM round, projected like OSQP does, is a certificate when A' dy ~ 0 and the support function of the row box is negative.
// Soft rows are penalties, not constraints: they stay out.  k_snapshot keeps (v, level) of the listed instances before the
// round; k_infeas compares with the state after it.
__global__ void k_snapshot(BmpcDims d, BmpcInst I, const int32_t* __restrict__ list, int count, double* vprev, int32_t* lprev) {
    const int idx = blockIdx.x;
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    for (int i = threadIdx.x; i < d.mc; i += blockDim.x) vprev[(size_t)inst * d.mc + i] = I.vw[(size_t)inst * d.mc + i];
    if (threadIdx.x == 0) lprev[inst] = I.lvl[inst];
}

__global__ void k_infeas(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, const int32_t* __restrict__ list, int count,
                         const double* __restrict__ vprev, const int32_t* __restrict__ lprev, double eps_pinf, double* u0_out, int32_t* counts) {
    extern __shared__ double smem[];
    __shared__ double sd[32];
    __shared__ int si[32];
    const int idx = blockIdx.x;
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    sys += (size_t)inst * I.sys_stride;
    BlockTeam t(sd, si);
    double* dy = smem; double* um1 = dy + d.mc;
    for (int q = t.tid; q < d.nu; q += t.n) um1[q] = I.um1[(size_t)inst * d.nu + q];
    t.sync();
    if (bmpc_primal_infeasible(t, d, o, sys, um1, I.cc + (size_t)inst * d.NX, vprev + (size_t)inst * d.mc, lprev[inst],
                               I.vw + (size_t)inst * d.mc, I.lvl[inst], dy, eps_pinf)) {
        if (t.tid == 0) { I.status[inst] = BMPC_PRIMAL_INFEASIBLE; atomicAdd(counts + 3, 1); }
        for (int q = t.tid; q < d.nu; q += t.n) bmpc_publish_u0(I, u0_out, (size_t)inst * d.nu + q, sys[o.uref + q]);
    }
}

// pure-ADMM mode: list of instances not yet converged by OSQP's criterion
__global__ void k_check_converged(BmpcInst I, const int32_t* __restrict__ list, int count, double eps_abs, double eps_rel,
                                  int32_t* next_list, int32_t* next_count) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const int inst = list ? list[idx] : idx;
    const double* res = I.res + (size_t)inst * 4;
    bool conv = res[0] <= eps_abs + eps_rel * res[2] && res[1] <= eps_abs + eps_rel * res[3];
    if (!conv && I.status[inst] != BMPC_PRIMAL_INFEASIBLE) { int pos = atomicAdd(next_count, 1); next_list[pos] = inst; }
}

__global__ void k_reset(BmpcInst I, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { I.status[i] = BMPC_UNSOLVED; I.iters[i] = 0; I.psteps[i] = 0; I.lvl[i] = BMPC_LEV0; }
}

// x_seq = Acal x0 + Bcal U, eps_seq = distance back to the box, objective of the reference QP (without J_CNST).
// Recomputes g and cc from (x0, u_-1, xref) so it does not depend on which solve path ran.
__global__ void k_sequences(BmpcDims d, BmpcSysOff o, const double* __restrict__ sys, BmpcInst I, int B, int xref_mode,
                            double* xseq, double* epsseq, double* obj) {
    extern __shared__ double smem[];
    WarpTeam t;
    const int slot = threadIdx.x >> 5;
    int inst = blockIdx.x * (blockDim.x >> 5) + slot;
    if (inst >= B) return;
    sys += (size_t)inst * I.sys_stride;
    double* g = smem + (size_t)slot * (d.NU + d.NX + d.nu + 1);
    double* cc = g + d.NU;
    double* um1 = cc + d.NX;
    const double *BcalT = sys + o.BcalT, *H = sys + o.H, *lo0 = sys + o.lo0, *hi0 = sys + o.hi0;
    const double *Qx = sys + o.Qx, *QxN = sys + o.QxN;
    const double rho_e = sys[o.scal + BMPC_S_RHOE];
    const double* U = I.Us + (size_t)inst * d.NU;
    const double* xr = I.xref + (size_t)inst * (xref_mode ? d.NX : d.nx);
    for (int q = t.tid; q < d.nu; q += 32) um1[q] = I.um1_solved[(size_t)inst * d.nu + q];
    t.sync();
    bmpc_prep(t, d, o, sys, I.x0 + (size_t)inst * d.nx, um1, xr, xref_mode, g, cc);
    double part = 0.0;
    for (int i = t.tid; i < d.NX; i += 32) {
        double xi = bmpc_Arow_dot(d, BcalT, U, i) + cc[i];
        double e = xi > hi0[i] ? hi0[i] - xi : (xi < lo0[i] ? lo0[i] - xi : 0.0);
        if (!(rho_e > 0.0)) e = 0.0;
        xseq[(size_t)inst * d.NX + i] = xi; epsseq[(size_t)inst * d.NX + i] = e;
        part += 0.5 * rho_e * e * e;
        // 0.5 cc' P_X cc + q_X' cc  with q_X = -P_X xref
        int k = i / d.nx, a = i % d.nx; const double* Q = k < d.Np ? Qx : QxN; double pc = 0.0, pr = 0.0;
        for (int q = 0; q < d.nx; q++) {
            pc += Q[a * d.nx + q] * cc[k * d.nx + q];
            pr += Q[a * d.nx + q] * (xref_mode ? xr[k * d.nx + q] : xr[q]);
        }
        part += cc[i] * (0.5 * pc - pr);
    }
    for (int a = t.tid; a < d.NU; a += 32) {
        double hu = 0.0;
        for (int b = 0; b < d.NU; b++) hu += H[a * d.NU + b] * U[b];
        part += U[a] * (0.5 * hu + g[a]);
    }
    part = t.sum(part);
    if (t.tid == 0) obj[inst] = part;
}

// ------------------------------------------------------------------------------------------------
// TPI fast path (bmpc_tpi.cuh): one thread per instance, one warp (32 instances) per CTA.
// The iterate travels between global [inst][mc] and shared [row][lane] through a coalesced transpose.
constexpr int TPI_STR = 33;   // padded lane stride of the shared-memory columns

template <class S>
__device__ __forceinline__ void tpi_load_v(const BmpcInst& I, int inst0, int nvalid, double* smem, int row_off, int nrows = S::MT) {
    // coalesced global read -> transposed shared write, issued as asynchronous 8-byte copies (LDGSTS) so that the 125
    // copies of a lane are all in flight at once instead of one load->store round trip per element
    const double* src = I.vw + (size_t)inst0 * S::mc;
    for (int idx = threadIdx.x; idx < nvalid * S::mc; idx += 32) {
        int t = idx / S::mc, i = idx - t * S::mc;
        if (i >= S::nx && i - S::nx < nrows) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (row_off + i - S::nx) * TPI_STR + t);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + idx) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <class S, bool TV>
__global__ void __launch_bounds__(32) k_tpi_admm(const __grid_constant__ TpiAdmmParams<S> P, BmpcInst I, const int32_t* __restrict__ list, int B, int niter, int cold,
                                                 int reset, int32_t* counts, double* um1_solved) {
    // list == nullptr: instances blockIdx*32 .. +31 (coalesced transpose of the iterate through shared memory);
    // list != nullptr: B arbitrary instances (straggler rounds): each thread moves its own rows directly.
    extern __shared__ double smem[];
    const int lane = threadIdx.x, idx0 = blockIdx.x * 32;
    const int nvalid = (B - idx0) < 32 ? (B - idx0) : 32;
    const bool valid = lane < nvalid;
    const int inst = list ? (valid ? list[idx0 + lane] : 0) : idx0 + lane;
    TpiAcc V{smem + lane, TPI_STR};
    if (!cold) {
        if (list) {
            if (valid) { const double* src = I.vw + (size_t)inst * S::mc + S::nx; for (int i = 0; i < S::MT; i++) V(i) = src[i]; }
        } else { tpi_load_v<S>(I, idx0, nvalid, smem, 0); }
        __syncwarp();
    }
    double x0[S::nx], um1[S::nu], xref[S::nx], x[S::NU];
#pragma unroll
    for (int q = 0; q < S::nx; q++) { x0[q] = valid ? I.x0[(size_t)inst * S::nx + q] : 0.0; xref[q] = (valid && !TV) ? I.xref[(size_t)inst * S::nx + q] : 0.0; }
    // constant reference: in registers; time-varying: this instance's (Np+1) x nx block in global memory
    const TpiXref<S, TV> xr{TV ? I.xref + (size_t)inst * S::NX : xref};
#pragma unroll
    for (int q = 0; q < S::nu; q++) um1[q] = valid ? I.um1[(size_t)inst * S::nu + q] : 0.0;
#pragma unroll
    for (int a = 0; a < S::NU; a++) x[a] = (valid && !cold) ? I.xw[(size_t)inst * S::NU + a] : 0.0;
    if (reset) {
        // first round of a solve: the per-solve bookkeeping that k_reset / two memsets would do rides here (3 launches saved)
        if (valid) {
            I.status[inst] = BMPC_UNSOLVED; I.iters[inst] = 0; I.psteps[inst] = 0; I.lvl[inst] = BMPC_LEV0;
#pragma unroll
            for (int q = 0; q < S::nu; q++) um1_solved[(size_t)inst * S::nu + q] = um1[q];
        }
        if (blockIdx.x == 0 && lane < BMPC_CNT) counts[lane] = 0;        // (the polish launch behind this one zeroes the other half)
    }
    // g' (read once per iteration) lives in the instance's global scratch row I.g, not in shared memory: the column is
    // MT rows instead of MT + NU, which lets one more warp reside per SM (measured: ADMM kernel 12 % faster)
    const TpiAcc G{I.g + (size_t)inst * S::NU, 1};
    if (valid) tpi_admm<S>(P, V, G, x0, um1, xr, x, niter, cold != 0);
    __syncwarp();
    if (list) {
        if (valid) {
            double* dst = I.vw + (size_t)inst * S::mc;
            for (int i = 0; i < S::MT; i++) dst[S::nx + i] = V(i);
#pragma unroll
            for (int q = 0; q < S::nx; q++) dst[q] = x0[q];
        }
    } else {
        double* dst = I.vw + (size_t)idx0 * S::mc;
#pragma unroll 5
        for (int idx = lane; idx < nvalid * S::mc; idx += 32) {
            int t = idx / S::mc, i = idx - t * S::mc;
            dst[idx] = (i >= S::nx) ? smem[(i - S::nx) * TPI_STR + t] : I.x0[(size_t)(idx0 + t) * S::nx + i];
        }
    }
    if (valid) {
#pragma unroll
        for (int a = 0; a < S::NU; a++) { I.xw[(size_t)inst * S::NU + a] = x[a]; I.Ua[(size_t)inst * S::NU + a] = x[a]; }
        I.iters[inst] += niter;
    }
}

// ------------------------------------------------------------------------------------------------
// K5 + K6 of the fast path: persistent, work-stealing Riccati polish (bmpc_tpi.cuh, second generation).
// One CTA per SM, TPI_POL_WARPS warps per CTA (as many as the gain workspace lets reside), every warp on its own: it
//   phase A  steals chunks of 32 instances off a cursor and runs capA active-set refinements on them (a warm solve: ONE, from
//            the stored working sets shifted by one stage); lanes that verified publish u*, U, status, v*; the others store
//            their updated working sets and enter a device-side queue;
//   phase B  when the cursor is exhausted the warp serves the queue: whatever instances are waiting (1..32) get up to capB more
//            refinements in-warp.  Stragglers therefore never make 31 verified lanes wait (the first-generation kernel lost
//            2/3 of its time to that on transient steps) and a warm solve is ONE launch whatever the refinement counts are.
// Instances still unverified after that go to next_list for the ADMM rounds of the host loop.
struct TpiPolArgs {
    const int32_t* list; int count;      // instances of phase A: list[0..count) or, list == nullptr, 0..count-1
    int mode;                            // 0 stored working sets, 1 stored and shifted one stage (receding horizon), 2 from the iterate v
    int capA, capB;                      // refinements per instance in phase A / in phase B (0: no phase B, failures go to next_list)
    int reset;                           // first round of a solve: per-solve bookkeeping rides here
    int32_t* counts_next;                // the 8 counters of the NEXT round: zeroed here (saves a memset launch per solve)
    int32_t* host_counts; int epoch;     // mapped pinned copy of the counters + the value the last warp stores into its epoch slot (0: off)
    long long* gflags; BmpcPeerFlags gpeers; int g_npeer, g_rank, g_world; long long g_epoch;   // K6 arrival folded into this launch (g_epoch 0: off)
    int32_t* next_list; int32_t* counts; // counts[0] unfinished (-> next_list), [1] refinements; counts[4..7] queue control, see below
    int32_t* queue; int qcap;            // phase-B queue (capacity qcap), all -1 between launches (consumers clear what they take)
    double* u0_out; double* um1_solved;
    const double* pview; int pstride;    // per-instance parameter blocks (PERINST instantiation): base and field stride (= batch)
    int cand_warm;                       // 1: an unverified instance hands its last candidate to the ADMM rounds when that candidate is sane
    unsigned char* codes; int code_stride;   // per instance: Np working-set codes + the multiplier scale (double) of the last refinement
};
constexpr int TPI_POL_WARPS = 7;

template <class S>
struct TpiPolLayout {
    using CT = typename TpiCode<S>::type;
    static constexpr int GROWS = S::Np * (S::nx + 2);                      // gain rows = TPI rows of v minus the spurious last one
    static constexpr size_t gain_bytes = (size_t)GROWS * TPI_STR * 8;
    static constexpr size_t code_bytes = (((size_t)S::Np * 32 * sizeof(CT)) + 15) & ~(size_t)15;
    static constexpr size_t per_warp = gain_bytes + code_bytes;
    static constexpr int code_stride = (int)(((S::Np * sizeof(CT) + 7) & ~(size_t)7) + 8);   // bytes per instance in global memory
    static constexpr int VROWS = S::MT - 1;                                 // TPI rows of v without the spurious last row
    static_assert(VROWS <= GROWS, "slot layout");
};

// one batch of up to 32 instances (one per lane) through up to cap refinements; returns the mask of verified lanes
template <class S, bool TV, class PP>
__device__ __forceinline__ void tpi_pol_batch(const PP& P, const BmpcInst& I, const TpiPolArgs& A, double* wsm, typename TpiCode<S>::type* csm,
                                              int inst, bool valid_in, int inst0_contig, int nvalid, int mode, int cap, bool to_queue, int reset) {
    bool valid = valid_in;
    using L = TpiPolLayout<S>; using CT = typename TpiCode<S>::type;
    constexpr int nx = S::nx;
    const int lane = threadIdx.x & 31;
    TpiAcc W{wsm + lane, TPI_STR};
    auto C = [&](int k) -> CT& { return csm[k * 32 + lane]; };
    double x0[nx], xref[nx], um1 = 0.0, mumax = 0.0, vq = 0.0;
#pragma unroll
    for (int q = 0; q < nx; q++) { x0[q] = valid ? I.x0[(size_t)inst * nx + q] : 0.0; xref[q] = (valid && !TV) ? I.xref[(size_t)inst * nx + q] : 0.0; }
    if (valid) um1 = I.um1[(size_t)inst];
    const TpiXref<S, TV> xr{TV ? I.xref + (size_t)inst * S::NX : xref};
    unsigned char* rec = A.codes + (size_t)inst * A.code_stride;
    if (mode == 2) {
        // working sets from the ADMM iterate: TPI rows of v into the (still unused) gain rows
        if (inst0_contig >= 0) tpi_load_v<S>(I, inst0_contig, nvalid, wsm, 0, L::VROWS);
        else if (valid) { const double* src = I.vw + (size_t)inst * S::mc + nx; for (int i = 0; i < L::VROWS; i++) W(i) = src[i]; }
        __syncwarp();
        if (valid) tpi2_codes_from_v<S>(P, um1, W, I.vw[(size_t)inst * S::mc + nx + S::MT - 1], C);
    } else if (valid) {
        // stored working sets (written by another SM when this batch comes from the queue: bypass L1)
        CT st[S::Np];
        const unsigned long long* src = (const unsigned long long*)rec;
        constexpr int NW = (int)((S::Np * sizeof(CT) + 7) / 8);
        unsigned long long wbuf[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) wbuf[i] = __ldcg(src + i);
#pragma unroll
        for (int k = 0; k < S::Np; k++) st[k] = (CT)(wbuf[(k * sizeof(CT)) / 8] >> (8 * ((k * sizeof(CT)) % 8)));
#pragma unroll
        for (int k = 0; k < S::Np; k++) C(k) = (CT)tpi2_shifted_code<S>(st, k, mode == 1);
        if (!reset) mumax = __ldcg((const double*)(rec + A.code_stride - 8));
    }
    double* udst = I.Us + (size_t)inst * S::NU;
    double u_first = 0.0;
    bool done = !valid, ok = false;
    if (mode == 2 && A.list != nullptr && !reset && valid) {
        // straggler round (an ADMM chunk of the team kernels ran just before): residuals far below any tolerance end the instance
        // as "solved, unpolished" — its iterate is the answer although the active-set iteration cannot certify it (bmpc_residuals_tight)
        if (bmpc_residuals_tight(I.res + (size_t)inst * 4, I.iters[inst])) {
            const double* ua = I.Ua + (size_t)inst * S::NU;
            for (int j = 0; j < S::NU; j++) udst[j] = ua[j];
            bmpc_publish_u0(I, A.u0_out, (size_t)inst, ua[0]);
            I.status[inst] = BMPC_SOLVED_UNPOLISHED; atomicAdd(A.counts + BMPC_CNT_TIGHT, 1);
            done = true; valid = false;
        }
    }
    int used = 0;
    for (int r = 0; r < cap; r++) {
        if (!done) {
            tpi2_backward<S>(P, W, C, xr);
            ok = tpi2_forward<S>(P, W, C, x0, um1, mumax, vq, [&](int j, double u) { udst[j] = u; if (j == 0) u_first = u; });
            used++;
            done = ok;
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
    ok = ok && valid;
    if (valid) {
        if (reset) {
            I.iters[inst] = 0; A.um1_solved[inst] = um1;
            if (!ok) { I.status[inst] = BMPC_UNSOLVED; I.lvl[inst] = BMPC_LEV0; }
        }
        if (ok) { bmpc_publish_u0(I, A.u0_out, (size_t)inst, u_first); I.status[inst] = BMPC_SOLVED; }
        // working sets of the accepted (or last) candidate: the next solve starts from them, phase B continues from them
        unsigned long long* dst = (unsigned long long*)rec;
        constexpr int NW = (int)((S::Np * sizeof(CT) + 7) / 8);
        unsigned long long wbuf[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) wbuf[i] = 0ull;
#pragma unroll
        for (int k = 0; k < S::Np; k++) wbuf[(k * sizeof(CT)) / 8] |= (unsigned long long)C(k) << (8 * ((k * sizeof(CT)) % 8));
#pragma unroll
        for (int i = 0; i < NW; i++) dst[i] = wbuf[i];
        *(double*)(rec + A.code_stride - 8) = mumax;
    }
    // refinement count: one atomic per warp
    int tot = used;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
    if (lane == 0 && tot) atomicAdd(A.counts + 1, tot);
    // unfinished lanes: to the queue of phase B (their working sets must be visible first) or to the host's list
    const bool fail = valid && !ok && (reset || I.status[inst] != BMPC_PRIMAL_INFEASIBLE);
    const unsigned fmask = __ballot_sync(0xffffffffu, fail);
    const int nf = __popc(fmask), pos = __popc(fmask & ((1u << lane) - 1u));
    if (to_queue) {
        // one 64-bit atomic publishes "nf more entries reserved" and "one more chunk of phase A finished" together, so a
        // consumer that reads the word sees a tail that is final once the chunk count is complete
        if (fmask) {
            __threadfence();
            unsigned long long old = 0ull;
            if (lane == 0) old = atomicAdd((unsigned long long*)(A.counts + TPI_Q_TAIL), (1ull << 32) | (unsigned long long)nf);
            const int basep = (int)(unsigned)__shfl_sync(0xffffffffu, old, 0);
            if (fail) ((volatile int32_t*)A.queue)[basep + pos] = inst;
        } else if (lane == 0) atomicAdd((unsigned long long*)(A.counts + TPI_Q_TAIL), 1ull << 32);     // no return value: fire and forget
    } else if (fmask) {
        int basep = 0;
        if (lane == 0) basep = atomicAdd(A.counts, nf);
        basep = __shfl_sync(0xffffffffu, basep, 0);
        if (fail) {
            A.next_list[basep + pos] = inst;
            // its ADMM rounds start from the exact fixed point v* of its last verified solve, unless the last candidate is sane (hard
            // rows feasible to 1e-2): then from the candidate's rows and multipliers v = z + mu / rho (bmpc_candidate_usable's policy)
            bool usable = A.cand_warm != 0 && mode == 2 && A.list != nullptr && bmpc_admm_stalled(I.res + (size_t)inst * 4, I.iters[inst]);   // straggler rounds only
            double up = um1;
            for (int j = 0; j < S::NU && usable; j++) {
                const double u = udst[j], dz = u - up;
                const double uh = P.uhi, ul = P.ulo, dh = P.dhi, dl = P.dlo;
                usable = u <= uh + 1e-2 * (1.0 + fabs(uh)) && u >= ul - 1e-2 * (1.0 + fabs(ul)) &&
                         dz <= dh + 1e-2 * (1.0 + fabs(dh)) && dz >= dl - 1e-2 * (1.0 + fabs(dl));
                up = u;
            }
            if (usable) {
                double* dst = I.vw + (size_t)inst * S::mc;
                for (int i = 0; i < L::VROWS; i++) dst[nx + i] = W(tpi_vstar_slot<S>(i));
                dst[S::mc - 1] = vq;
#pragma unroll
                for (int q = 0; q < nx; q++) dst[q] = x0[q];
            }
        }
    }
    // v* = z* + y*/rho, the exact ADMM fixed point of this problem, staged in the consumed gain slots by the forward sweep: the
    // warm start of the ADMM rounds a later solve may need for this instance (measured: starting those rounds from the failed
    // candidate instead costs 10x the rounds).  Verified lanes only.
    const unsigned okmask = __ballot_sync(0xffffffffu, ok);
    if (okmask) {
        __syncwarp();
        if (ok) {
            double* dst = I.vw + (size_t)inst * S::mc;
#pragma unroll
            for (int q = 0; q < nx; q++) dst[q] = x0[q];
            dst[S::mc - 1] = vq;
        }
        if (inst0_contig >= 0) {
            // coalesced: for every verified instance t of the chunk the warp writes its rows 32 at a time (lane = row)
            constexpr int NG = (L::VROWS + 31) / 32;
            int soff[NG];
#pragma unroll
            for (int c = 0; c < NG; c++) { const int i = c * 32 + lane; soff[c] = (i < L::VROWS) ? tpi_vstar_slot<S>(i) * TPI_STR : -1; }
            double* dst = I.vw + (size_t)inst0_contig * S::mc + nx + lane;
            if (okmask == 0xffffffffu) {
                // every instance of the chunk verified (the steady state): 4 instances per trip, all loads ahead of the stores
#pragma unroll 1
                for (int t = 0; t < 32; t += 4) {
                    double val[4][NG];
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int c = 0; c < NG; c++) val[j][c] = (soff[c] >= 0) ? wsm[soff[c] + t + j] : 0.0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int c = 0; c < NG; c++) if (soff[c] >= 0) dst[(size_t)(t + j) * S::mc + c * 32] = val[j][c];
                }
            } else {
#pragma unroll 2
                for (int t = 0; t < nvalid; t++) {
                    if (!((okmask >> t) & 1u)) continue;
#pragma unroll
                    for (int c = 0; c < NG; c++) if (soff[c] >= 0) dst[(size_t)t * S::mc + c * 32] = wsm[soff[c] + t];
                }
            }
        } else if (ok) {
            double* dst = I.vw + (size_t)inst * S::mc;
            for (int i = 0; i < L::VROWS; i++) dst[nx + i] = W(tpi_vstar_slot<S>(i));
        }
    }
    __syncwarp();
}

// per-instance parameter blocks (n_sys = batch): field-major in global memory, see TpiPolView
template <class S>
__global__ void k_tpi_fill_view(BmpcSysOff o, const double* __restrict__ sys, size_t sys_stride, int B, double* __restrict__ pg) {
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    TpiPolParams<S> P;
    tpi_fill_pol<S>(sys + (size_t)inst * sys_stride, o, P);
    const double* src = (const double*)&P;
    for (int f = 0; f < TpiPolView<S>::NF; f++) pg[(size_t)f * B + inst] = src[f];
}

template <class S, bool TV, bool PERINST>
__global__ void __launch_bounds__(TPI_POL_WARPS * 32, 1) k_tpi_pol(const __grid_constant__ TpiPolParams<S> P, BmpcInst I, TpiPolArgs A) {
    using L = TpiPolLayout<S>; using CT = typename TpiCode<S>::type;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wsm = (double*)((char*)smem + (size_t)warp * L::per_warp);
    CT* csm = (CT*)((char*)wsm + L::gain_bytes);
    const int nchunks = (A.count + 31) / 32;
    if (blockIdx.x == 0 && threadIdx.x < BMPC_CNT) A.counts_next[threadIdx.x] = 0;
    if (threadIdx.x == 0) atomicMax((unsigned long long*)(A.counts + BMPC_CNT_T0), ~bmpc_globaltimer());
    // ---- phase A
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(A.counts + TPI_Q_CURSOR, 1);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= nchunks) break;
        const int idx0 = chunk * 32, nvalid = (A.count - idx0) < 32 ? (A.count - idx0) : 32;
        const bool valid = lane < nvalid;
        const int inst = A.list ? (valid ? A.list[idx0 + lane] : 0) : idx0 + lane;
        if constexpr (PERINST) { const TpiPolView<S> Pv(A.pview + inst, (size_t)A.pstride); tpi_pol_batch<S, TV>(Pv, I, A, wsm, csm, inst, valid, A.list ? -1 : idx0, nvalid, A.mode, A.capA, A.capB > 0, A.reset); }
        else tpi_pol_batch<S, TV>(P, I, A, wsm, csm, inst, valid, A.list ? -1 : idx0, nvalid, A.mode, A.capA, A.capB > 0, A.reset);
    }
    auto leave = [&]() {
        if (lane != 0) return;
        atomicMax((unsigned long long*)(A.counts + BMPC_CNT_T1), bmpc_globaltimer());
        if (A.epoch == 0) return;
        __threadfence_system();                 // this warp's u* stores (to peers / to mapped host memory too) before its exit count
        if (atomicAdd(A.counts + BMPC_CNT_EXIT, 1) == (int)(gridDim.x * TPI_POL_WARPS) - 1) {
            // last warp out: every counter of the launch is final.
            __threadfence();
            volatile int32_t* src = (volatile int32_t*)A.counts; volatile int32_t* dst = (volatile int32_t*)A.host_counts;
            int gathered = 0;
            if (A.g_epoch > 0 && src[0] == 0) {
                // K6 arrival without a launch of its own: every instance of this rank is final and stored in every peer's gathered
                // buffer -> raise this rank's flag at the peers, then wait for theirs: when this kernel ends the gathered buffer of
                // the step is complete here.  (Instances left for straggler rounds: the host calls bmpc_gather_arrive afterwards.)
                __threadfence_system();
                for (int pr = 0; pr < A.g_npeer; pr++) *(volatile long long*)(A.gpeers.p[pr] + A.g_rank) = A.g_epoch;
                const long long t0 = clock64();
                for (int r = 0; r < A.g_world; r++) {
                    if (r == A.g_rank) continue;
                    while (*(volatile long long*)(A.gflags + r) < A.g_epoch) {
                        if (clock64() - t0 > 20000000000ll) __trap();        // ~10 s: a peer died; fail loudly instead of hanging the GPU
                        __nanosleep(100);
                    }
                }
                __threadfence_system();
                gathered = 1;
            }
            // ... and hand the counters to the host through mapped memory
            for (int i = 0; i < BMPC_CNT_EXIT; i++) dst[i] = src[i];
            dst[BMPC_CNT_TIGHT + 1] = gathered;
            __threadfence_system();
            dst[BMPC_CNT_EPOCH] = A.epoch;
        }
    };
    if (A.capB <= 0) { leave(); return; }
    // ---- phase B: serve the queue until every chunk of phase A is finished and the queue is empty.  Tickets: a warp takes the
    // next 32 queue slots with one fetch-and-add (no compare-and-swap races between a thousand warps) and waits until they are
    // filled or phase A has ended short of them, so batches are full while phase A still produces.
    volatile unsigned long long* ctl = (volatile unsigned long long*)(A.counts + TPI_Q_TAIL);
    for (;;) {
        int h = 0;
        if (lane == 0) h = atomicAdd(A.counts + TPI_Q_HEAD, 32);
        h = __shfl_sync(0xffffffffu, h, 0);
        const int idx = h + lane;
        int inst = -1;
        for (;;) {
            unsigned wlo = 0u, whi = 0u;
            if (lane == 0) { const unsigned long long w = *ctl; wlo = (unsigned)w; whi = (unsigned)(w >> 32); }   // tail | finished chunks
            wlo = __shfl_sync(0xffffffffu, wlo, 0); whi = __shfl_sync(0xffffffffu, whi, 0);
            const bool all_done = (int)whi >= nchunks;
            if (inst < 0 && idx < A.qcap && (idx < (int)wlo)) inst = ((volatile int32_t*)A.queue)[idx];
            const bool ready = inst >= 0 || (all_done && idx >= (int)wlo);
            if (__all_sync(0xffffffffu, ready)) break;
            __nanosleep(200);
        }
        const bool valid = inst >= 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (!vmask) break;                                            // ticket beyond the end of the queue: nothing left
        if (valid) ((volatile int32_t*)A.queue)[idx] = -1;            // the queue is all -1 again when the launch ends
        __threadfence();
        if constexpr (PERINST) { const TpiPolView<S> Pv(A.pview + (valid ? inst : 0), (size_t)A.pstride); tpi_pol_batch<S, TV>(Pv, I, A, wsm, csm, valid ? inst : 0, valid, -1, __popc(vmask), 0, A.capB, false, 0); }
        else tpi_pol_batch<S, TV>(P, I, A, wsm, csm, valid ? inst : 0, valid, -1, __popc(vmask), 0, A.capB, false, 0);
    }
    leave();
}


// ------------------------------------------------------------------------------------------------
// Multi-input fast path (bmpc_tpm.cuh): K5 of shapes with nu > 1 and one shared system — the constrained Riccati polish with one
// THREAD per instance and one warp (32 instances) per CTA.  Gain rows in global memory (per resident warp: slots x 32 lanes,
// lane-interleaved), working sets in local memory, the system in the constant bank (kernel parameter).
//  mode -1: warm first round of a solve: per instance, the stored working sets shifted one stage (record flag set: its last
//           solve was verified here) or, failing that, the working sets read off its ADMM fixed point v* of the previous
//           problem (left by whichever kernel finished it), shifted;
//  mode  2: straggler rounds: an ADMM chunk of the team / tile kernels ran just before; working sets from the iterate v.
// Verified: plan, u0 (published like every epilogue: output slice, shadow, peers), status, v* (standard row order: the warm
// start of any later ADMM round), x (ADMM warm start), working-set record.  Not verified: listed for the next round.
struct TpmArgs {
    const int32_t* list; int count, mode, cap, reset;
    int32_t* counts; int32_t* next_list; double* u0_out; double* um1_solved;
    double* W; unsigned long long* rec; int rec_stride;               // record of instance i: Np code words, mumax, flag (rec_stride 64-bit words)
};
struct TpmDiscard { unsigned long long sink; };

template <class S, bool TV, bool EX>
__global__ void __launch_bounds__(32) k_tpm_pol(const __grid_constant__ TpmParams<S> P, BmpcInst I, TpmArgs A) {
    using L = TpmLayout<S>; using CD = TpmCode<S>;
    constexpr int nx = S::nx, nu = S::nu, Np = S::Np, NU = S::NU, mc = S::mc, NX = S::NX;
    const int lane = threadIdx.x, idx = blockIdx.x * 32 + lane;
    const bool valid = idx < A.count;
    const int inst = valid ? (A.list ? A.list[idx] : idx) : 0;
    TpiStreamAcc W{A.W + (size_t)blockIdx.x * L::slots * 32 + lane, 32};
    uint64_t cur[Np], atb[Np];
    uint64_t dump = 0ull; uint64_t* const dp = &dump;
    auto C = [&](int k) -> uint64_t& { return cur[k]; };
    auto CB = [&](int k) -> uint64_t& { return atb[k]; };
    auto CK = [dp](int) -> uint64_t& { return *dp; };
    double x0[nx], um1[nu], xref[nx];
#pragma unroll
    for (int q = 0; q < nx; q++) { x0[q] = valid ? I.x0[(size_t)inst * nx + q] : 0.0; xref[q] = (valid && !TV) ? I.xref[(size_t)inst * nx + q] : 0.0; }
#pragma unroll
    for (int j = 0; j < nu; j++) um1[j] = valid ? I.um1[(size_t)inst * nu + j] : 0.0;
    const TpiXref<S, TV> xr{TV ? I.xref + (size_t)inst * NX : xref};
    unsigned long long* rec = A.rec + (size_t)inst * A.rec_stride;
    double mumax = 0.0;
    bool done = !valid, ok = false;
    if (valid) {
        const bool stored = A.mode < 0 && rec[Np + 1] == 1ull;
        if (stored) {
            uint64_t st[Np];
            for (int k = 0; k < Np; k++) st[k] = rec[k];
            unsigned first[nu];
#pragma unroll
            for (int j = 0; j < nu; j++) first[j] = (S::Nc > 1) ? tpm_first_label<S>(P, j, I.Us[(size_t)inst * NU + j], I.Us[(size_t)inst * NU + nu + j]) : 0u;
            const int tail = tpm_tail_start<S>(st);
            for (int k = 0; k < Np; k++) cur[k] = tpm_shifted_code<S>(st, k, true, first, tail);
            mumax = __longlong_as_double((long long)rec[Np]);
        } else {
            const double* v = I.vw + (size_t)inst * mc;
            tpm_codes_from_v<S>(P, um1, [&](int i) { return v[i]; }, C, A.mode < 0 ? 1 : 0);
        }
        if (A.mode == 2 && !A.reset && bmpc_residuals_tight(I.res + (size_t)inst * 4, I.iters[inst])) {
            // residuals far below any tolerance: the iterate is the answer although the active-set iteration cannot certify it
            const double* ua = I.Ua + (size_t)inst * NU;
            for (int a = 0; a < NU; a++) { I.Us[(size_t)inst * NU + a] = ua[a]; if (a < nu) bmpc_publish_u0(I, A.u0_out, (size_t)inst * nu + a, ua[a]); }
            I.status[inst] = BMPC_SOLVED_UNPOLISHED; atomicAdd(A.counts + BMPC_CNT_TIGHT, 1);
            done = true;
        }
    }
    const bool live = valid && !done;
    double vfirst[nu], vq = 0.0;
    double* udst = I.Us + (size_t)inst * NU;
    int used = 0;
    for (int r = 0; r < A.cap; r++) {
        if (!done) {
            tpm_backward<S>(P, W, C, xr, um1);
            const int fl = tpm_forward<S, EX>(P, W, C, CB, CK, x0, um1, mumax, vfirst, vq, [&](int k, int j, double u) { udst[k * nu + j] = u; });
            used++;
            ok = fl == 0; done = ok;
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
    ok = ok && live;
    if (valid && A.reset) {
        I.iters[inst] = 0; I.psteps[inst] = 0; I.lvl[inst] = BMPC_LEV0;
#pragma unroll
        for (int j = 0; j < nu; j++) A.um1_solved[(size_t)inst * nu + j] = um1[j];
        if (!ok) I.status[inst] = BMPC_UNSOLVED;
    }
    if (live) {
        I.psteps[inst] += used;
        if (ok) {
#pragma unroll
            for (int j = 0; j < nu; j++) bmpc_publish_u0(I, A.u0_out, (size_t)inst * nu + j, udst[j]);
            I.status[inst] = BMPC_SOLVED;
            double* xw = I.xw + (size_t)inst * NU;
            for (int a = 0; a < NU; a++) xw[a] = udst[a];
            // v* in the standard row order
            double* v = I.vw + (size_t)inst * mc;
#pragma unroll
            for (int q = 0; q < nx; q++) v[q] = x0[q];
            for (int k = 0; k < Np; k++)
                for (int p = 0; p < 2 * nu + nx; p++) { const int row = tpm_vstar_row<S>(k, p); if (row >= 0) v[row] = W.ld(k * L::per_stage + p); }
#pragma unroll
            for (int j = 0; j < nu; j++) v[NX + NU + j] = vfirst[j];
            v[mc - 1] = vq;
            for (int k = 0; k < Np; k++) rec[k] = atb[k];
            rec[Np] = (unsigned long long)__double_as_longlong(mumax); rec[Np + 1] = 1ull;
        } else {
            rec[Np + 1] = 0ull;                                         // whoever finishes this instance leaves v*: the next solve starts from that
        }
    }
    int tot = live ? used : 0;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
    if (lane == 0 && tot) atomicAdd(A.counts + 1, tot);
    const bool fail = live && !ok && (A.reset || I.status[inst] != BMPC_PRIMAL_INFEASIBLE);
    const unsigned fmask = __ballot_sync(0xffffffffu, fail);
    if (fmask) {
        int basep = 0;
        if (lane == 0) basep = atomicAdd(A.counts, __popc(fmask));
        basep = __shfl_sync(0xffffffffu, basep, 0);
        if (fail) A.next_list[basep + __popc(fmask & ((1u << lane) - 1u))] = inst;
    }
}

// Compiled multi-input fast-path shapes: one line per shape in csrc/tpm_shapes.inc
struct TpmEntry {
    int nx, nu, Np, Nc; unsigned long long amask; unsigned bmask; size_t par_bytes; int slots;
    bool (*fill)(const double* hs, const BmpcSysOff& o, void* pp);
    void (*launch)(struct bmpc_handle* h, const int32_t* list, int count, int mode, int cap, int reset, int32_t* next_list, int exchange);
};

// Compiled fast-path shapes (nx, nu, Np, Nc) with nu == 1 and Nc == Np: one line per shape in csrc/tpi_shapes.inc
// (`python -m pympc_b200.build --add-shape nx,1,Np` appends one and rebuilds).
struct TpiEntry {
    int nx, nu, Np, Nc;
    size_t admm_bytes, ric_bytes; int code_stride;
    void (*fill)(const double* hs, const BmpcSysOff& o, void* pa, void* pp);
    int view_fields; void (*fill_view)(struct bmpc_handle* h);
    int (*configure)();
    void (*launch)(struct bmpc_handle* h, const int32_t* list, int count, int niter, int32_t* next_list, cudaEvent_t mid);
    void (*launch_polish)(struct bmpc_handle* h, const int32_t* list, int count, int32_t* next_list);
};

// ------------------------------------------------------------------------------------------------
// host side
struct bmpc_handle {
    bmpc_config cfg;
    BmpcDims d;
    BmpcSysOff o;
    int team;        // 32 or CTA size
    int wpb;         // warps per block (warp team)
    int rmax;
    cudaStream_t stream, own_stream;
    double* sys = nullptr;
    BmpcInst I;
    double *x0 = nullptr, *um1 = nullptr, *um1_alt = nullptr, *um1_solved = nullptr, *xref = nullptr, *u0_own = nullptr, *u0_bound = nullptr;
    const double *x0_cur = nullptr, *um1_cur = nullptr;    // what the next solve reads: the handle's copies or buffers borrowed from the caller (bmpc_update on_device = 2)
    int cpar = 0;                                          // which half of counts[16] the round in flight uses (the fast-path kernel zeroes the other half for the next one)
    double *seq_x = nullptr, *seq_e = nullptr, *seq_obj = nullptr;
    int32_t *listA = nullptr, *listB = nullptr, *counts = nullptr;  // counts[8]: [0..3] round counters, [4..7] queue control of k_tpi_pol
    long long* gflags = nullptr; BmpcPeerFlags gpeers = {}; int g_npeer = 0, g_rank = 0, g_world = 1;   // K6 arrival flags
    long long g_epoch = 0, g_done_epoch = 0;       // one epoch per solve while flags are bound; the last epoch whose arrival is complete
    int32_t* queue = nullptr; unsigned char* codes = nullptr;        // phase-B queue and stored working sets of the fast-path polish
    int32_t* h_count = nullptr;                                      // pinned, mapped
    int epoch = 0, spin_epoch = 0;                                   // launch epochs of the fast-path kernel's host notification
    cudaEvent_t ev[4];
    int xref_mode = 0;
    const int32_t* status_filled = nullptr;   // host status array last filled with BMPC_SOLVED everywhere (skip refilling it)
    bool sync_after_round = false;     // bmpc_output queued a result copy behind the round: wait for the stream, not only for the kernel
    bool is_setup = false, cold = true, solved = false, committed = true;   // committed: this solve's u0 already is the next u_-1
    bmpc_stats stats;
    std::string err;
    size_t smem_admm = 0, smem_polish = 0;
    // low-latency CTA-per-instance variant for the few stragglers of a warp-team / TPI handle
    int fb_team = 0, fb_rmax = 0; size_t fb_smem_polish = 0;
    int rmax_small = 0; size_t smem_polish_small = 0; int32_t* ovf = nullptr;   // small-capacity polish tier (CTA teams) + its overflow list
    int sm_count = 148;
    double* vprev = nullptr; int32_t* lprev = nullptr;   // snapshot of (v, level) for the infeasibility check of straggler rounds
    void (*tile_fn[3])(BmpcDims, BmpcSysOff, const double*, BmpcInst, const int32_t*, int, int, int, int, int, int) = {nullptr, nullptr, nullptr};
    int tile_T = 0, tile_threads = 0;                  // > 0: the ADMM of this shape runs on tiles of T instances per CTA
    struct { const int32_t* list; int count; int32_t *cur, *nxt; int total, chunk, round; bool need_prep; int tight; bool cold; } st = {};
    bool pending = false;              // a round is in flight and has not been retired by the host yet
    int tpi_kind = 0;                  // 0 none, else 1 + index into g_tpi_table (compiled fast-path shapes)
    void *tpi_admm_params = nullptr, *tpi_polish_params = nullptr;   // host copies of the parameter blocks
    int tpi_pdas_steps = 8;
    double* tpi_view = nullptr;       // per-instance parameter blocks of the fast path (n_sys = batch), field-major
    bool codes_valid = false;         // the stored working sets come from a fast-path polish of the whole batch
    int tpm_kind = 0;                  // 0 none, else 1 + index into g_tpm_table (compiled multi-input fast-path shapes)
    void* tpm_params = nullptr;        // host copy of the parameter block
    double* tpm_W = nullptr; unsigned long long* tpm_rec = nullptr;   // gain rows of the resident warps, working-set records
    // refinements of the first launch / of a straggler round, ADMM iterations of the first straggler round.  (12, 4, 100) was measured on
    // the MIMO side bench with the all-at-once update in the straggler rounds too (tools/sweep_tpm_caps.sh: 13.3 ms/step at 8,6,25 ->
    // 11.4 ms); the straggler rounds now run in exchange mode, where 12 refinements finish every straggler of the host study in the
    // first round (all-at-once, cap 4: 54 %, three rounds for the rest) — chosen in host emulation, not yet swept on the GPU
    int tpm_first_cap = 12, tpm_round_cap = 12, tpm_chunk = 100;
    // cold solves: working sets read off a rough ADMM iterate need 8 - 14 refinements, not 4 (host study on 48 random MIMO starts,
    // first polish after 25 iterations: cap 4 verifies none and the solve takes 333 iterations in 4.5 rounds, cap 12 79 %, cap 16 92 %,
    // cap 24 all of them in the first round; after 50 iterations the same picture): every polish attempt of a cold solve gets this cap
    int tpm_cold_cap = 24;
};

static std::string g_create_err;

#define BMPC_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return BMPC_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

template <class S>
static void launch_tpi_pol(bmpc_handle* h, const int32_t* list, int count, int mode, int capA, int capB, int reset, int32_t* next_list) {
    using L = TpiPolLayout<S>;
    const TpiPolParams<S>& PP = *(const TpiPolParams<S>*)h->tpi_polish_params;
    TpiPolArgs A;
    A.list = list; A.count = count; A.mode = mode; A.capA = capA; A.capB = capB; A.reset = reset;
    A.next_list = next_list; A.counts = h->counts + BMPC_CNT * h->cpar; A.counts_next = h->counts + BMPC_CNT * (1 - h->cpar); A.queue = h->queue; A.qcap = h->cfg.batch; A.u0_out = h->I.u0; A.um1_solved = h->um1_solved;
    A.codes = h->codes; A.code_stride = L::code_stride; A.cand_warm = h->cfg.candidate_warm;
    const int nchunks = (count + 31) / 32;
    int grid = (nchunks + TPI_POL_WARPS - 1) / TPI_POL_WARPS;
    if (grid > h->sm_count) grid = h->sm_count;
    if (grid < 1) grid = 1;
    const size_t sm = L::per_warp * TPI_POL_WARPS;
    A.pview = h->tpi_view; A.pstride = h->cfg.batch;
    A.host_counts = h->h_count; A.epoch = h->spin_epoch;
    A.gflags = h->gflags; A.gpeers = h->gpeers; A.g_npeer = h->g_npeer; A.g_rank = h->g_rank; A.g_world = h->g_world;
    A.g_epoch = (h->gflags && h->spin_epoch && h->g_world > 1) ? h->g_epoch : 0;
    if (h->tpi_view) {
        if (h->xref_mode) k_tpi_pol<S, true, true><<<grid, TPI_POL_WARPS * 32, sm, h->stream>>>(PP, h->I, A);
        else k_tpi_pol<S, false, true><<<grid, TPI_POL_WARPS * 32, sm, h->stream>>>(PP, h->I, A);
    } else {
        if (h->xref_mode) k_tpi_pol<S, true, false><<<grid, TPI_POL_WARPS * 32, sm, h->stream>>>(PP, h->I, A);
        else k_tpi_pol<S, false, false><<<grid, TPI_POL_WARPS * 32, sm, h->stream>>>(PP, h->I, A);
    }
    h->stats.launches++;
}

// first round of a fast-path solve.  niter == 0 (warm start): ONE launch, the polish starts from the previous solution's working
// sets shifted by one stage (measured: a few ADMM iterations do not improve that first guess); niter > 0 (cold start,
// or first_iters set): thread-per-instance ADMM, then the polish takes its working sets from the iterate.
template <class S>
static void launch_tpi_round(bmpc_handle* h, const int32_t* list, int count, int niter, int32_t* next_list, cudaEvent_t mid) {
    const int grid = (count + 31) / 32;
    const TpiAdmmParams<S>& PA = *(const TpiAdmmParams<S>*)h->tpi_admm_params;
    const int cold = h->cold ? 1 : 0, reset = h->st.round == 0 ? 1 : 0;
    const size_t sa = S::MT * TPI_STR * 8;
    // (a cold solve reads its working sets off a rough ADMM iterate: twice the refinements — host study on 300 random pendulum starts,
    // first polish after 25 iterations: 88 % verify within 8 refinements, 99 % within 16, at most 3 rounds instead of 5)
    const int capA = 1, capB = (h->st.cold ? 2 * h->tpi_pdas_steps : h->tpi_pdas_steps) - 1;
    if (niter > 0) {
        if (h->xref_mode)   // one (Np+1) x nx reference per instance
            k_tpi_admm<S, true><<<grid, 32, sa, h->stream>>>(PA, h->I, list, count, niter, cold, reset, h->counts + BMPC_CNT * h->cpar, h->um1_solved);
        else
            k_tpi_admm<S, false><<<grid, 32, sa, h->stream>>>(PA, h->I, list, count, niter, cold, reset, h->counts + BMPC_CNT * h->cpar, h->um1_solved);
        h->stats.launches++;
        cudaEventRecord(mid, h->stream);
        launch_tpi_pol<S>(h, list, count, 2, capA, capB, 0, next_list);
    } else {
        // (this round's counters were zeroed by the previous fast-path launch, or by a memset)
        // working sets: stored ones (shifted one stage), or — when the last solve of the batch did not go through this kernel (cold
        // start on the team kernels) — from the iterate v
        launch_tpi_pol<S>(h, list, count, h->codes_valid ? (h->cfg.shift_warm ? 1 : 0) : 2, capA, capB, reset, next_list);
    }
    if (list == nullptr) h->codes_valid = true;
}

// straggler rounds: the listed instances come back from an ADMM chunk of the team kernels, working sets from their iterate
template <class S>
static void launch_tpi_polish_only(bmpc_handle* h, const int32_t* list, int count, int32_t* next_list) {
    h->spin_epoch = 0;
    launch_tpi_pol<S>(h, list, count, 2, 2, (h->st.cold ? 2 * h->tpi_pdas_steps : h->tpi_pdas_steps) - 2, 0, next_list);
}

template <class S>
static void tpi_fill_entry(const double* hs, const BmpcSysOff& o, void* pa, void* pp) {
    tpi_fill_admm<S>(hs, o, *(TpiAdmmParams<S>*)pa); tpi_fill_pol<S>(hs, o, *(TpiPolParams<S>*)pp);
}
template <class S>
static void tpi_fill_view_entry(bmpc_handle* h) {
    const int B = h->cfg.batch;
    k_tpi_fill_view<S><<<(B + 63) / 64, 64, 0, h->stream>>>(h->o, h->sys, (size_t)h->o.total, B, h->tpi_view);
}
template <class S>
static int tpi_configure_entry() {
    if (cudaFuncSetAttribute(k_tpi_admm<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(S::MT * TPI_STR * 8)) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tpi_admm<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(S::MT * TPI_STR * 8)) != cudaSuccess) return 1;
    const int sm = (int)(TpiPolLayout<S>::per_warp * TPI_POL_WARPS);
    if (cudaFuncSetAttribute(k_tpi_pol<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tpi_pol<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tpi_pol<S, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(k_tpi_pol<S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess) return 1;
    return 0;
}
#define BMPC_TPI_SHAPE(NX_, NU_, NP_, NC_)                                                                                   \
    {NX_, NU_, NP_, NC_, sizeof(TpiAdmmParams<TpiShape<NX_, NU_, NP_, NC_>>), sizeof(TpiPolParams<TpiShape<NX_, NU_, NP_, NC_>>), \
     TpiPolLayout<TpiShape<NX_, NU_, NP_, NC_>>::code_stride,                                                                \
     tpi_fill_entry<TpiShape<NX_, NU_, NP_, NC_>>, TpiPolView<TpiShape<NX_, NU_, NP_, NC_>>::NF, tpi_fill_view_entry<TpiShape<NX_, NU_, NP_, NC_>>, \
     tpi_configure_entry<TpiShape<NX_, NU_, NP_, NC_>>,                                                                        \
     launch_tpi_round<TpiShape<NX_, NU_, NP_, NC_>>, launch_tpi_polish_only<TpiShape<NX_, NU_, NP_, NC_>>},
// BMPC_TPI_SHAPES_FILE: pympc_b200.build.jit_shape() compiles a copy of the library whose table holds the one shape a controller
// asked for (any nx, Np with nu == 1) when the in-tree table does not have it
#ifndef BMPC_TPI_SHAPES_FILE
#define BMPC_TPI_SHAPES_FILE "tpi_shapes.inc"
#endif
static const TpiEntry g_tpi_table[] = {
#include BMPC_TPI_SHAPES_FILE
};
#undef BMPC_TPI_SHAPE
static const int g_tpi_count = (int)(sizeof(g_tpi_table) / sizeof(g_tpi_table[0]));

template <class S>
static void launch_tpm(bmpc_handle* h, const int32_t* list, int count, int mode, int cap, int reset, int32_t* next_list, int exchange) {
    TpmArgs A;
    A.list = list; A.count = count; A.mode = mode; A.cap = cap; A.reset = reset;
    A.counts = h->counts + BMPC_CNT * h->cpar; A.next_list = next_list; A.u0_out = h->I.u0; A.um1_solved = h->um1_solved;
    A.W = h->tpm_W; A.rec = h->tpm_rec; A.rec_stride = S::Np + 2;
    const TpmParams<S>& P = *(const TpmParams<S>*)h->tpm_params;
    const int grid = (count + 31) / 32;
    // exchange: the straggler rounds' instantiation (hard rows change by single exchanges, tpm_forward<S, true>)
    if (exchange) {
        if (h->xref_mode) k_tpm_pol<S, true, true><<<grid, 32, 0, h->stream>>>(P, h->I, A);
        else k_tpm_pol<S, false, true><<<grid, 32, 0, h->stream>>>(P, h->I, A);
    } else {
        if (h->xref_mode) k_tpm_pol<S, true, false><<<grid, 32, 0, h->stream>>>(P, h->I, A);
        else k_tpm_pol<S, false, false><<<grid, 32, 0, h->stream>>>(P, h->I, A);
    }
    h->stats.launches++;
}
template <class S>
static bool tpm_fill_entry(const double* hs, const BmpcSysOff& o, void* pp) { return tpm_fill<S>(hs, o, *(TpmParams<S>*)pp); }
#define BMPC_TPM_ENTRY(S_) {S_::nx, S_::nu, S_::Np, S_::Nc, S_::amask, S_::bmask, sizeof(TpmParams<S_>), TpmLayout<S_>::slots, tpm_fill_entry<S_>, launch_tpm<S_>},
#define BMPC_TPM_SHAPE(NX_, NU_, NP_, NC_) BMPC_TPM_ENTRY(BMPC_TPM_T4(TpiShape, NX_, NU_, NP_, NC_))
#define BMPC_TPM_SPARSE_SHAPE(NX_, NU_, NP_, NC_, AM_, BM_) BMPC_TPM_ENTRY(BMPC_TPM_T6(TpmSparseShape, NX_, NU_, NP_, NC_, AM_, BM_))
#define BMPC_TPM_T4(T_, a, b, c, d) T_<a, b, c, d>
#define BMPC_TPM_T6(T_, a, b, c, d, e, f) T_<a, b, c, d, e, f>
// (entries are tried in order: sparse patterns first, the dense instantiation of a shape last)
#ifndef BMPC_TPM_SHAPES_FILE
#define BMPC_TPM_SHAPES_FILE "tpm_shapes.inc"
#endif
static const TpmEntry g_tpm_table[] = {
#include BMPC_TPM_SHAPES_FILE
};
#undef BMPC_TPM_SHAPE
#undef BMPC_TPM_SPARSE_SHAPE
static const int g_tpm_count = (int)(sizeof(g_tpm_table) / sizeof(g_tpm_table[0]));


extern "C" {

static int finish_solve(bmpc_handle* h);

void bmpc_default_config(bmpc_config* c) {
    memset(c, 0, sizeof(*c));
    c->Np = 20; c->Nc = 0; c->batch = 1; c->device = 0; c->soft_on = 1;
    c->max_iter = 4000; c->first_iters = 0; c->pdas_steps = 10; c->rmax = 0; c->polish = 1;
    c->team_threads = 0; c->warps_per_block = 0; c->fast_path = 1; c->n_sys = 1; c->shift_warm = 1; c->candidate_warm = 0; c->cold_iters = 0;
    c->eps_feas = 1e6; c->rho = 0.0; c->sigma = 1e-6; c->alpha = 1.6; c->eps_abs = 1e-3; c->eps_rel = 1e-3;
}

int bmpc_has_fast_path(int nx, int nu, int Np, int Nc) {
    if (Nc <= 0) Nc = Np;
    for (int k = 0; k < g_tpi_count; k++)
        if (g_tpi_table[k].nx == nx && g_tpi_table[k].nu == nu && g_tpi_table[k].Np == Np && g_tpi_table[k].Nc == Nc) return 1;
    return 0;
}

int bmpc_has_multi_input_fast_path(int nx, int nu, int Np, int Nc) {
    if (Nc <= 0) Nc = Np;
    for (int k = 0; k < g_tpm_count; k++)
        if (g_tpm_table[k].nx == nx && g_tpm_table[k].nu == nu && g_tpm_table[k].Np == Np && g_tpm_table[k].Nc == Nc) return 1;
    return 0;
}

int bmpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* bmpc_last_error(const bmpc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

void* bmpc_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    // mapped + portable: the device can read / write the buffer in place (bmpc_update on_device = 2, bmpc_bind_output): with unified
    // addressing the host pointer is the device pointer
    if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void bmpc_host_free(void* p) { if (p) cudaFreeHost(p); }

static int configure_launch(bmpc_handle* h) {
    const BmpcDims& d = h->d;
    int dev = h->cfg.device, max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    size_t budget = (size_t)max_optin - 1024;
    int team = h->cfg.team_threads;
    if (team <= 0) team = (d.mc <= 192 && d.NU <= 64) ? 32 : 256;
    if (team != 32) { team = ((team + 31) / 32) * 32; if (team > 1024) team = 1024; }
    h->team = team;
    int rmax = h->cfg.rmax > 0 ? h->cfg.rmax : (team == 32 ? 64 : 224);     // S is packed: r(r+1)/2 doubles
    if (rmax > d.mc) rmax = d.mc;
    // shrink rmax until one instance fits
    while (rmax > 8 && polish_smem_doubles(d, rmax) * 8 > budget) rmax -= 8;
    if (polish_smem_doubles(d, rmax) * 8 > budget || admm_smem_doubles(d) * 8 > budget) {
        h->err = "problem too large for the shared-memory resident kernels"; return BMPC_ERR_ARG;
    }
    h->rmax = rmax;
    auto setup_tiles = [&]() -> int {
        h->tile_T = 0;
        if (h->cfg.team_threads == 0 && h->cfg.n_sys <= 1) {
            for (int T : {8, 4, 2}) if (!h->tile_T && bmpc_tile_smem_doubles(d, T) * 8 <= budget) h->tile_T = T;
            // kernels per tile size {8, 4, 2}; shapes with a compiled (nx, nu) get the unrolled instantiation
            if (d.nx == 8 && d.nu == 4) { h->tile_fn[0] = k_admm_tile<8, 2, 8, 4>; h->tile_fn[1] = k_admm_tile<4, 2, 8, 4>; h->tile_fn[2] = k_admm_tile<2, 1, 8, 4>; }
            else { h->tile_fn[0] = k_admm_tile<8, 2, 0, 0>; h->tile_fn[1] = k_admm_tile<4, 2, 0, 0>; h->tile_fn[2] = k_admm_tile<2, 1, 0, 0>; }
            for (int k = 0; k < 3; k++) {
                const int T = 8 >> k;
                if (h->tile_T >= T) BMPC_CUDA(cudaFuncSetAttribute((const void*)h->tile_fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bmpc_tile_smem_doubles(d, T) * 8)));
            }
            int th = d.NX > 2 * d.NU ? d.NX : 2 * d.NU;
            th = ((th + 31) / 32) * 32; if (th < 128) th = 128; if (th > 512) th = 512;
            if (h->cfg.warps_per_block > 0 && h->team != 32) th = h->cfg.warps_per_block * 32;      // tuning override
            if (th > 512) th = 512;
            h->tile_threads = th;
        }
        return BMPC_OK;
    };
    if (team == 32) {

        int wpb = h->cfg.warps_per_block > 0 ? h->cfg.warps_per_block : 8;
        while (wpb > 1 && (size_t)wpb * polish_smem_doubles(d, rmax) * 8 > budget) wpb--;
        h->wpb = wpb;
        h->smem_admm = (size_t)wpb * admm_smem_doubles(d) * 8;
        h->smem_polish = (size_t)wpb * polish_smem_doubles(d, rmax) * 8;
        BMPC_CUDA(cudaFuncSetAttribute(k_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_admm));
        BMPC_CUDA(cudaFuncSetAttribute(k_polish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_polish));
        // stragglers (instances the first round did not verify) are few and latency-bound: give each a whole CTA
        int frmax = d.mc < 128 ? d.mc : 128;
        while (frmax > 8 && polish_smem_doubles(d, frmax) * 8 > budget) frmax -= 8;
        if (polish_smem_doubles(d, frmax) * 8 <= budget) {
            h->fb_team = 128; h->fb_rmax = frmax;
            h->fb_smem_polish = polish_smem_doubles(d, frmax) * 8;
            BMPC_CUDA(cudaFuncSetAttribute(k_polish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fb_smem_polish));
        }
    } else {
        // large shapes with one shared system: the ADMM runs on tiles of T instances per CTA (bmpc_tile.cuh)
        { int rc = setup_tiles(); if (rc) return rc; }
        if (h->cfg.rmax == 0 && rmax > 96) { h->rmax_small = 80; h->smem_polish_small = polish_smem_doubles(d, 80) * 8; }
        h->wpb = team / 32;
        h->smem_admm = admm_smem_doubles(d) * 8;
        h->smem_polish = polish_smem_doubles(d, rmax) * 8;
        BMPC_CUDA(cudaFuncSetAttribute(k_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_admm));
        BMPC_CUDA(cudaFuncSetAttribute(k_polish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_polish));
    }
    return BMPC_OK;
}

int bmpc_create(const bmpc_config* cfg, bmpc_handle** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return BMPC_ERR_ARG; }
    *out = nullptr;
    if (cfg->nx < 1 || cfg->nu < 1 || cfg->Np < 2 || cfg->batch < 1 || (cfg->Nc > cfg->Np)) {
        g_create_err = "invalid dimensions (need nx,nu >= 1, Np > 1, Nc <= Np, batch >= 1)"; return BMPC_ERR_ARG;
    }
    int ndev = bmpc_device_count();
    if (ndev <= 0) { g_create_err = "no CUDA device visible: libbmpc has no CPU fallback"; return BMPC_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "device ordinal out of range"; return BMPC_ERR_ARG; }
    if (cfg->n_sys != 0 && cfg->n_sys != 1 && cfg->n_sys != cfg->batch) { g_create_err = "n_sys must be 1 (shared system) or batch (one system per instance)"; return BMPC_ERR_ARG; }
    bmpc_handle* h = new bmpc_handle();
    h->cfg = *cfg;
    if (h->cfg.n_sys <= 0) h->cfg.n_sys = 1;
    if (h->cfg.Nc <= 0) h->cfg.Nc = h->cfg.Np;
    if (h->cfg.max_iter <= 0) h->cfg.max_iter = 4000;
    if (h->cfg.pdas_steps <= 0) h->cfg.pdas_steps = 10;
    h->d = bmpc_make_dims(cfg->nx, cfg->nu, h->cfg.Np, h->cfg.Nc);
    h->o = bmpc_make_off(h->d);
    memset(&h->I, 0, sizeof(h->I)); memset(&h->stats, 0, sizeof(h->stats));
    auto fail = [&](int code) { g_create_err = h->err; bmpc_destroy(h); return code; };
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return fail(BMPC_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream create failed"; return fail(BMPC_ERR_CUDA); }
    h->stream = h->own_stream;
    for (int i = 0; i < 4; i++) cudaEventCreate(&h->ev[i]);
    int rc = configure_launch(h);
    if (rc) return fail(rc);
    const BmpcDims& d = h->d; size_t B = cfg->batch;
    auto dalloc = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 8) == cudaSuccess; };
    bool ok = true;
    ok &= dalloc((void**)&h->sys, sizeof(double) * (size_t)h->o.total * h->cfg.n_sys);
    ok &= dalloc((void**)&h->x0, sizeof(double) * B * d.nx);
    ok &= dalloc((void**)&h->um1, sizeof(double) * B * d.nu);
    ok &= dalloc((void**)&h->um1_alt, sizeof(double) * B * d.nu);
    ok &= dalloc((void**)&h->um1_solved, sizeof(double) * B * d.nu);
    ok &= dalloc((void**)&h->xref, sizeof(double) * B * d.NX);
    ok &= dalloc((void**)&h->u0_own, sizeof(double) * B * d.nu);
    ok &= dalloc((void**)&h->I.g, sizeof(double) * B * d.NU);
    ok &= dalloc((void**)&h->I.cc, sizeof(double) * B * d.NX);
    ok &= dalloc((void**)&h->I.xw, sizeof(double) * B * d.NU);
    ok &= dalloc((void**)&h->I.vw, sizeof(double) * B * d.mc);
    ok &= dalloc((void**)&h->I.Ua, sizeof(double) * B * d.NU);
    ok &= dalloc((void**)&h->I.Us, sizeof(double) * B * d.NU);
    ok &= dalloc((void**)&h->I.res, sizeof(double) * B * 4);
    ok &= dalloc((void**)&h->I.status, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->I.iters, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->I.psteps, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->I.lvl, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->listA, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->listB, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->counts, sizeof(int32_t) * 2 * BMPC_CNT);
    ok &= dalloc((void**)&h->queue, sizeof(int32_t) * B);
    ok &= dalloc((void**)&h->codes, (size_t)B * 72);       // largest record: 32 stages x 2 bytes + 8 (Np < 32 on the fast path)
    ok &= dalloc((void**)&h->ovf, sizeof(int32_t) * (size_t)B);
    ok &= dalloc((void**)&h->vprev, sizeof(double) * (size_t)B * d.mc);
    ok &= dalloc((void**)&h->lprev, sizeof(int32_t) * (size_t)B);
    if (!ok) { h->err = "cudaMalloc failed"; cudaGetLastError(); return fail(BMPC_ERR_CUDA); }
    if (cudaHostAlloc((void**)&h->h_count, sizeof(int32_t) * BMPC_CNT, cudaHostAllocMapped) != cudaSuccess) { h->err = "cudaHostAlloc failed"; return fail(BMPC_ERR_CUDA); }
    cudaMemset(h->sys, 0, sizeof(double) * (size_t)h->o.total * h->cfg.n_sys);
    cudaMemset(h->x0, 0, sizeof(double) * B * d.nx);
    cudaMemset(h->um1, 0, sizeof(double) * B * d.nu);
    cudaMemset(h->xref, 0, sizeof(double) * B * d.NX);
    cudaMemset(h->u0_own, 0, sizeof(double) * B * d.nu);
    cudaMemset(h->I.Us, 0, sizeof(double) * B * d.NU);
    cudaMemset(h->I.Ua, 0, sizeof(double) * B * d.NU);
    cudaMemset(h->queue, 0xff, sizeof(int32_t) * B);
    cudaMemset(h->codes, 0, (size_t)B * 72);
    cudaMemset(h->counts, 0, sizeof(int32_t) * 2 * BMPC_CNT);
    cudaMemset(h->um1_alt, 0, sizeof(double) * B * d.nu);
    h->I.sys_stride = h->cfg.n_sys > 1 ? (size_t)h->o.total : 0;
    h->I.x0 = h->x0; h->I.um1 = h->um1; h->I.um1_solved = h->um1_solved; h->I.xref = h->xref; h->I.u0 = h->u0_own; h->I.u0_shadow = h->um1_alt;
    h->x0_cur = h->x0; h->um1_cur = h->um1;
    k_reset<<<(int)((B + 255) / 256), 256, 0, h->stream>>>(h->I, (int)B);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "device initialisation failed"; return fail(BMPC_ERR_CUDA); }
    *out = h;
    return BMPC_OK;
}

void bmpc_destroy(bmpc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->pending) cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->sys, h->x0, h->um1, h->um1_alt, h->um1_solved, h->xref, h->u0_own, h->I.g, h->I.cc, h->I.xw, h->I.vw, h->I.Ua, h->I.Us, h->I.res,
                    h->I.status, h->I.iters, h->I.psteps, h->I.lvl, h->listA, h->listB, h->counts, h->queue, h->codes, h->ovf, h->vprev, h->lprev, h->seq_x, h->seq_e, h->seq_obj};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_count) cudaFreeHost(h->h_count);
    free(h->tpi_admm_params); free(h->tpi_polish_params); free(h->tpm_params);
    if (h->tpi_view) cudaFree(h->tpi_view);
    if (h->tpm_W) cudaFree(h->tpm_W);
    if (h->tpm_rec) cudaFree(h->tpm_rec);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int bmpc_set_stream(bmpc_handle* h, void* s) {
    if (!h) return BMPC_ERR_ARG;
    if (h->pending) { cudaSetDevice(h->cfg.device); int rc = finish_solve(h); if (rc) return rc; }
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return BMPC_OK;
}

int bmpc_synchronize(bmpc_handle* h) {
    if (!h) return BMPC_ERR_ARG;
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    if (h->pending) { int rc = finish_solve(h); if (rc) return rc; }
    BMPC_CUDA(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_bind_output(bmpc_handle* h, double* dev_u0) {
    if (!h) return BMPC_ERR_ARG;
    // kernels in flight captured the old pointer: retire them first (the committed u_-1 of that solve stays what it was)
    if (h->pending) { cudaSetDevice(h->cfg.device); int rc = finish_solve(h); if (rc) return rc; }
    h->u0_bound = dev_u0;
    h->I.u0 = dev_u0 ? dev_u0 : h->u0_own;
    return BMPC_OK;
}

int bmpc_bind_output_peers(bmpc_handle* h, double* const* peer_u0, int n) {
    if (!h || n < 0 || n > 8 || (n > 0 && !peer_u0)) return BMPC_ERR_ARG;
    if (h->pending) { cudaSetDevice(h->cfg.device); int rc = finish_solve(h); if (rc) return rc; }
    h->I.n_peer = n;
    for (int p = 0; p < n; p++) h->I.u0_peer[p] = peer_u0[p];
    return BMPC_OK;
}

int bmpc_bind_gather_flags(bmpc_handle* h, int64_t* my_flags, int64_t* const* peer_flags, int n_peers, int rank, int world, int64_t base_epoch) {
    if (!h || n_peers < 0 || n_peers > 8 || world < 1 || world > 9 || rank < 0 || rank >= world || (n_peers > 0 && (!my_flags || !peer_flags))) return BMPC_ERR_ARG;
    h->gflags = (long long*)my_flags; h->g_npeer = n_peers; h->g_rank = rank; h->g_world = world;
    h->g_epoch = h->g_done_epoch = base_epoch;
    for (int p = 0; p < n_peers; p++) h->gpeers.p[p] = (long long*)peer_flags[p];
    return BMPC_OK;
}

int bmpc_gather_arrive(bmpc_handle* h, int64_t epoch) {
    if (!h) return BMPC_ERR_ARG;
    if (!h->gflags) { h->err = "bmpc_gather_arrive before bmpc_bind_gather_flags"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    if (h->pending) { int rc = finish_solve(h); if (rc) return rc; }
    (void)epoch;                                       // the library counts the epochs itself (one per solve on every rank)
    if (h->g_done_epoch == h->g_epoch) return BMPC_OK;  // the solver kernel's last warp already signalled and waited
    k_gather_arrive<<<1, 32, 0, h->stream>>>(h->gflags, h->gpeers, h->g_npeer, h->g_rank, h->g_world, h->g_epoch);
    h->g_done_epoch = h->g_epoch;
    BMPC_CUDA(cudaGetLastError());
    return BMPC_OK;
}

int bmpc_get_dims(const bmpc_handle* h, int32_t* dims) {
    if (!h || !dims) return BMPC_ERR_ARG;
    dims[0] = h->d.nx; dims[1] = h->d.nu; dims[2] = h->d.Np; dims[3] = h->d.Nc; dims[4] = h->d.NX; dims[5] = h->d.NU;
    dims[6] = h->d.mc; dims[7] = h->team;
    return BMPC_OK;
}

int bmpc_setup(bmpc_handle* h, const double* Ad, const double* Bd, const double* Qx, const double* QxN, const double* Qu,
               const double* QDu, const double* xmin, const double* xmax, const double* umin, const double* umax,
               const double* Dumin, const double* Dumax, const double* uref) {
    if (!h) return BMPC_ERR_ARG;
    if (!Ad || !Bd || !Qx || !QxN || !Qu || !QDu || !xmin || !xmax || !umin || !umax || !Dumin || !Dumax || !uref) {
        h->err = "bmpc_setup: null pointer"; return BMPC_ERR_ARG;
    }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    const BmpcDims& d = h->d; const BmpcSysOff& o = h->o;
    const int ns = h->cfg.n_sys;
    std::vector<double> in((size_t)o.pw * ns, 0.0);
    for (int si = 0; si < ns; si++) {
        double* blk = in.data() + (size_t)si * o.pw;
        auto put = [&](int off, const double* src, int cnt) { memcpy(blk + off, src + (size_t)si * cnt, sizeof(double) * cnt); };
        put(o.Ad, Ad, d.nx * d.nx); put(o.Bd, Bd, d.nx * d.nu); put(o.Qx, Qx, d.nx * d.nx); put(o.QxN, QxN, d.nx * d.nx);
        put(o.Qu, Qu, d.nu * d.nu); put(o.QDu, QDu, d.nu * d.nu); put(o.xmin, xmin, d.nx); put(o.xmax, xmax, d.nx);
        put(o.umin, umin, d.nu); put(o.umax, umax, d.nu); put(o.Dumin, Dumin, d.nu); put(o.Dumax, Dumax, d.nu); put(o.uref, uref, d.nu);
    }
    BMPC_CUDA(cudaMemcpy2DAsync(h->sys, sizeof(double) * o.total, in.data(), sizeof(double) * o.pw, sizeof(double) * o.pw, ns,
                                cudaMemcpyHostToDevice, h->stream));
    k_condense<<<ns, 256, 0, h->stream>>>(d, o, h->sys, h->cfg.rho, h->cfg.sigma, h->cfg.alpha, h->cfg.eps_feas, h->cfg.soft_on);
    BMPC_CUDA(cudaGetLastError());
    std::vector<double> errs(ns);
    BMPC_CUDA(cudaMemcpy2DAsync(errs.data(), sizeof(double), h->sys + o.scal + BMPC_S_ERR, sizeof(double) * o.total, sizeof(double), ns,
                                cudaMemcpyDeviceToHost, h->stream));
    BMPC_CUDA(cudaStreamSynchronize(h->stream));
    for (int si = 0; si < ns; si++) if (errs[si] != 0.0) {
        h->err = "condensed Hessian is not positive definite (Qu/QDu/Qx make the QP non-strictly convex in U)";
        return BMPC_ERR_NOT_PD;
    }
    // thread-per-instance fast path for the compiled small shapes (pendulum, point mass)
    h->tpi_kind = 0;
    // (the Riccati polish treats state rows as the soft penalty they are by default; hard state rows -> team kernels)
    if (h->cfg.fast_path && h->team == 32 && h->cfg.soft_on) {
        for (int k = 0; k < g_tpi_count; k++)
            if (g_tpi_table[k].nx == d.nx && g_tpi_table[k].nu == d.nu && g_tpi_table[k].Np == d.Np && g_tpi_table[k].Nc == d.Nc) { h->tpi_kind = k + 1; break; }
    }
    if (h->tpi_kind) {
        const TpiEntry& te = g_tpi_table[h->tpi_kind - 1];
        std::vector<double> hs(o.total);
        BMPC_CUDA(cudaMemcpyAsync(hs.data(), h->sys, sizeof(double) * o.total, cudaMemcpyDeviceToHost, h->stream));
        BMPC_CUDA(cudaStreamSynchronize(h->stream));
        free(h->tpi_admm_params); free(h->tpi_polish_params);
        h->tpi_admm_params = malloc(te.admm_bytes); h->tpi_polish_params = malloc(te.ric_bytes);
        te.fill(hs.data(), o, h->tpi_admm_params, h->tpi_polish_params);          // system 0 (the only one when the system is shared)
        if (te.configure()) { h->err = "fast path: shared-memory configuration failed"; return BMPC_ERR_CUDA; }
        if (h->tpi_view) { cudaFree(h->tpi_view); h->tpi_view = nullptr; }
        if (ns > 1) {
            // one parameter block per instance, built on the device from each instance's condensed system block
            BMPC_CUDA(cudaMalloc((void**)&h->tpi_view, sizeof(double) * (size_t)te.view_fields * h->cfg.batch));
            te.fill_view(h);
            BMPC_CUDA(cudaGetLastError());
        }
    }
    // multi-input fast path: compiled shape, one shared system, soft state rows, diagonal QDu
    h->tpm_kind = 0;
    if (h->cfg.fast_path && !h->tpi_kind && h->cfg.soft_on && ns == 1 && h->cfg.polish) {
        // sparsity pattern of (Ad, Bd): an entry fits when its compile-time masks contain it
        unsigned long long am = 0ull; unsigned bm = 0u;
        if (d.nx * d.nx <= 64 && d.nx * d.nu <= 32) {
            for (int i = 0; i < d.nx * d.nx; i++) if (in[o.Ad + i] != 0.0) am |= 1ull << i;
            for (int i = 0; i < d.nx * d.nu; i++) if (in[o.Bd + i] != 0.0) bm |= 1u << i;
        } else { am = ~0ull; bm = ~0u; }
        for (int k = 0; k < g_tpm_count; k++)
            if (g_tpm_table[k].nx == d.nx && g_tpm_table[k].nu == d.nu && g_tpm_table[k].Np == d.Np && g_tpm_table[k].Nc == d.Nc &&
                (am & ~g_tpm_table[k].amask) == 0ull && (bm & ~g_tpm_table[k].bmask) == 0u) { h->tpm_kind = k + 1; break; }
    }
    if (h->tpm_kind) {
        const TpmEntry& te = g_tpm_table[h->tpm_kind - 1];
        std::vector<double> hs(o.total);
        BMPC_CUDA(cudaMemcpyAsync(hs.data(), h->sys, sizeof(double) * o.total, cudaMemcpyDeviceToHost, h->stream));
        BMPC_CUDA(cudaStreamSynchronize(h->stream));
        free(h->tpm_params); h->tpm_params = malloc(te.par_bytes);
        if (!te.fill(hs.data(), o, h->tpm_params)) h->tpm_kind = 0;      // QDu not diagonal: team kernels
    }
    if (h->tpm_kind) {
        if (const char* e = getenv("BMPC_TPM_CAPS")) sscanf(e, "%d,%d,%d,%d", &h->tpm_first_cap, &h->tpm_round_cap, &h->tpm_chunk, &h->tpm_cold_cap);   // tuning knob (tools/gpu_sweep.py)
        const TpmEntry& te = g_tpm_table[h->tpm_kind - 1];
        const size_t B = h->cfg.batch, nwarp = (B + 31) / 32;
        if (!h->tpm_W) BMPC_CUDA(cudaMalloc((void**)&h->tpm_W, sizeof(double) * nwarp * 32 * (size_t)te.slots));
        if (!h->tpm_rec) BMPC_CUDA(cudaMalloc((void**)&h->tpm_rec, sizeof(unsigned long long) * B * (size_t)(d.Np + 2)));
        BMPC_CUDA(cudaMemsetAsync(h->tpm_rec, 0, sizeof(unsigned long long) * B * (size_t)(d.Np + 2), h->stream));
    }
    // uminus1 default = uref for every instance (mpc.py:141); caller overrides through bmpc_update
    h->is_setup = true; h->cold = true; h->solved = false; h->pending = false; h->codes_valid = false;
    return BMPC_OK;
}



int bmpc_update(bmpc_handle* h, const double* x0, const double* uminus1, const double* xref, int xref_rows, int on_device) {
    if (!h) return BMPC_ERR_ARG;
    if (!h->is_setup) { h->err = "bmpc_update before bmpc_setup"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    if (h->pending) { int rc = finish_solve(h); if (rc) return rc; }
    const BmpcDims& d = h->d; size_t B = h->cfg.batch;
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (on_device == 2) {
        // borrowed device buffers: the solver kernels read them in place (no copy); the caller keeps them unchanged until the
        // results of the next solve have been retired (bmpc_output / bmpc_synchronize)
        if (x0) h->x0_cur = x0;
        if (uminus1) h->um1_cur = uminus1;
    } else {
        if (x0) { BMPC_CUDA(cudaMemcpyAsync(h->x0, x0, sizeof(double) * B * d.nx, kind, h->stream)); h->x0_cur = h->x0; }
        if (uminus1) { BMPC_CUDA(cudaMemcpyAsync(h->um1, uminus1, sizeof(double) * B * d.nu, kind, h->stream)); h->um1_cur = h->um1; }
    }
    h->I.x0 = h->x0_cur; h->I.um1 = h->um1_cur;
    if (xref) {
        if (xref_rows != 1 && xref_rows != d.Np + 1) { h->err = "xref_rows must be 1 or Np+1"; return BMPC_ERR_ARG; }
        h->xref_mode = xref_rows == 1 ? 0 : 1;
        BMPC_CUDA(cudaMemcpyAsync(h->xref, xref, sizeof(double) * B * (h->xref_mode ? d.NX : d.nx), kind, h->stream));
    }
    return BMPC_OK;
}

static bool use_fallback_team(const bmpc_handle* h, const int32_t* list, int count) {
    return h->team == 32 && h->fb_team && list != nullptr && count <= 4096;
}

static void launch_admm(bmpc_handle* h, const int32_t* list, int count, int niter, int do_prep) {
    const int cold = h->cold ? 1 : 0;
    if (use_fallback_team(h, list, count)) {
        // stragglers of a warp-team / fast-path handle: one warp per CTA (the lanes-own-rows ADMM has the shortest
        // dependent chain per iteration for these small shapes: measured 5 % on the random-instance bench vs a 128-thread CTA)
        k_admm<true><<<count, 32, admm_smem_doubles(h->d) * 8, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, niter, do_prep, cold, h->xref_mode, h->cfg.polish ? 1 : 0);
    } else if (h->team == 32 && h->tile_T == 0) {
        int grid = (count + h->wpb - 1) / h->wpb;
        k_admm<true><<<grid, h->wpb * 32, h->smem_admm, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, niter, do_prep, cold, h->xref_mode, h->cfg.polish ? 1 : 0);
    } else if (h->tile_T > 0) {
        // throughput tiles (8 instances share every K^-1 entry) while the batch fills the GPU twice over; straggler rounds
        // with few instances left use small tiles: more CTAs, a shorter dependent chain per iteration
        int k = (h->tile_T == 8 && count > 2 * h->sm_count * 8) ? 0 : ((h->tile_T >= 4 && count > 2 * h->sm_count * 2) ? 1 : 2);
        while ((8 >> k) > h->tile_T) k++;
        const int T = 8 >> k;
        h->tile_fn[k]<<<(count + T - 1) / T, h->tile_threads, bmpc_tile_smem_doubles(h->d, T) * 8, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, niter, do_prep, cold, h->xref_mode, h->cfg.polish ? 1 : 0);
    } else {
        k_admm<false><<<count, h->team, h->smem_admm, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, niter, do_prep, cold, h->xref_mode, h->cfg.polish ? 1 : 0);
    }
    h->stats.launches++;
}

static void launch_polish(bmpc_handle* h, const int32_t* list, int count, int32_t* next_list, int32_t* next_count) {
    const int steps = bmpc_polish_steps(h->cfg.pdas_steps, h->st.round);
    if (use_fallback_team(h, list, count)) {
        k_polish<false><<<count, h->fb_team, h->fb_smem_polish, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, h->fb_rmax, steps, next_list, next_count, h->I.u0, nullptr, nullptr, h->cfg.candidate_warm);
    } else if (h->team == 32) {
        int grid = (count + h->wpb - 1) / h->wpb;
        k_polish<true><<<grid, h->wpb * 32, h->smem_polish, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, h->rmax, steps, next_list, next_count, h->I.u0, nullptr, nullptr, h->cfg.candidate_warm);
    } else if (h->rmax_small > 0) {
        // two capacity tiers: the small one keeps several CTAs resident per SM (the polish is latency-bound); working
        // sets that outgrow it are listed and redone by the full-capacity launch right after (count read on the device)
        k_polish<false><<<count, h->team, h->smem_polish_small, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, h->rmax_small, steps, next_list, next_count, h->I.u0, nullptr, h->ovf, h->cfg.candidate_warm);
        k_polish<false><<<count, h->team, h->smem_polish, h->stream>>>(h->d, h->o, h->sys, h->I, h->ovf, count, h->rmax, steps, next_list, next_count, h->I.u0, next_count + 2, nullptr, h->cfg.candidate_warm);
        h->stats.launches++;
    } else {
        k_polish<false><<<count, h->team, h->smem_polish, h->stream>>>(h->d, h->o, h->sys, h->I, list, count, h->rmax, steps, next_list, next_count, h->I.u0, nullptr, nullptr, h->cfg.candidate_warm);
    }
    h->stats.launches++;
}

// ---- solve = rounds of [ADMM chunk -> polish]; the host only needs the count of unfinished instances after each
// round.  bmpc_solve enqueues the first round and returns; whoever needs results next (bmpc_output, ...) waits once,
// and only if stragglers remain runs further rounds.  In the warm closed loop this is ONE host sync per step.
static int enqueue_round(bmpc_handle* h) {
    auto& st = h->st;
    if (st.chunk > h->cfg.max_iter - st.total) st.chunk = h->cfg.max_iter - st.total;
    // fast path (thread-per-instance kernels, throughput-optimised) for the first round; the few stragglers are
    // latency-bound and go to the CTA-per-instance team kernels
    const bool tpi_ok = h->tpi_kind != 0;
    // (per-instance systems have no thread-per-instance ADMM — its matrices live in the constant bank —: their cold start and any
    // first_iters > 0 go through the team kernels, the warm polish-only round through k_tpi_pol with per-instance parameter blocks)
    const bool tpi = st.round == 0 && tpi_ok && h->cfg.polish && (h->tpi_view == nullptr || st.chunk == 0);
    // multi-input fast path: a warm solve starts with the Riccati polish alone (no ADMM, no prep)
    const bool tpm0 = st.round == 0 && h->tpm_kind != 0 && st.chunk == 0 && !h->cold;
    // straggler rounds read u_-1 from the snapshot the first round took: bmpc_output may already have queued the commit of
    // this solve's u0 into um1 (speculating that the first round finishes everything)
    h->cpar ^= 1;                                        // counters of this round: the half the previous round left zeroed
    int32_t* cnt = h->counts + BMPC_CNT * h->cpar;
    h->I.um1 = st.round > 0 ? h->um1_solved : h->um1_cur;
    h->I.x0 = h->x0_cur; h->I.u0_shadow = h->um1_alt;
    if (!tpi) {
        if (st.round == 0 && !tpm0) {
            const int B = h->cfg.batch;
            k_reset<<<(B + 255) / 256, 256, 0, h->stream>>>(h->I, B);
            h->stats.launches++;
            BMPC_CUDA(cudaMemcpyAsync(h->um1_solved, h->um1_cur, sizeof(double) * (size_t)B * h->d.nu, cudaMemcpyDefault, h->stream));   // (um1_cur may be mapped host memory)
        }
        BMPC_CUDA(cudaMemsetAsync(h->counts, 0, sizeof(int32_t) * 2 * BMPC_CNT, h->stream));     // this round's half and the next one's
    }
    h->spin_epoch = 0;
    if (tpi && st.chunk == 0) { h->epoch = (h->epoch % 1000000000) + 1; h->spin_epoch = h->epoch; ((volatile int32_t*)h->h_count)[BMPC_CNT_EPOCH] = 0; }
    // (a warm fast-path round times itself with the global timer and reports through mapped memory: no event records in the stream)
    if (!h->spin_epoch) BMPC_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if (tpi) {
        g_tpi_table[h->tpi_kind - 1].launch(h, st.list, st.count, st.chunk, st.nxt, h->ev[1]);
    } else if (tpm0) {
        BMPC_CUDA(cudaEventRecord(h->ev[1], h->stream));
        g_tpm_table[h->tpm_kind - 1].launch(h, nullptr, st.count, -1, h->tpm_first_cap, 1, st.nxt, 0);
    } else {
        // infeasible instances never pass the polish: from the third round on, look for OSQP's certificate
        const bool chk = st.total >= 25 && st.list != nullptr;
        if (chk) { k_snapshot<<<st.count, 128, 0, h->stream>>>(h->d, h->I, st.list, st.count, h->vprev, h->lprev); h->stats.launches++; }
        launch_admm(h, st.list, st.count, st.chunk, st.need_prep ? 1 : 0);
        st.need_prep = false;
        if (chk) {
            k_infeas<<<st.count, 128, sizeof(double) * (h->d.mc + h->d.nu + 2), h->stream>>>(h->d, h->o, h->sys, h->I, st.list, st.count, h->vprev, h->lprev, 1e-4, h->I.u0, cnt);
            h->stats.launches++;
        }
        BMPC_CUDA(cudaEventRecord(h->ev[1], h->stream));
        // stragglers of a fast-path shape: the Riccati polish (list mode) has ~3x lower latency than the team Schur polish
        if (h->cfg.polish && tpi_ok && st.total + st.chunk <= 200)
            g_tpi_table[h->tpi_kind - 1].launch_polish(h, st.list, st.count, st.nxt);
        // multi-input fast-path shapes: the Riccati polish (a refinement costs about one ADMM iteration of this shape) instead of
        // the Schur-form one, which takes over for the instances that are still open after 200 iterations (any working set)
        else if (h->cfg.polish && h->tpm_kind && st.total + st.chunk <= 200)           // (also the first attempt of a cold start: working sets from the iterate)
            // (first attempt of a cold solve: all-at-once updates; every straggler round: single exchanges — with the cold cap in a
            // cold solve, whose iterate is still rough: host study on 45 cold stragglers: cap 24 finishes all of them in their first
            // straggler round, cap 12 needs two or three)
            g_tpm_table[h->tpm_kind - 1].launch(h, st.list, st.count, 2, st.cold ? h->tpm_cold_cap : h->tpm_round_cap, 0, st.nxt, (st.cold && st.round == 0) ? 0 : 1);
        else if (h->cfg.polish) launch_polish(h, st.list, st.count, st.nxt, cnt);
        else { k_check_converged<<<(st.count + 255) / 256, 256, 0, h->stream>>>(h->I, st.list, st.count, h->cfg.eps_abs, h->cfg.eps_rel, st.nxt, cnt); h->stats.launches++; }
    }
    h->I.um1 = h->um1_cur;
    if (!h->spin_epoch) BMPC_CUDA(cudaEventRecord(h->ev[2], h->stream));
    if (!h->spin_epoch) BMPC_CUDA(cudaMemcpyAsync(h->h_count, cnt, sizeof(int32_t) * BMPC_CNT, cudaMemcpyDeviceToHost, h->stream));
    BMPC_CUDA(cudaGetLastError());
    return BMPC_OK;
}

// waits for the round in flight; returns 1 in *more if another round was enqueued (stragglers), 0 if the solve is complete
static int retire_round(bmpc_handle* h, int* more) {
    auto& st = h->st;
    float a = 0.f, p = 0.f;
    bool spun = false;
    if (h->spin_epoch && !h->sync_after_round) {
        // the kernel's last warp raises the epoch word in mapped host memory when every counter is final: a few microseconds
        // after the kernel ends, instead of a D2H copy + cudaStreamSynchronize (the round trip sits in every step of a control loop)
        volatile int32_t* flag = (volatile int32_t*)h->h_count + BMPC_CNT_EPOCH;
        for (long spins = 0; *flag != h->spin_epoch; spins++)
            if (spins > 200000000L) break;                              // ~ seconds: fall back to the stream
        spun = *flag == h->spin_epoch;
    }
    if (!spun) {
        BMPC_CUDA(cudaStreamSynchronize(h->stream));
        if (h->spin_epoch && ((volatile int32_t*)h->h_count)[BMPC_CNT_EPOCH] != h->spin_epoch) { h->err = "fast-path kernel did not report"; return BMPC_ERR_CUDA; }
        if (!h->spin_epoch) { cudaEventElapsedTime(&a, h->ev[0], h->ev[1]); cudaEventElapsedTime(&p, h->ev[1], h->ev[2]); }
    }
    {   // a fast-path polish launch timed itself (global timer): more faithful than events, which also see launch gaps
        unsigned long long t0n, t1; memcpy(&t0n, h->h_count + BMPC_CNT_T0, 8); memcpy(&t1, h->h_count + BMPC_CNT_T1, 8);
        if (t1 != 0ull && t0n != 0ull && t1 > ~t0n) p = (float)((double)(t1 - ~t0n) * 1e-6);
    }
    h->stats.ms_admm += a; h->stats.ms_polish += p;
    h->stats.admm_iters += (int64_t)st.count * st.chunk;
    st.total += st.chunk; st.round++;
    h->cold = false;
    st.count = h->h_count[0];
    h->stats.polish_steps += h->h_count[1];
    h->stats.infeasible += h->h_count[3];
    st.tight += h->h_count[BMPC_CNT_TIGHT];
    if (spun && h->h_count[BMPC_CNT_TIGHT + 1] == 1) h->g_done_epoch = h->g_epoch;
    st.list = st.nxt; int32_t* tmp = st.cur; st.cur = st.nxt; st.nxt = tmp;
    // polish mode: cumulative first_iters, 25, 50, 100, ...; pure ADMM: OSQP's check_termination = 25
    st.chunk = h->cfg.polish ? (st.total < 25 ? 25 - st.total : st.total) : 25;
    if (h->tpm_kind && st.total == 0 && h->cfg.polish) st.chunk = h->tpm_chunk;        // first straggler round of a multi-input fast-path solve
    if (st.count > 0 && st.total < h->cfg.max_iter) { *more = 1; return enqueue_round(h); }
    *more = 0;
    const int B = h->cfg.batch;
    if (!h->cfg.polish) {
        // pure-ADMM mode: every instance gets its status from OSQP's criterion on its last residuals
        k_finalize<<<(B + 127) / 128, 128, 0, h->stream>>>(h->d, h->o, h->sys, h->I, nullptr, B, h->cfg.eps_abs, h->cfg.eps_rel, h->I.u0);
        h->stats.launches++;
    } else if (st.count > 0) {
        k_finalize<<<(st.count + 127) / 128, 128, 0, h->stream>>>(h->d, h->o, h->sys, h->I, st.list, st.count, h->cfg.eps_abs, h->cfg.eps_rel, h->I.u0);
        h->stats.launches++;
    }
    BMPC_CUDA(cudaGetLastError());
    h->stats.rounds = st.round; h->stats.unsolved = h->cfg.polish ? st.count + st.tight : -1;   // not KKT-verified: status 2 or -2
    h->pending = false;
    return BMPC_OK;
}

static int finish_solve(bmpc_handle* h) {
    while (h->pending) { int more = 0; int rc = retire_round(h, &more); if (rc) return rc; }
    return BMPC_OK;
}

int bmpc_solve(bmpc_handle* h) {
    if (!h) return BMPC_ERR_ARG;
    if (!h->is_setup) { h->err = "bmpc_solve before bmpc_setup"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    if (h->pending) { int rc = finish_solve(h); if (rc) return rc; }
    const int B = h->cfg.batch;
    memset(&h->stats, 0, sizeof(h->stats));
    auto& st = h->st;
    st.list = nullptr; st.count = B; st.cur = h->listA; st.nxt = h->listB;
    st.total = 0; st.round = 0; st.need_prep = true; st.tight = 0; st.cold = h->cold;
    if (h->gflags) h->g_epoch++;                     // every rank solves in lockstep: the arrival epoch of this step
    // first round: NO ADMM iterations on a warm fast-path solve (the previous solution's working sets are the best first guess
    // the active-set polish can get: measured 0 vs 1..10 iterations), 10 on the team kernels; first_iters > 0 overrides
    const bool fast = h->tpi_kind != 0 || (h->tpm_kind != 0 && !h->cold);
    st.chunk = h->cfg.polish ? (h->cfg.first_iters > 0 ? h->cfg.first_iters : (fast ? 0 : 10)) : 25;
    // a cold start has no active-set guess to refresh: 25 iterations at once on the fast path, so that the first polish
    // usually verifies and the whole batch does not take the straggler route
    if (h->cfg.polish && h->cfg.first_iters <= 0 && fast && h->cold) st.chunk = 25;
    // team / tile kernels, cold: 10 iterations from the free response leave working sets of hundreds of rows for the Schur polish
    // (round 1 measured 180 ms launches on the MIMO shape): iterate longer before the first attempt
    // (multi-input fast-path shapes: the Riccati polish takes the first attempt there and copes with the iterate after 25)
    if (h->cfg.polish && h->cfg.first_iters <= 0 && !fast && h->cold) st.chunk = h->cfg.cold_iters > 0 ? h->cfg.cold_iters : (h->tpm_kind ? 25 : 50);
    if (h->cfg.polish && fast && h->tpi_view && !h->cold && h->cfg.first_iters > 0) st.chunk = h->cfg.first_iters;
    if (st.chunk > h->cfg.max_iter) st.chunk = h->cfg.max_iter;
    int rc = enqueue_round(h);
    if (rc) return rc;
    h->pending = true; h->solved = true; h->committed = false;
    return BMPC_OK;
}

int bmpc_output(bmpc_handle* h, double* u0, int32_t* status, int commit_uminus1, int on_device) {
    if (!h) return BMPC_ERR_ARG;
    if (!h->solved) { h->err = "bmpc_output before bmpc_solve"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    const BmpcDims& d = h->d; size_t B = h->cfg.batch;
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const bool want_u = u0 && u0 != h->I.u0;
    auto copy_u = [&]() -> int {
        if (want_u) BMPC_CUDA(cudaMemcpyAsync(u0, h->I.u0, sizeof(double) * B * d.nu, kind, h->stream));
        return BMPC_OK;
    };
    bool u_copied = false;
    if (h->pending) {
        // speculate that the round in flight finishes everything (the common case): queue the result copy behind it so that a
        // single wait covers the solve and the read-back; redone if stragglers needed more rounds (they rewrite u0)
        if (!on_device) { int rc = copy_u(); if (rc) return rc; u_copied = true; }
        h->sync_after_round = !on_device && want_u;
        int more = 0; int rc = retire_round(h, &more); h->sync_after_round = false; if (rc) return rc;
        if (more) { u_copied = false; rc = finish_solve(h); if (rc) return rc; }
        else if (h->stats.launches && (h->st.count > 0 || !h->cfg.polish)) u_copied = false;   // k_finalize rewrote u0/status
    }
    if (commit_uminus1 && !h->committed) {
        // every kernel that publishes u0 also wrote it into the shadow buffer: committing it as the next u_-1 (mpc.py:330) is a
        // pointer swap, no copy
        double* t = h->um1; h->um1 = h->um1_alt; h->um1_alt = t;
        h->um1_cur = h->um1; h->I.um1 = h->um1; h->I.u0_shadow = h->um1_alt;
        h->committed = true;
    }
    // every instance KKT-verified and none certified infeasible: the status array is all BMPC_SOLVED, no need to fetch it
    const bool all_solved = h->cfg.polish && h->stats.unsolved == 0 && h->stats.infeasible == 0;
    bool need_sync = false;
    if (!u_copied && want_u) { int rc = copy_u(); if (rc) return rc; need_sync = !on_device; }
    if (status) {
        if (all_solved && !on_device) {
            // (a closed loop asks with the same array every step: filled once while it stays all-solved)
            if (h->status_filled != status) { for (size_t i = 0; i < B; i++) status[i] = BMPC_SOLVED; h->status_filled = status; }
        }
        else { BMPC_CUDA(cudaMemcpyAsync(status, h->I.status, sizeof(int32_t) * B, kind, h->stream)); need_sync = !on_device; h->status_filled = nullptr; }
    }
    if (need_sync) BMPC_CUDA(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_get_sequences(bmpc_handle* h, double* u_seq, double* x_seq, double* eps_seq, double* obj_val, int32_t* iters) {
    if (!h) return BMPC_ERR_ARG;
    if (!h->solved) { h->err = "bmpc_get_sequences before bmpc_solve"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    if (h->pending) { int rc = finish_solve(h); if (rc) return rc; }
    const BmpcDims& d = h->d; size_t B = h->cfg.batch;
    if (x_seq || eps_seq || obj_val) {
        if (!h->seq_x) {
            BMPC_CUDA(cudaMalloc((void**)&h->seq_x, sizeof(double) * B * d.NX));
            BMPC_CUDA(cudaMalloc((void**)&h->seq_e, sizeof(double) * B * d.NX));
            BMPC_CUDA(cudaMalloc((void**)&h->seq_obj, sizeof(double) * B));
        }
        int wpb = 4;
        size_t sm = (size_t)wpb * (d.NU + d.NX + d.nu + 1) * sizeof(double);
        BMPC_CUDA(cudaFuncSetAttribute(k_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_sequences<<<(int)((B + wpb - 1) / wpb), wpb * 32, sm, h->stream>>>(d, h->o, h->sys, h->I, (int)B, h->xref_mode, h->seq_x, h->seq_e, h->seq_obj);
        BMPC_CUDA(cudaGetLastError());
        if (x_seq) BMPC_CUDA(cudaMemcpyAsync(x_seq, h->seq_x, sizeof(double) * B * d.NX, cudaMemcpyDeviceToHost, h->stream));
        if (eps_seq) BMPC_CUDA(cudaMemcpyAsync(eps_seq, h->seq_e, sizeof(double) * B * d.NX, cudaMemcpyDeviceToHost, h->stream));
        if (obj_val) BMPC_CUDA(cudaMemcpyAsync(obj_val, h->seq_obj, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
    }
    if (u_seq) BMPC_CUDA(cudaMemcpyAsync(u_seq, h->I.Us, sizeof(double) * B * d.NU, cudaMemcpyDeviceToHost, h->stream));
    if (iters) BMPC_CUDA(cudaMemcpyAsync(iters, h->I.iters, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, h->stream));
    BMPC_CUDA(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_get_stats(bmpc_handle* h, bmpc_stats* out) {
    if (!h || !out) return BMPC_ERR_ARG;
    if (h->pending) { cudaSetDevice(h->cfg.device); int rc = finish_solve(h); if (rc) return rc; }
    *out = h->stats;    // host-side counters only
    return BMPC_OK;
}

int bmpc_get_sys(bmpc_handle* h, const char* name, double* out, int capacity) {
    if (!h || !name || !out) return BMPC_ERR_ARG;
    if (!h->is_setup) { h->err = "bmpc_get_sys before bmpc_setup"; return BMPC_ERR_STATE; }
    BMPC_CUDA(cudaSetDevice(h->cfg.device));
    const BmpcDims& d = h->d; const BmpcSysOff& o = h->o;
    struct { const char* n; int off; int cnt; } tab[] = {
        {"Bcal", o.Bcal, d.NX * d.NU}, {"Acal", o.Acal, d.NX * d.nx}, {"H", o.H, d.NU * d.NU}, {"Hinv", o.Hinv, d.NU * d.NU},
        {"K", o.K, d.NU * d.NU}, {"Kinv", o.Kinv, d.NU * d.NU}, {"M", o.M, d.mc * d.mc}, {"AHinv", o.AHinv, d.mc * d.NU},
        {"Gx0", o.Gx0, d.NU * d.nx}, {"Gref", o.Gref, d.NU * d.nx}, {"g0", o.g0, d.NU}, {"lo0", o.lo0, d.mc},
        {"hi0", o.hi0, d.mc}, {"rho", o.rho, d.mc}, {"scal", o.scal, BMPC_S_COUNT}};
    for (auto& e : tab) if (strcmp(e.n, name) == 0) {
        if (capacity < e.cnt) { h->err = "bmpc_get_sys: buffer too small"; return BMPC_ERR_ARG; }
        BMPC_CUDA(cudaMemcpyAsync(out, h->sys + e.off, sizeof(double) * e.cnt, cudaMemcpyDeviceToHost, h->stream));
        BMPC_CUDA(cudaStreamSynchronize(h->stream));
        return e.cnt;
    }
    h->err = std::string("bmpc_get_sys: unknown array ") + name;
    return BMPC_ERR_ARG;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// batched linear state estimator (kalman.py:109-134): one thread per instance, shared matrices in shared memory
__global__ void k_est(int nx, int nu, int ny, int B, const double* __restrict__ mats, double* __restrict__ x, double* __restrict__ y,
                      const double* __restrict__ in, int mode) {
    extern __shared__ double sm[];
    const int nm = nx * nx + nx * nu + ny * nx + nx * ny;
    for (int i = threadIdx.x; i < nm; i += blockDim.x) sm[i] = mats[i];
    __syncthreads();
    const double *A = sm, *Bm = A + nx * nx, *C = Bm + nx * nu, *L = C + ny * nx;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* xb = x + (size_t)b * nx; double* yb = y + (size_t)b * ny;
    double xn[32];
    if (mode == 0) {            // predict: x <- A x + B u ; y <- C x
        const double* u = in + (size_t)b * nu;
        for (int i = 0; i < nx; i++) {
            double acc = 0.0;
            for (int j = 0; j < nx; j++) acc += A[i * nx + j] * xb[j];
            for (int j = 0; j < nu; j++) acc += Bm[i * nu + j] * u[j];
            xn[i] = acc;
        }
        for (int i = 0; i < nx; i++) xb[i] = xn[i];
        for (int i = 0; i < ny; i++) { double acc = 0.0; for (int j = 0; j < nx; j++) acc += C[i * nx + j] * xn[j]; yb[i] = acc; }
    } else {                    // update: x <- x + L (y_meas - y)      (y is NOT refreshed, as in the reference)
        const double* ym = in + (size_t)b * ny;
        for (int i = 0; i < nx; i++) {
            double acc = xb[i];
            for (int j = 0; j < ny; j++) acc += L[i * ny + j] * (ym[j] - yb[j]);
            xn[i] = acc;
        }
        for (int i = 0; i < nx; i++) xb[i] = xn[i];
    }
}

struct bmpc_estimator {
    int nx, nu, ny, B, device;
    double *mats = nullptr, *x = nullptr, *y = nullptr, *in = nullptr;
    cudaStream_t stream = nullptr, own = nullptr;
    bmpc_handle* mpc = nullptr;        // attached controller: same stream, and its deferred solve is retired before its u0 is read
};

extern "C" {

int bmpc_est_create(int32_t nx, int32_t nu, int32_t ny, int32_t batch, int32_t device, const double* A, const double* B,
                    const double* C, const double* L, const double* x0, bmpc_estimator** out) {
    if (!out || !A || !B || !C || !L || !x0 || nx < 1 || nx > 32 || nu < 1 || ny < 1 || batch < 1) { g_create_err = "bmpc_est_create: bad argument (1 <= nx <= 32)"; return BMPC_ERR_ARG; }
    *out = nullptr;
    if (bmpc_device_count() <= device) { g_create_err = "no CUDA device visible: libbmpc has no CPU fallback"; return BMPC_ERR_NO_DEVICE; }
    cudaSetDevice(device);
    bmpc_estimator* e = new bmpc_estimator();
    e->nx = nx; e->nu = nu; e->ny = ny; e->B = batch; e->device = device;
    const size_t nm = (size_t)nx * nx + (size_t)nx * nu + (size_t)ny * nx + (size_t)nx * ny;
    const int mx = nu > ny ? nu : ny;
    bool ok = cudaMalloc((void**)&e->mats, nm * 8) == cudaSuccess && cudaMalloc((void**)&e->x, (size_t)batch * nx * 8) == cudaSuccess &&
              cudaMalloc((void**)&e->y, (size_t)batch * ny * 8) == cudaSuccess && cudaMalloc((void**)&e->in, (size_t)batch * mx * 8) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->own, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { g_create_err = "bmpc_est_create: allocation failed"; bmpc_est_destroy(e); return BMPC_ERR_CUDA; }
    e->stream = e->own;
    std::vector<double> m(nm);
    memcpy(m.data(), A, sizeof(double) * nx * nx); memcpy(m.data() + nx * nx, B, sizeof(double) * nx * nu);
    memcpy(m.data() + nx * nx + nx * nu, C, sizeof(double) * ny * nx); memcpy(m.data() + nx * nx + nx * nu + ny * nx, L, sizeof(double) * nx * ny);
    cudaMemcpyAsync(e->mats, m.data(), nm * 8, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(e->x, x0, (size_t)batch * nx * 8, cudaMemcpyHostToDevice, e->stream);
    // y = C x0 (kalman.py:113): a predict with A = I, B = 0 would do; compute on the host instead (setup only)
    std::vector<double> y0((size_t)batch * ny);
    for (int b = 0; b < batch; b++) for (int i = 0; i < ny; i++) { double acc = 0; for (int j = 0; j < nx; j++) acc += C[i * nx + j] * x0[(size_t)b * nx + j]; y0[(size_t)b * ny + i] = acc; }
    cudaMemcpyAsync(e->y, y0.data(), y0.size() * 8, cudaMemcpyHostToDevice, e->stream);
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) { g_create_err = "bmpc_est_create: upload failed"; bmpc_est_destroy(e); return BMPC_ERR_CUDA; }
    *out = e;
    return BMPC_OK;
}

void bmpc_est_destroy(bmpc_estimator* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->mats); cudaFree(e->x); cudaFree(e->y); cudaFree(e->in);
    if (e->own) cudaStreamDestroy(e->own);
    delete e;
}

static int est_run(bmpc_estimator* e, const double* in, int n, int on_device, int mode) {
    if (!e || !in) return BMPC_ERR_ARG;
    if (cudaSetDevice(e->device) != cudaSuccess) return BMPC_ERR_CUDA;
    if (e->mpc) {
        // on-device chain estimator <-> controller: one stream orders the kernels; a solve whose first round is still in flight may
        // need straggler rounds (host driven) before its u0 is final, so it is retired here
        if (on_device && e->mpc->pending) { int rc = finish_solve(e->mpc); if (rc) return rc; }
        e->stream = e->mpc->stream;
    }
    const double* src = in;
    if (!on_device) {
        if (cudaMemcpyAsync(e->in, in, (size_t)e->B * n * 8, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) return BMPC_ERR_CUDA;
        src = e->in;
    }
    const size_t nm = (size_t)e->nx * e->nx + (size_t)e->nx * e->nu + (size_t)e->ny * e->nx + (size_t)e->nx * e->ny;
    k_est<<<(e->B + 127) / 128, 128, nm * 8, e->stream>>>(e->nx, e->nu, e->ny, e->B, e->mats, e->x, e->y, src, mode);
    return cudaGetLastError() == cudaSuccess ? BMPC_OK : BMPC_ERR_CUDA;
}
int bmpc_est_predict(bmpc_estimator* e, const double* u, int on_device) { return est_run(e, u, e ? e->nu : 0, on_device, 0); }
int bmpc_est_update(bmpc_estimator* e, const double* y_meas, int on_device) { return est_run(e, y_meas, e ? e->ny : 0, on_device, 1); }
int bmpc_est_get(bmpc_estimator* e, double* x, double* y) {
    if (!e) return BMPC_ERR_ARG;
    cudaSetDevice(e->device);
    if (x) cudaMemcpyAsync(x, e->x, (size_t)e->B * e->nx * 8, cudaMemcpyDeviceToHost, e->stream);
    if (y) cudaMemcpyAsync(y, e->y, (size_t)e->B * e->ny * 8, cudaMemcpyDeviceToHost, e->stream);
    return cudaStreamSynchronize(e->stream) == cudaSuccess ? BMPC_OK : BMPC_ERR_CUDA;
}
double* bmpc_est_state_ptr(bmpc_estimator* e) { return e ? e->x : nullptr; }
int bmpc_est_set_stream(bmpc_estimator* e, void* s) { if (!e) return BMPC_ERR_ARG; e->stream = s ? (cudaStream_t)s : e->own; return BMPC_OK; }
int bmpc_est_attach(bmpc_estimator* e, bmpc_handle* h) {
    if (!e) return BMPC_ERR_ARG;
    if (h && h->cfg.device != e->device) return BMPC_ERR_ARG;
    cudaSetDevice(e->device); cudaStreamSynchronize(e->stream);
    e->mpc = h; e->stream = h ? h->stream : e->own;
    return BMPC_OK;
}

}  // extern "C"
