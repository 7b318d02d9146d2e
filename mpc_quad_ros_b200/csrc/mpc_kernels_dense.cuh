// mpc_kernels_dense.cuh — K2b: condensed (dense) box-QP solver, one CTA per OCP, for the OCPs whose warm-started
// active-set rounds did not settle in the screening kernel (qmpc_ipm_kernel with a hard list) and for cold starts.
//
// Why: a Riccati sweep is 20 strictly sequential stages of warp-serial work; the ~5 % of OCPs that need the
// interior-point method kept one warp busy for 1.3-2.7 ms while the rest of the GPU idled (profiles/r01_timeline.txt).
// Eliminating the states gives  min 1/2 du' H du + f' du,  lb <= ubar + du <= ub  with H = Rbar + sum_k G_k' Q_k G_k
// (G_k = impulse responses), E = 4N <= 88 unknowns.  Everything is then wide: condensing is 4x4-tile rank-13 updates
// held in registers (one thread per tile of the lower triangle), every factorisation is a blocked Cholesky over the
// same tiles (2 barriers per block column), gradients/multipliers are dense mat-vecs; only the two triangular solves
// per right-hand side stay serial (one warp, shuffles).  Same algorithm as the Riccati path: Mehrotra IPM from the
// box centre with gradient-scaled multipliers, hand-over to exact primal-dual active-set rounds, full step.
#pragma once
#include "mpc_kernels_v2.cuh"

namespace qmpc {

constexpr int DN_THREADS = 256;
constexpr int DN_MAX_N = 22;        // N(N+1)/2 tiles <= 256 threads and E = 4N <= 96 (three rows per lane in the solves)

template <typename real>
struct DenseArgs {
    IpmArgs<real> b;
    const int* hard_list;           // OCP indices to solve, or null = all of 0..B-1
    const int* hard_count;          // device counter written by the screening kernel (read here, no host sync)
};

// shared-memory carve-up (reals)
struct DenseLayout {
    int E, T, GS, Ht, Lt, G, ev, sml, vec, total;
};
__host__ __device__ inline DenseLayout dense_layout(int N)
{
    DenseLayout L;
    L.E = 4 * N; L.T = N * (N + 1) / 2; L.GS = L.E + 4;
    L.Ht = 0; L.Lt = L.T * 16; L.G = 2 * L.T * 16; L.ev = L.G + 13 * L.GS; L.sml = L.ev + 16; L.vec = L.sml + 48;
    L.total = L.vec + 14 * L.E;
    return L;
}
constexpr int DN_NVEC = 14;

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // tile (i, j), j <= i

template <typename real>
struct DenseCtx {
    const IpmArgs<real>& a;
    int tid, lane, N, E, T, GS, ti, tj;
    real *Ht, *Lt, *G, *ev;
    real *f, *ubar, *ucur, *tl, *tu, *ll, *lu, *cl, *cu, *ua, *usol, *rt, *dR, *tv, *fx, *fv, *grad;

    // tv = H x  (thread per row; H symmetric, stored as 4x4 tiles of the lower triangle)
    __device__ __forceinline__ void matvec(const real* x)
    {
        if (tid < E) {
            const int I = tid >> 2, ar = tid & 3;
            real s0 = 0, s1 = 0;
            for (int J = 0; J <= I; ++J) {
                const real* p = Ht + tri(I, J) * 16 + ar * 4;
                real h0, h1, h2, h3, x0, x1, x2, x3;
                ld2(p, h0, h1); ld2(p + 2, h2, h3);
                ld2(x + 4 * J, x0, x1); ld2(x + 4 * J + 2, x2, x3);
                s0 += h0 * x0 + h2 * x2; s1 += h1 * x1 + h3 * x3;
            }
            for (int J = I + 1; J < N; ++J) {
                const real* p = Ht + tri(J, I) * 16 + ar;
                real x0, x1, x2, x3;
                ld2(x + 4 * J, x0, x1); ld2(x + 4 * J + 2, x2, x3);
                s0 += p[0] * x0 + p[8] * x2; s1 += p[4] * x1 + p[12] * x3;
            }
            tv[tid] = s0 + s1;
        }
    }

    // Lt = chol(K): K = H + diag(dR) (IPM) or H with the inputs flagged in fx replaced by identity rows/columns.
    // One thread per 4x4 tile; right-looking over block columns, the current panel goes through shared memory.
    __device__ __forceinline__ void factor(const bool fixed)
    {
        real acc[16];
        if (tid < T) {
            const real* p = Ht + tid * 16;
#pragma unroll
            for (int t = 0; t < 16; t += 2) ld2(p + t, acc[t], acc[t + 1]);
            real ki[4], kj[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ki[q] = (fixed && fx[4 * ti + q] != real(0)) ? real(0) : real(1);
                kj[q] = (fixed && fx[4 * tj + q] != real(0)) ? real(0) : real(1);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q * 4 + r] *= ki[q] * kj[r];
            if (ti == tj) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q * 5] = ki[q] != real(0) ? acc[q * 5] + (fixed ? real(0) : dR[4 * ti + q]) : real(1);
            }
        }
        for (int K = 0; K < N; ++K) {
            if (tid < T && ti == K && tj == K) {
                Chol4<real> L;
                L.factor(acc);
                L.store(Lt + tid * 16);
            }
            __syncthreads();
            if (tid < T && tj == K && ti > K) {
                Chol4<real> L;
                L.load(Lt + tri(K, K) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    real z[4];
                    L.fsolve(acc + q * 4, z);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[q * 4 + r] = z[r];
                }
                real* o = Lt + tid * 16;
#pragma unroll
                for (int t = 0; t < 16; ++t) o[t] = acc[t];
            }
            __syncthreads();
            if (tid < T && tj > K) {
                real li[16], lj[16];
                const real* pi = Lt + tri(ti, K) * 16;
                const real* pj = Lt + tri(tj, K) * 16;
#pragma unroll
                for (int t = 0; t < 16; t += 2) { ld2(pi + t, li[t], li[t + 1]); ld2(pj + t, lj[t], lj[t + 1]); }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        real s = acc[q * 4 + r];
#pragma unroll
                        for (int c = 0; c < 4; ++c) s = fma(-li[q * 4 + c], lj[r * 4 + c], s);
                        acc[q * 4 + r] = s;
                    }
            }
        }
        __syncthreads();
    }

    // warp 0: dst = K^-1 rhs with the factor in Lt (rows lane, lane+32, lane+64 per lane)
    __device__ __forceinline__ void solve(const real* rhs, real* dst)
    {
        real r[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { const int e = lane + 32 * m; r[m] = e < E ? rhs[e] : real(0); }
        for (int K = 0; K < N; ++K) {
            const int m0 = K >> 3, l0 = (4 * K) & 31;
            const real rs = m0 == 0 ? r[0] : (m0 == 1 ? r[1] : r[2]);
            real rk[4], yk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rk[q] = __shfl_sync(FULL, rs, l0 + q);
            Chol4<real> L;
            L.load(Lt + tri(K, K) * 16);
            L.fsolve(rk, yk);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int e = lane + 32 * m, I = e >> 2;
                if (e < E) {
                    if (I == K) r[m] = sel4(yk, e & 3);
                    else if (I > K) {
                        const real* p = Lt + tri(I, K) * 16 + (e & 3) * 4;
                        real l0_, l1_, l2_, l3_;
                        ld2(p, l0_, l1_); ld2(p + 2, l2_, l3_);
                        r[m] -= l0_ * yk[0] + l1_ * yk[1] + l2_ * yk[2] + l3_ * yk[3];
                    }
                }
            }
        }
        for (int K = N - 1; K >= 0; --K) {
            const int m0 = K >> 3, l0 = (4 * K) & 31;
            const real rs = m0 == 0 ? r[0] : (m0 == 1 ? r[1] : r[2]);
            real rk[4], xk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rk[q] = -__shfl_sync(FULL, rs, l0 + q);
            Chol4<real> L;
            L.load(Lt + tri(K, K) * 16);
            L.bsolve_neg(rk, xk);                  // Lam^T x = y
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int e = lane + 32 * m, I = e >> 2;
                if (e < E) {
                    if (I == K) r[m] = sel4(xk, e & 3);
                    else if (I < K) {
                        const real* p = Lt + tri(K, I) * 16 + (e & 3);
                        r[m] -= p[0] * xk[0] + p[4] * xk[1] + p[8] * xk[2] + p[12] * xk[3];
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) { const int e = lane + 32 * m; if (e < E) dst[e] = r[m]; }
        __syncwarp();
    }
};

#ifdef QMPC_DENSE_PROF
#define DPROF_DECL long long pf_t0 = clock64(), pf_last = pf_t0, pf_acc[6] = {0, 0, 0, 0, 0, 0}; int pf_n[6] = {0, 0, 0, 0, 0, 0}
#define DPROF(slot) { const long long pf_now = clock64(); pf_acc[slot] += pf_now - pf_last; ++pf_n[slot]; pf_last = pf_now; }
#else
#define DPROF_DECL
#define DPROF(slot)
#endif

template <typename real>
__global__ void __launch_bounds__(DN_THREADS, 2) qmpc_dense_kernel(DenseArgs<real> da)
{
    QMPC_DYN_SMEM(smem_raw);
    const IpmArgs<real>& a = da.b;
    real* sm = reinterpret_cast<real*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, E = 4 * N;
    const DenseLayout lay = dense_layout(N);
    const int T = lay.T, GS = lay.GS;
    DenseCtx<real> c{a};
    c.tid = tid; c.lane = lane; c.N = N; c.E = E; c.T = T; c.GS = GS;
    c.Ht = sm + lay.Ht; c.Lt = sm + lay.Lt; c.G = sm + lay.G; c.ev = sm + lay.ev;
    real* small = sm + lay.sml;                 // wv(16) xp(16) + control words
    int* ctl = reinterpret_cast<int*>(small + 32);
    real* v = sm + lay.vec;
    c.f = v; c.ubar = v + E; c.ucur = v + 2 * E; c.tl = v + 3 * E; c.tu = v + 4 * E; c.ll = v + 5 * E; c.lu = v + 6 * E;
    c.cl = v + 7 * E; c.cu = v + 8 * E; c.ua = v + 9 * E; c.usol = v + 10 * E; c.rt = v + 11 * E; c.dR = v + 12 * E;
    c.tv = v + 13 * E;
    c.fx = c.cl; c.fv = c.cu; c.grad = c.ua;
    {   // tile owned by this thread
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= tid) ++i;
        c.ti = i; c.tj = tid - i * (i + 1) / 2;
    }
    const int ti = c.ti, tj = c.tj;
    const int count = da.hard_list ? *da.hard_count : a.B;
    const real lb = a.lb, ub = a.ub;
    enum { T_FIXED, T_ADJ, T_PRED, T_CORR, T_GRAD, T_DONE };

    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const int ocp = da.hard_list ? da.hard_list[item] : item;
        if (a.timeline && tid == 0) a.timeline[2 * ocp] = global_ns();
        const real* Wv = a.W + (size_t)ocp * N * WT;
        const double* x0 = a.x0 + (size_t)ocp * NX;
        const double* yref = a.yref + (size_t)ocp * N * NY;
        const double* yref_e = a.yref_e + (size_t)ocp * NX;
        double* xit = a.xit + (size_t)ocp * (N + 1) * NX;
        double* uit = a.uit + (size_t)ocp * N * NU;
        unsigned char* actset = a.act + (size_t)ocp * E;

        DPROF_DECL;
        // ---- condensing
        for (int idx = tid; idx < 13 * GS; idx += DN_THREADS) c.G[idx] = 0;
        __syncthreads();
        if (tid < NX) c.G[tid * GS + E] = real(x0[tid] - xit[tid]);
        if (tid < E) {
            const real ub_ = real(uit[tid]);
            c.ubar[tid] = ub_;
            c.f[tid] = a.Rd[tid & 3] * (ub_ - real(yref[(size_t)(tid >> 2) * NY + NX + (tid & 3)]));
        }
        real acc[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = 0;
        real fc = 0;
        __syncthreads();
        for (int k = 0; k < N; ++k) {
            const real* tile = Wv + (size_t)k * WT;
            const bool colact = (tid < E && (tid >> 2) <= k) || tid == E;
            real gn[NX];
            if (colact) {
                if (tid < E && (tid >> 2) == k) {
#pragma unroll
                    for (int r = 0; r < NX; ++r) gn[r] = __ldg(tile + r * 16 + (tid & 3));
                } else {
                    real g[NX];
#pragma unroll
                    for (int s = 0; s < NX; ++s) g[s] = c.G[s * GS + tid];
#pragma unroll
                    for (int r = 0; r < NX; ++r) {
                        real s0 = r < 3 ? g[r] : real(0), s1 = 0;
#pragma unroll
                        for (int s = 0; s < 10; s += 2) {
                            real t0, t1;
                            ldg2(tile + r * 16 + 4 + s, t0, t1);
                            s0 += t0 * g[3 + s]; s1 += t1 * g[4 + s];
                        }
                        gn[r] = s0 + s1;
                    }
                    if (tid == E) {
#pragma unroll
                        for (int r = 0; r < NX; ++r) gn[r] += __ldg(tile + r * 16 + 14);
                    }
                }
            }
            __syncthreads();                    // every read of the previous G is done
            if (colact) {
#pragma unroll
                for (int r = 0; r < NX; ++r) c.G[r * GS + tid] = gn[r];
                if (tid == E) {
#pragma unroll
                    for (int r = 0; r < NX; ++r) {
                        const real w = (k + 1 < N) ? a.Qd[r] : a.QNd[r];
                        const double ref = (k + 1 < N) ? yref[(size_t)(k + 1) * NY + r] : yref_e[r];
                        c.ev[r] = w * (gn[r] + real(xit[(size_t)(k + 1) * NX + r] - ref));
                    }
                }
            }
            __syncthreads();
            if (tid < T && ti <= k) {           // H tile += G_i' Q G_j
#pragma unroll
                for (int r = 0; r < NX; ++r) {
                    const real w = (k + 1 < N) ? a.Qd[r] : a.QNd[r];
                    real gi[4], gj[4];
                    ld2(c.G + r * GS + 4 * ti, gi[0], gi[1]); ld2(c.G + r * GS + 4 * ti + 2, gi[2], gi[3]);
                    ld2(c.G + r * GS + 4 * tj, gj[0], gj[1]); ld2(c.G + r * GS + 4 * tj + 2, gj[2], gj[3]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) gj[q] *= w;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[p * 4 + q] = fma(gi[p], gj[q], acc[p * 4 + q]);
                }
            }
            if (tid < E && (tid >> 2) <= k) {
#pragma unroll
                for (int r = 0; r < NX; ++r) fc = fma(gn[r], c.ev[r], fc);
            }
        }
        if (tid < T) {
            if (ti == tj) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q * 5] += a.Rd[q];
            }
            real* o = c.Ht + tid * 16;
#pragma unroll
            for (int t = 0; t < 16; ++t) o[t] = acc[t];
        }
        if (tid < E) c.f[tid] += fc;
        if (tid == 0) ctl[0] = T_GRAD;
        __syncthreads();
        DPROF(0);

        // ---- box-QP: warp 0 drives (element-wise work, triangular solves, decisions); the CTA factors and multiplies
        int it = 0, rounds = 0, status = QMPC_STATUS_MAXITER_;
        bool exact = false, refine = a.max_refine > 0, ipm_started = false;
        int rounds_left = 0, prev_changed = 1 << 30, round_no = 0, cpass = 0;
        real target = refine ? a.mu_switch : a.mu_tol, mu = 0, sigma = 0, so = 1, resfac = 1;
        const real inv2E = real(1) / real(2 * E);
        int trip = T_GRAD;
        if (warp == 0) {
            for (int e = lane; e < E; e += 32) c.usol[e] = real(0.5) * (lb + ub) - c.ubar[e];
        }
        __syncthreads();
        while (true) {
            trip = ctl[0];
            if (trip == T_DONE) break;
            DPROF(5);
            // (1) products with H
            if (trip == T_GRAD || trip == T_ADJ) { c.matvec(c.usol); __syncthreads(); }
            else if (trip == T_FIXED) {
                if (warp == 0) {
                    for (int e = lane; e < E; e += 32)
                        c.fv[e] = c.fx[e] == real(1) ? lb - c.ubar[e] : (c.fx[e] == real(2) ? ub - c.ubar[e] : real(0));
                }
                __syncthreads();
                c.matvec(c.fv);
                __syncthreads();
            }
            DPROF(1);
            // (2) right-hand side, factorisation
            if (trip == T_FIXED || trip == T_PRED) {
                if (warp == 0) {
                    if (trip == T_FIXED) {
                        for (int e = lane; e < E; e += 32) c.rt[e] = c.fx[e] != real(0) ? c.fv[e] : -c.f[e] - c.tv[e];
                    } else {
                        for (int e = lane; e < E; e += 32) {
                            const real d = c.ll[e] / c.tl[e] + c.lu[e] / c.tu[e];
                            c.dR[e] = d;
                            c.rt[e] = -c.f[e] + d * (c.ucur[e] - c.ubar[e]);
                        }
                    }
                }
                __syncthreads();
                c.factor(trip == T_FIXED);
            }
            DPROF(2);
            // (3) warp 0: solves, step logic, next trip
            if (warp == 0) {
                int next = T_DONE;
                if (trip == T_GRAD) {
                    real gs = 0;
                    for (int e = lane; e < E; e += 32) gs += fabs(c.tv[e] + c.f[e]);
                    gs = warp_sum(gs) / real(E);
                    const real lam0 = rfinite(gs) ? fmin(fmax(a.lam0_scale * gs, a.lam0_min), a.lam0_max) : a.lam0_min;
                    for (int e = lane; e < E; e += 32) {
                        const real u0 = real(0.5) * (lb + ub);
                        c.ucur[e] = u0; c.tl[e] = u0 - lb; c.tu[e] = ub - u0; c.ll[e] = lam0; c.lu[e] = lam0;
                    }
                    ipm_started = true;
                    next = T_PRED;
                } else if (trip == T_FIXED) {
                    c.solve(c.rt, c.usol);
                    next = T_ADJ;
                } else if (trip == T_ADJ) {
                    ++rounds;
                    int changed = 0;
                    for (int e = lane; e < E; e += 32) {
                        const real fxe = c.fx[e], un = c.ubar[e] + c.usol[e], gr = c.tv[e] + c.f[e];
                        if (fxe == real(1)) { if (gr < -a.refine_gtol) { c.fx[e] = 0; ++changed; } }
                        else if (fxe == real(2)) { if (gr > a.refine_gtol) { c.fx[e] = 0; ++changed; } }
                        else if (un < lb) { c.fx[e] = 1; ++changed; }
                        else if (un > ub) { c.fx[e] = 2; ++changed; }
                    }
                    changed = warp_sum(changed);
                    ++round_no;
                    if (!changed) { exact = true; status = QMPC_STATUS_OK_; next = T_DONE; }
                    else if (--rounds_left > 0 && !(round_no >= 3 && changed >= prev_changed)) { prev_changed = changed; next = T_FIXED; }
                    else { refine = false; target = a.mu_tol; next = T_PRED; }
                } else if (trip == T_PRED) {
                    c.solve(c.rt, c.usol);
                    real apm = 1, adm = 1;
                    for (int e = lane; e < E; e += 32) {
                        const real tl = c.tl[e], tu = c.tu[e];
                        const real du = c.ubar[e] + c.usol[e] - c.ucur[e];
                        const real dl = -c.ll[e] - c.ll[e] / tl * du;
                        const real dv = -c.lu[e] + c.lu[e] / tu * du;
                        c.ua[e] = c.usol[e];
                        c.cl[e] = du * dl; c.cu[e] = -du * dv;
                        c.rt[e] = dl; c.tv[e] = dv;
                        if (du < 0) apm = fmin(apm, -tl / du);
                        if (du > 0) apm = fmin(apm, tu / du);
                        if (dl < 0) adm = fmin(adm, -c.ll[e] / dl);
                        if (dv < 0) adm = fmin(adm, -c.lu[e] / dv);
                    }
                    const real apa = warp_min(apm), ada = warp_min(adm);
                    real s = 0;
                    for (int e = lane; e < E; e += 32) {
                        const real du = c.ubar[e] + c.ua[e] - c.ucur[e];
                        s += (c.ll[e] + ada * c.rt[e]) * (c.tl[e] + apa * du) + (c.lu[e] + ada * c.tv[e]) * (c.tu[e] - apa * du);
                    }
                    const real muaff = warp_sum(s) * inv2E;
                    sigma = muaff / mu; sigma = sigma * sigma * sigma;
                    so = 1; cpass = 0;
                    next = T_CORR;
                }
                if (trip == T_CORR || next == T_CORR) {
                    // corrector (increment on the predictor); blocked Mehrotra step -> once more as a centring step
                    while (true) {
                        const real smu = sigma * mu;
                        for (int e = lane; e < E; e += 32)
                            c.rt[e] = (smu - so * c.cl[e]) / c.tl[e] - (smu - so * c.cu[e]) / c.tu[e];
                        __syncwarp();
                        c.solve(c.rt, c.usol);
                        real apx = real(1e30), adx = real(1e30);
                        for (int e = lane; e < E; e += 32) {
                            const real tl = c.tl[e], tu = c.tu[e];
                            const real du = c.ubar[e] + c.ua[e] + c.usol[e] - c.ucur[e];
                            const real dl = (smu - so * c.cl[e]) / tl - c.ll[e] - c.ll[e] / tl * du;
                            const real dv = (smu - so * c.cu[e]) / tu - c.lu[e] + c.lu[e] / tu * du;
                            c.usol[e] = du; c.rt[e] = dl; c.tv[e] = dv;
                            if (du < 0) apx = fmin(apx, -tl / du);
                            if (du > 0) apx = fmin(apx, tu / du);
                            if (dl < 0) adx = fmin(adx, -c.ll[e] / dl);
                            if (dv < 0) adx = fmin(adx, -c.lu[e] / dv);
                        }
                        real ap = warp_min(apx), ad = warp_min(adx);
                        if (cpass == 0 && fmin(ap, ad) < real(0.5)) { so = 0; sigma = fmax(sigma, real(0.5)); cpass = 1; continue; }
                        ap = fmin(real(1), real(0.995) * ap); ad = fmin(real(1), real(0.995) * ad);
                        for (int e = lane; e < E; e += 32) {
                            const real du = ap * c.usol[e];
                            c.ucur[e] += du; c.tl[e] += du; c.tu[e] -= du;
                            c.ll[e] += ad * c.rt[e];
                            c.lu[e] += ad * c.tv[e];
                        }
                        resfac *= real(1) - fmin(ap, ad);
                        ++it;
                        break;
                    }
                    next = T_PRED;
                }
                if (next == T_PRED) {           // complementarity, convergence / hand-over test
                    __syncwarp();
                    real s = 0;
                    for (int e = lane; e < E; e += 32) s += c.ll[e] * c.tl[e] + c.lu[e] * c.tu[e];
                    mu = warp_sum(s) * inv2E;
                    if (!rfinite(mu)) { status = QMPC_STATUS_NAN_; next = T_DONE; }
                    else if (mu < target && resfac < real(1e-3)) {
                        if (refine) {
                            for (int e = lane; e < E; e += 32)
                                c.fx[e] = c.tl[e] < c.ll[e] ? real(1) : (c.tu[e] < c.lu[e] ? real(2) : real(0));
                            next = T_FIXED; rounds_left = a.max_refine; prev_changed = 1 << 30; round_no = 0;
                        } else { status = QMPC_STATUS_OK_; next = T_DONE; }
                    } else if (it >= a.max_iter) next = T_DONE;
                }
                __syncwarp();
                if (lane == 0) ctl[0] = next;
            }
            DPROF(3);
            __syncthreads();
        }
        // ---- result: new iterate, roll-out through the linearised dynamics (warp 0, lanes 0..15 carry the states)
        if (warp == 0) {
            real chk = 0;
            for (int e = lane; e < E; e += 32) { const real un = exact ? c.usol[e] : c.ucur[e]; chk += un - un; }
            chk = warp_sum(chk);
            if (!(chk == real(0))) status = QMPC_STATUS_NAN_;
            const bool good = status != QMPC_STATUS_NAN_;
            if (!good) {
                for (int e = lane; e < E; e += 32) actset[e] = 255;
                if (lane < 4) a.u0[(size_t)ocp * 4 + lane] = double(fmin(fmax(c.ubar[lane], lb), ub));
                if (lane == 0) { a.cost[ocp] = nan(""); a.status[ocp] = status; a.iters[ocp] = it; a.rounds[ocp] = rounds; }
            } else {
                for (int e = lane; e < E; e += 32) {
                    real un;
                    unsigned char fl;
                    if (exact) {
                        fl = c.fx[e] == real(1) ? 1 : (c.fx[e] == real(2) ? 2 : 0);
                        un = fl == 1 ? lb : (fl == 2 ? ub : c.ubar[e] + c.usol[e]);
                    } else {
                        un = fmin(fmax(c.ucur[e], lb), ub);
                        fl = status == QMPC_STATUS_OK_ ? (c.tl[e] < c.ll[e] ? 1 : (c.tu[e] < c.lu[e] ? 2 : 0)) : 255;
                    }
                    actset[e] = fl;
                    c.ucur[e] = un;
                }
                __syncwarp();
                for (int e = lane; e < E; e += 32) c.usol[e] = c.ucur[e] - c.ubar[e];
            }
            __syncwarp();
            HalfCtx<real> hc{a};
            hc.j = lane & 15; hc.N = N; hc.E = E; hc.hmask = 0xffffu << (lane & 16); hc.valid = lane < 16;
            hc.sidx = -1;
            hc.wv = small; hc.xp = small + 16; hc.usol = c.usol; hc.ubar = c.ubar;
            hc.Wv = Wv; hc.x0 = x0; hc.yref = yref; hc.yref_e = yref_e; hc.xit = xit; hc.uit = uit;
            real cost = hc.template forward<true>(0, false, good && lane < 16);
            cost = hc.hsum(cost);
            if (good) {
                for (int e = lane; e < E; e += 32) uit[e] = double(c.ucur[e]);
                if (lane < 4) a.u0[(size_t)ocp * 4 + lane] = double(c.ucur[lane]);
                if (lane == 0) { a.cost[ocp] = double(cost); a.status[ocp] = status; a.iters[ocp] = it; a.rounds[ocp] = rounds; }
            }
            if (a.timeline && lane == 0) a.timeline[2 * ocp + 1] = global_ns();
        }
        DPROF(4);
#ifdef QMPC_DENSE_PROF
        if (tid == 0 && item < 3)
            printf("dense ocp %d it %d rounds %d | cycles: condense %lld | matvec %lld (%d) | factor %lld (%d) | warp0 solve+logic %lld (%d) | rollout %lld | loop-top %lld | total %lld\n",
                   ocp, it, rounds, pf_acc[0], pf_acc[1], pf_n[1], pf_acc[2], pf_n[2], pf_acc[3], pf_n[3], pf_acc[4], pf_acc[5], clock64() - pf_t0);
#endif
        __syncthreads();
    }
}

}  // namespace qmpc
