// mpc_kernels_v2.cuh — K2, second mapping: TWO OCPs per warp (16 lanes each).
//
// Same algorithm as qmpc_ipm_kernel (mpc_kernels.cuh): warm-started primal-dual active-set rounds, cold Mehrotra IPM
// with Riccati factorisations, exact refinement, full step.  What changes is the mapping:
//   * a half-warp owns one OCP; lane j of the half is tile column j (0..3 B, 4..13 A', 14 b, 15 q) in the backward
//     sweeps and state row j in the forward sweep.  Each lane computes its whole column of y = P w and of M = W^T P W
//     (no half exchange), and the per-stage "glue" (4x4 Cholesky, triangular solves, masks) - which is executed
//     redundantly by every lane and dominates the instruction count - now serves two OCPs per warp instruction.
//   * per-OCP shared memory is cut to 8 KB (the IPM-only vectors and the state trajectory of the adjoint live in an
//     L2-resident global scratch), so 28 OCPs stay resident per SM with 14 warps.
//   * the two OCPs of a warp run the same trip sequence; trips of different kinds (active-set round vs IPM iteration)
//     share the factorisation and forward sweeps (per-half flags) and only the gradient sweeps are issued separately.
#pragma once
#include "mpc_kernels.cuh"

namespace qmpc {

template <typename real>
struct Ipm2Args {
    IpmArgs<real> b;       // everything of the one-OCP-per-warp kernel
    real* xtr;             // [B][(N+1)*13]  state trajectory of the last pinned forward sweep (adjoint input)
    real* ws;              // [B][5][4N]     ucur, ll, lu, tl, tu of the IPM
};

// per-OCP shared-memory carve-up (reals).  1000 reals = 8000 B = 62.5 x 128 B: the two OCPs of a warp sit 64 B apart
// in bank space, so their simultaneous 16-byte broadcasts never collide.
constexpr int S2_P = 0, S2_PV = 200, S2_WV = 216, S2_XP = 232, S2_HV = 248, S2_LS = 264, S2_CS = 328, S2_VEC = 360;
constexpr int S2_NVEC = 8;   // ubar, rdel, fx|cl, fv|cu, usol, grad|ua, dR, rt

template <typename real>
struct HalfCtx {
    const IpmArgs<real>& a;
    int j, sidx, N, E;
    unsigned hmask;
    bool valid;                      // this half owns a real OCP (stores allowed)
    real *P, *pv, *wv, *xp, *hv, *Ls, *cs;
    real *ubar, *rdel, *fx, *fv, *usol, *grad, *dR, *rt, *cl, *cu, *ua;
    real *ucur, *ll, *lu, *tl, *tu, *xtr;          // global scratch
    const real* Wv;
    real* facv;
    const double *x0, *yref, *yref_e;
    double *xit, *uit;

    __device__ __forceinline__ real hsum(real v) const { return half_sum(hmask, v); }
    __device__ __forceinline__ real hmin(real v) const
    {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) { const real w = __shfl_xor_sync(hmask, v, o); v = w < v ? w : v; }
        return v;
    }
    __device__ __forceinline__ int hsumi(int v) const { return half_sum(hmask, v); }

    // factorisation sweep.  fixed: active inputs of fx pinned at fv (gradient rdel, no barrier term); else IPM (dR, rt).
    // act: this half takes part (stores enabled).
    __device__ __forceinline__ void backward_full(const bool fixed, const bool act)
    {
        for (int idx = j; idx < 196; idx += 16) if (act) P[idx] = 0;
        __syncwarp();
        if (act && j < NX) {
            P[j * PS + j] = a.QNd[j];
            pv[j] = a.QNd[j] * real(xit[(size_t)N * NX + j] - yref_e[j]);
        }
        __syncwarp();
        for (int k = N - 1; k >= 0; --k) {
            const real* tile = Wv + (size_t)k * WT;
            real y[NX];
            real g;
            const real qj = sidx >= 0 ? __ldg(tile + sidx * 16 + 15) : real(0);
            {
                real w[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) w[i] = __ldg(tile + i * 16 + j);
                if (k > 0 && j < NX) prefetch_l1(Wv + (size_t)(k - 1) * WT + j * 16);
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const real* pr = P + i * PS;
                    real s0 = 0, s1 = 0;
#pragma unroll
                    for (int c = 0; c < 12; c += 2) {
                        real p0, p1;
                        ld2(pr + c, p0, p1);
                        s0 += p0 * w[c]; s1 += p1 * w[c + 1];
                    }
                    y[i] = s0 + s1 + pr[12] * w[12];
                }
                if (act && j == 14) {
#pragma unroll
                    for (int i = 0; i < NX; ++i) hv[i] = y[i] + pv[i];
                }
                __syncwarp();                      // hv visible; every read of the old P is done
                real g0 = 0, g1 = 0;
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    real h0, h1;
                    ld2(hv + c, h0, h1);
                    g0 += w[c] * h0; g1 += w[c + 1] * h1;
                }
                g = g0 + g1 + w[12] * hv[12];
            }
            g += (j < 4) ? (fixed ? rdel[k * 4 + j] : rt[k * 4 + j]) : qj;
            // rows 0..13 (tile columns) of this lane's column of M = W^T P W
            real m[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) m[i] = 0;
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                const real* tr = tile + r * 16;
#pragma unroll
                for (int c = 0; c < 14; c += 2) {
                    real t0, t1;
                    ldg2(tr + c, t0, t1);
                    m[c] += t0 * y[r]; m[c + 1] += t1 * y[r];
                }
            }
            const real dg = j < 4 ? a.Rd[j] + (fixed ? real(0) : dR[k * 4 + j]) : (j < 14 ? a.Qd[j - 1] : real(0));
            real keep[4] = {1, 1, 1, 1}, fva[4] = {0, 0, 0, 0};
            if (fixed) {
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) { keep[aa] = fx[k * 4 + aa] != real(0) ? real(0) : real(1); fva[aa] = fv[k * 4 + aa]; }
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) g = fma(m[aa], fva[aa], g);             // M[:,a] fv_a -> gradient
            }
            if (act && j < 4) {
                const real kme = sel4(keep, j);
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) cs[aa * 4 + j] = m[aa] * (keep[aa] * kme);
#pragma unroll
                for (int pi = 0; pi < 3; ++pi) cs[16 + pi * 4 + j] = y[pi];             // M[p_i][a] = (P w_a)[p_i]
                cs[28 + j] = g * kme;
                cs[j * 5] = (kme != real(0)) ? cs[j * 5] + dg : real(1);
            }
            __syncwarp();
            Chol4<real> L;
            real lg[4], lp[3][4], lj[4], lpme[4];
            real gpme = 0;
            {
                real Muu[16];
#pragma unroll
                for (int t = 0; t < 16; t += 2) ld2(cs + t, Muu[t], Muu[t + 1]);
                L.factor(Muu);
                real gu[4];
                ld2(cs + 28, gu[0], gu[1]); ld2(cs + 30, gu[2], gu[3]);
                L.fsolve(gu, lg);
#pragma unroll
                for (int pi = 0; pi < 3; ++pi) {
                    real mpu[4];
                    ld2(cs + 16 + pi * 4, mpu[0], mpu[1]); ld2(cs + 18 + pi * 4, mpu[2], mpu[3]);
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) mpu[aa] *= keep[aa];
                    L.fsolve(mpu, lp[pi]);
                }
                real mpu[4];
                const int pme = j < 3 ? j : 2;
                ld2(cs + 16 + pme * 4, mpu[0], mpu[1]); ld2(cs + 18 + pme * 4, mpu[2], mpu[3]);
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) { gpme = fma(mpu[aa], fva[aa], gpme); mpu[aa] *= keep[aa]; }
                L.fsolve(mpu, lpme);
            }
            {
                real mu4[4];
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) mu4[aa] = m[aa] * keep[aa];
                L.fsolve(mu4, lj);
            }
            if (act) {
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) Ls[j * 4 + aa] = lj[aa];
            }
            __syncwarp();
            if (act && k > 0) {
                if (j >= 4 && j < 14) {
                    const int sj = j - 1;
#pragma unroll
                    for (int i = 4; i < 14; ++i) {
                        real l0, l1, l2, l3;
                        ld2(Ls + i * 4, l0, l1); ld2(Ls + i * 4 + 2, l2, l3);
                        real v = fma(-l0, lj[0], m[i]);
                        v = fma(-l1, lj[1], v); v = fma(-l2, lj[2], v); v = fma(-l3, lj[3], v);
                        P[(i - 1) * PS + sj] = v;
                    }
                    P[sj * PS + sj] += dg;
#pragma unroll
                    for (int pi = 0; pi < 3; ++pi) {
                        const real v = y[pi] - dot4(lp[pi], lj);
                        P[pi * PS + sj] = v;
                        P[sj * PS + pi] = v;
                    }
                    pv[sj] = g - dot4(lj, lg);
                }
                if (j < 3) {
#pragma unroll
                    for (int pi = 0; pi < 3; ++pi) P[pi * PS + j] -= dot4(lp[pi], lpme);
                    P[j * PS + j] += a.Qd[j];
                    pv[j] = hv[j] + qj + gpme - dot4(lpme, lg);
                }
            }
            if (act) {
                real* f = facv + (size_t)k * FAC;
                if (j >= 4 && j < 14) {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) f[(j - 1) * 4 + aa] = lj[aa];
                } else if (j < 3) {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) f[j * 4 + aa] = lpme[aa];
                } else if (j == 14) {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) f[52 + aa] = lg[aa];
                } else if (j == 15) {
                    L.store(f + 56);
                }
            }
            __syncwarp();
        }
    }

    // gradient-only backward sweep.  adjoint = false: IPM corrector (input gradient rt, homogeneous), writes lgc.
    //                                adjoint = true : multipliers at (xtr, usol) -> grad.
    __device__ __forceinline__ void backward_grad(const bool adjoint, const bool act)
    {
        if (act && j < 16) {
            real v = 0;
            if (adjoint && j < NX) v = a.QNd[j] * (xtr[(size_t)N * NX + j] + real(xit[(size_t)N * NX + j] - yref_e[j]));
            pv[j] = v;
        }
        __syncwarp();
        for (int k = N - 1; k >= 0; --k) {
            const real* tile = Wv + (size_t)k * WT;
            const real* f = facv + (size_t)k * FAC;
            real g0 = 0, g1 = 0;
#pragma unroll
            for (int c = 0; c < 12; c += 2) {
                real h0, h1;
                ld2(pv + c, h0, h1);
                g0 += __ldg(tile + c * 16 + j) * h0; g1 += __ldg(tile + (c + 1) * 16 + j) * h1;
            }
            real g = g0 + g1 + __ldg(tile + 12 * 16 + j) * pv[12];
            if (k > 0 && j < NX) prefetch_l1(Wv + (size_t)(k - 1) * WT + j * 16);
            const real pold = j < 3 ? pv[j] : real(0);
            real corr = 0;          // what is subtracted from / added to the propagated costate of "my" state
            if (!adjoint) {
                Chol4<real> L;
                L.load(f + 56);
                if (j < 4) g += rt[k * 4 + j];
                real gu[4], lgc[4], lx[4] = {0, 0, 0, 0};
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) gu[aa] = __shfl_sync(hmask, g, (threadIdx.x & 16) + aa);
                L.fsolve(gu, lgc);
                if (sidx >= 0) { ld2(f + sidx * 4, lx[0], lx[1]); ld2(f + sidx * 4 + 2, lx[2], lx[3]); }
                corr = -dot4(lx, lgc);
                if (act && j == 14) {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) facv[(size_t)k * FAC + 66 + aa] = lgc[aa];
                }
            } else {
                const real qj = sidx >= 0 ? __ldg(tile + sidx * 16 + 15) : real(0);
                const real xk = sidx >= 0 ? xtr[k * NX + sidx] : real(0);
                corr = (sidx >= 0 ? a.Qd[sidx] : real(0)) * xk + qj;
                if (act && j < 4) grad[k * 4 + j] = g + a.Rd[j] * usol[k * 4 + j] + rdel[k * 4 + j];
            }
            __syncwarp();
            if (act) {
                if (j >= 4 && j < 14) pv[j - 1] = g + corr;
                else if (j < 3) pv[j] = pold + corr;
            }
            __syncwarp();
        }
    }

    // forward sweep (row j of the tile per lane).  mode 0: feedback with lg and offset b; mode 1: feedback with lgc,
    // homogeneous; fixed: pinned inputs take fv and the trajectory is stored in xtr (adjoint input).
    // ROLLOUT: open loop with usol, writes the new iterate, returns this lane's share of the objective.
    template <bool ROLLOUT>
    __device__ __forceinline__ real forward(const int mode, const bool fixed, const bool act)
    {
        const bool hom = !ROLLOUT && mode == 1;
        if (act) {
            if (j < NX) {
                const real v = hom ? real(0) : real(x0[j] - xit[j]);
                if (j < 3) xp[j] = v; else wv[j + 1] = v;
                if (fixed) xtr[j] = v;
            } else if (j == 14) wv[14] = hom ? real(0) : real(1);
            else if (j == 15) wv[15] = 0;
        }
        real cost = 0;
        if (ROLLOUT && act && j < NX) {
            const real e0 = real(x0[j] - yref[j]);
            cost = real(0.5) * a.Qd[j] * e0 * e0;
            xit[j] = x0[j];
        }
        __syncwarp();
        const int irow = j < NX ? j : NX - 1;
        for (int k = 0; k < N; ++k) {
            const real* tr = Wv + (size_t)k * WT + irow * 16;
            real wr[16];
#pragma unroll
            for (int c = 0; c < 16; c += 2) ldg2(tr + c, wr[c], wr[c + 1]);
            if (k + 1 < N && j < NX) prefetch_l1(Wv + (size_t)(k + 1) * WT + j * 16);
            real u[4];
            if (!ROLLOUT) {
                const real* f = facv + (size_t)k * FAC;
                Chol4<real> L;
                L.load(f + 56);
                const real xo = j < 3 ? xp[j] : (j < NX ? wv[j + 1] : real(0));
                real v[4] = {0, 0, 0, 0};
                const int src = j < NX ? j * 4 : (mode == 0 ? 52 : 66);
                if (j <= NX) { ld2(f + src, v[0], v[1]); ld2(f + src + 2, v[2], v[3]); }
                const real sc = j < NX ? xo : real(1);
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) v[aa] = hsum(v[aa] * sc);
                L.bsolve_neg(v, u);
                if (fixed) {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) if (fx[k * 4 + aa] != real(0)) u[aa] = fv[k * 4 + aa];
                }
            } else {
                ld2(usol + k * 4, u[0], u[1]); ld2(usol + k * 4 + 2, u[2], u[3]);
            }
            __syncwarp();                       // everyone has read the old state
            if (act && j < 4) {
                const real ul = sel4(u, j);
                wv[j] = ul;
                if (!ROLLOUT) usol[k * 4 + j] = ul;
            }
            __syncwarp();
            real acc0 = 0, acc1 = 0;
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                real v0, v1;
                ld2(wv + c, v0, v1);
                acc0 += wr[c] * v0; acc1 += wr[c + 1] * v1;
            }
            real acc = acc0 + acc1;
            if (j < 3) acc += xp[j];
            __syncwarp();
            if (act && j < NX) {
                if (j < 3) xp[j] = acc; else wv[j + 1] = acc;
                if (fixed) xtr[(k + 1) * NX + j] = acc;
                if (ROLLOUT) {
                    double* xo = xit + (size_t)(k + 1) * NX + j;
                    const double xnew = *xo + double(acc);
                    *xo = xnew;
                    const real wgt = (k + 1 < N) ? a.Qd[j] : a.QNd[j];
                    const double ref = (k + 1 < N) ? yref[(size_t)(k + 1) * NY + j] : yref_e[j];
                    const real e = real(xnew - ref);
                    cost += real(0.5) * wgt * e * e;
                }
                if (ROLLOUT && j < 4) {
                    const real e = real(double(ubar[k * 4 + j]) + double(sel4(u, j)) - yref[(size_t)k * NY + NX + j]);
                    cost += real(0.5) * a.Rd[j] * e * e;
                }
            }
            __syncwarp();
        }
        return cost;
    }
};

__device__ __forceinline__ bool warp_any(bool p) { return warp_max(int(p)) != 0; }

#ifndef QMPC_IPM2_MIN_WARPS
#define QMPC_IPM2_MIN_WARPS 14
#endif

template <typename real, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, QMPC_IPM2_MIN_WARPS / WARPS) qmpc_ipm2_kernel(Ipm2Args<real> aa)
{
    QMPC_DYN_SMEM(smem_raw);
    const IpmArgs<real>& a = aa.b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4;
    const int ocp_raw = (blockIdx.x * WARPS + warp) * 2 + h;
    if ((blockIdx.x * WARPS + warp) * 2 >= a.B) return;      // whole warp out of range
    const int N = a.N, E = 4 * N;
    HalfCtx<real> c{a};
    c.valid = ocp_raw < a.B;
    const int ocp = c.valid ? ocp_raw : a.B - 1;               // an odd tail half shadows the last OCP without storing
    real* sm = reinterpret_cast<real*>(smem_raw) + (size_t)(warp * 2 + h) * a.smem_per_warp;
    c.j = lane & 15; c.N = N; c.E = E;
    c.hmask = 0xffffu << (lane & 16);
    c.sidx = (c.j < 3) ? c.j : ((c.j >= 4 && c.j < 14) ? c.j - 1 : -1);
    c.P = sm + S2_P; c.pv = sm + S2_PV; c.wv = sm + S2_WV; c.xp = sm + S2_XP; c.hv = sm + S2_HV;
    c.Ls = sm + S2_LS; c.cs = sm + S2_CS;
    real* v = sm + S2_VEC;
    c.ubar = v; c.rdel = v + E; c.fx = v + 2 * E; c.fv = v + 3 * E; c.usol = v + 4 * E; c.grad = v + 5 * E;
    c.dR = v + 6 * E; c.rt = v + 7 * E;
    c.cl = c.fx; c.cu = c.fv; c.ua = c.grad;
    real* ws = aa.ws + (size_t)ocp * 5 * E;
    c.ucur = ws; c.ll = ws + E; c.lu = ws + 2 * E; c.tl = ws + 3 * E; c.tu = ws + 4 * E;
    c.xtr = aa.xtr + (size_t)ocp * (N + 1) * NX;
    c.Wv = a.W + (size_t)ocp * N * WT;
    c.facv = a.fac + (size_t)ocp * N * FAC;
    c.x0 = a.x0 + (size_t)ocp * NX;
    c.yref = a.yref + (size_t)ocp * N * NY;
    c.yref_e = a.yref_e + (size_t)ocp * NX;
    c.xit = a.xit + (size_t)ocp * (N + 1) * NX;
    c.uit = a.uit + (size_t)ocp * N * NU;
    unsigned char* actset = a.act + (size_t)ocp * E;
    const int j = c.j;
    const bool valid = c.valid;
    const real lb = a.lb, ub = a.ub;
    if (a.timeline && valid && j == 0) a.timeline[2 * ocp] = global_ns();

    for (int e = j; e < E; e += 16) {
        const real ub_ = real(c.uit[e]);
        c.ubar[e] = ub_;
        c.rdel[e] = a.Rd[e & 3] * (ub_ - real(c.yref[(size_t)(e >> 2) * NY + NX + (e & 3)]));
    }
    __syncwarp();

    // trips: see qmpc_ipm_kernel; G_FWD/G_ADJ = roll-out + adjoint at the box centre that scales the IPM's start
    enum { T_FIXED, T_ADJ, T_PRED, T_CORR, T_GFWD, T_GADJ, T_DONE };
    int it = 0, rounds = 0, status = QMPC_STATUS_MAXITER_;
    bool exact = false, refine = a.max_refine > 0, ipm_started = false, handed = false;
    int rounds_left = 0, prev_changed = 1 << 30, round_no = 0, trip = T_GFWD, cpass = 0, attempts = 0;
    real target = refine ? a.mu_switch : a.mu_tol, mu = 0, sigma = 0, so = 1, resfac = 1;
    const real inv2E = real(1) / real(2 * E);
    if (a.warm_rounds > 0) {
        int known = 1;
        for (int e = j; e < E; e += 16) { const unsigned char f = actset[e]; if (f > 2) known = 0; c.fx[e] = real(f <= 2 ? f : 0); }
        known = c.hmin(real(known)) > real(0.5) ? 1 : 0;
        if (known) { trip = T_FIXED; rounds_left = a.warm_rounds; }
    }
    if (a.hard_count && trip == T_GFWD) { handed = true; trip = T_DONE; }
    if (!valid) { trip = T_DONE; handed = false; }

    while (warp_any(trip != T_DONE)) {
        // ---- prepare this half's trip
        if (trip == T_GFWD) {
            for (int e = j; e < E; e += 16) { c.fx[e] = 1; c.fv[e] = real(0.5) * (lb + ub) - c.ubar[e]; }
        } else if (trip == T_PRED) {
            if (!ipm_started) {
                real gs = 0;
                for (int e = j; e < E; e += 16) gs += fabs(c.grad[e]);
                gs = c.hsum(gs) / real(E);
                const real lam0 = rfinite(gs) ? fmin(fmax(a.lam0_scale * gs, a.lam0_min), a.lam0_max) : a.lam0_min;
                for (int e = j; e < E; e += 16) {
                    const real u0 = real(0.5) * (lb + ub);
                    c.ucur[e] = u0; c.tl[e] = u0 - lb; c.tu[e] = ub - u0; c.ll[e] = lam0; c.lu[e] = lam0;
                }
                ipm_started = true;
            }
            real s = 0;
            for (int e = j; e < E; e += 16) s += c.ll[e] * c.tl[e] + c.lu[e] * c.tu[e];
            mu = c.hsum(s) * inv2E;
            if (!rfinite(mu)) { status = QMPC_STATUS_NAN_; trip = T_DONE; }
            else if (mu < target && resfac < (refine ? real(1e-3) : a.resfac_final)) {
                if (refine) {
                    for (int e = j; e < E; e += 16)
                        c.fx[e] = c.tl[e] < c.ll[e] ? real(1) : (c.tu[e] < c.lu[e] ? real(2) : real(0));
                    trip = T_FIXED; rounds_left = a.max_refine; prev_changed = 1 << 30; round_no = 0;
                } else { status = QMPC_STATUS_OK_; trip = T_DONE; }
            } else if (it >= ((a.fail_streak && a.fail_streak[ocp] >= 2) ? a.max_iter_failed : a.max_iter)) trip = T_DONE;
            if (trip == T_PRED) {
                for (int e = j; e < E; e += 16) {
                    const real d = c.ll[e] / c.tl[e] + c.lu[e] / c.tu[e];
                    c.dR[e] = d;
                    c.rt[e] = c.rdel[e] - d * (c.ucur[e] - c.ubar[e]);
                }
            }
        }
        if (trip == T_FIXED) {
            for (int e = j; e < E; e += 16) c.fv[e] = c.fx[e] == real(1) ? lb - c.ubar[e] : (c.fx[e] == real(2) ? ub - c.ubar[e] : real(0));
        }
        __syncwarp();
        // ---- sweeps (issued once per warp, per-half flags)
        const bool kF = trip == T_FIXED, kP = trip == T_PRED, kC = trip == T_CORR, kA = trip == T_ADJ, kGF = trip == T_GFWD, kGA = trip == T_GADJ;
        if (warp_any(kF || kP)) c.backward_full(kF, kF || kP);
        if (warp_any(kC)) c.backward_grad(false, kC);
        if (warp_any(kA || kGA)) c.backward_grad(true, kA || kGA);
        if (warp_any(kF || kP || kC || kGF)) c.template forward<false>(kC ? 1 : 0, kF || kGF, kF || kP || kC || kGF);
        __syncwarp();
        // ---- what the trip was for
        if (kGF) trip = T_GADJ;
        else if (kGA) trip = T_PRED;
        else if (kF) trip = T_ADJ;
        else if (kA) {
            ++rounds;
            int changed = 0;
            for (int e = j; e < E; e += 16) {
                const real f = c.fx[e], un = c.ubar[e] + c.usol[e], gr = c.grad[e];
                if (f == real(1)) { if (gr < -a.refine_gtol) { c.fx[e] = 0; ++changed; } }
                else if (f == real(2)) { if (gr > a.refine_gtol) { c.fx[e] = 0; ++changed; } }
                else if (un < lb) { c.fx[e] = 1; ++changed; }
                else if (un > ub) { c.fx[e] = 2; ++changed; }
            }
            changed = c.hsumi(changed);
            ++round_no;
            if (!changed) { exact = true; status = QMPC_STATUS_OK_; trip = T_DONE; }
            else if (--rounds_left > 0 && !((!ipm_started || a.post_bail) && round_no >= 3 && changed >= prev_changed)) { prev_changed = changed; trip = T_FIXED; }
            else {                                   // not settling: (re)enter the IPM
                if (ipm_started) { if (++attempts < 2) target *= real(1e-2); else { refine = false; target = a.mu_tol; } trip = T_PRED; }
                else if (a.hard_count) { handed = true; trip = T_DONE; }     // screening mode: the dense kernel takes it
                else trip = T_GFWD;
            }
        } else if (kP) {
            real apm = 1, adm = 1;
            for (int e = j; e < E; e += 16) {
                const real tl = c.tl[e], tu = c.tu[e];
                const real du = c.ubar[e] + c.usol[e] - c.ucur[e];
                const real dl = -c.ll[e] - c.ll[e] / tl * du;
                const real dv = -c.lu[e] + c.lu[e] / tu * du;
                c.ua[e] = c.usol[e];
                c.cl[e] = du * dl; c.cu[e] = -du * dv;
                c.rt[e] = dl; c.dR[e] = dv;
                if (du < 0) apm = fmin(apm, -tl / du);
                if (du > 0) apm = fmin(apm, tu / du);
                if (dl < 0) adm = fmin(adm, -c.ll[e] / dl);
                if (dv < 0) adm = fmin(adm, -c.lu[e] / dv);
            }
            const real apa = c.hmin(apm), ada = c.hmin(adm);
            real s = 0;
            for (int e = j; e < E; e += 16) {
                const real du = c.ubar[e] + c.ua[e] - c.ucur[e];
                s += (c.ll[e] + ada * c.rt[e]) * (c.tl[e] + apa * du) + (c.lu[e] + ada * c.dR[e]) * (c.tu[e] - apa * du);
            }
            const real muaff = c.hsum(s) * inv2E;
            sigma = muaff / mu; sigma = sigma * sigma * sigma;
            so = 1; cpass = 0;
            const real smu = sigma * mu;
            for (int e = j; e < E; e += 16)
                c.rt[e] = -(smu - so * c.cl[e]) / c.tl[e] + (smu - so * c.cu[e]) / c.tu[e];
            trip = T_CORR;
        } else if (kC) {
            const real smu = sigma * mu;
            real apx = real(1e30), adx = real(1e30);
            for (int e = j; e < E; e += 16) {
                const real tl = c.tl[e], tu = c.tu[e];
                const real du = c.ubar[e] + c.ua[e] + c.usol[e] - c.ucur[e];
                const real dl = (smu - so * c.cl[e]) / tl - c.ll[e] - c.ll[e] / tl * du;
                const real dv = (smu - so * c.cu[e]) / tu - c.lu[e] + c.lu[e] / tu * du;
                c.usol[e] = du; c.rt[e] = dl; c.dR[e] = dv;
                if (du < 0) apx = fmin(apx, -tl / du);
                if (du > 0) apx = fmin(apx, tu / du);
                if (dl < 0) adx = fmin(adx, -c.ll[e] / dl);
                if (dv < 0) adx = fmin(adx, -c.lu[e] / dv);
            }
            real ap = c.hmin(apx), ad = c.hmin(adx);
            if (cpass == 0 && fmin(ap, ad) < real(0.5)) {      // blocked Mehrotra step: redo as a centring step
                so = 0; sigma = fmax(sigma, real(0.5)); cpass = 1;
                const real smu2 = sigma * mu;
                for (int e = j; e < E; e += 16) c.rt[e] = -smu2 / c.tl[e] + smu2 / c.tu[e];
            } else {
                ap = fmin(real(1), real(0.995) * ap); ad = fmin(real(1), real(0.995) * ad);
                for (int e = j; e < E; e += 16) {
                    const real du = ap * c.usol[e];
                    c.ucur[e] += du; c.tl[e] += du; c.tu[e] -= du;
                    c.ll[e] += ad * c.rt[e];
                    c.lu[e] += ad * c.dR[e];
                }
                resfac *= real(1) - fmin(ap, ad);
                ++it;
                trip = T_PRED;
            }
        }
        __syncwarp();
    }
    // ---- result
    if (handed) {
        if (j == 0) a.hard_list[atomicAdd(a.hard_count, 1)] = ocp;
    }
    if (valid && !handed) {
        real chk = 0;
        for (int e = j; e < E; e += 16) { const real un = exact ? c.usol[e] : (ipm_started ? c.ucur[e] : real(0)); chk += un - un; }
        chk = c.hsum(chk);
        if (!(chk == real(0)) || (!exact && !ipm_started)) status = QMPC_STATUS_NAN_;
    }
    const bool good = valid && !handed && status != QMPC_STATUS_NAN_;
    if (valid && !handed && !good) {
        for (int e = j; e < E; e += 16) actset[e] = 255;
        if (j < 4) a.u0[(size_t)ocp * 4 + j] = double(fmin(fmax(c.ubar[j], lb), ub));
        if (j == 0) { a.cost[ocp] = nan(""); a.status[ocp] = status; a.iters[ocp] = it; a.rounds[ocp] = rounds; }
    }
    if (good) {
        for (int e = j; e < E; e += 16) {
            real un;
            unsigned char f;
            if (exact) {
                f = c.fx[e] == real(1) ? 1 : (c.fx[e] == real(2) ? 2 : 0);
                un = f == 1 ? lb : (f == 2 ? ub : c.ubar[e] + c.usol[e]);
            } else {
                un = fmin(fmax(c.ucur[e], lb), ub);
                f = status == QMPC_STATUS_OK_ ? (c.tl[e] < c.ll[e] ? 1 : (c.tu[e] < c.lu[e] ? 2 : 0)) : 255;
            }
            actset[e] = f;
            c.ucur[e] = un;
        }
        for (int e = j; e < E; e += 16) c.usol[e] = c.ucur[e] - c.ubar[e];
    }
    __syncwarp();
    real cost = c.template forward<true>(0, false, good);
    cost = c.hsum(cost);
    if (good) {
        for (int e = j; e < E; e += 16) c.uit[e] = double(c.ucur[e]);
        if (j < 4) a.u0[(size_t)ocp * 4 + j] = double(c.ucur[j]);
        if (j == 0) { a.cost[ocp] = double(cost); a.status[ocp] = status; a.iters[ocp] = it; a.rounds[ocp] = rounds; }
    }
    if (a.timeline && valid && !handed && j == 0) a.timeline[2 * ocp + 1] = global_ns();
}

}  // namespace qmpc
