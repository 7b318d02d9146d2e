// rgp_kernels.cuh — batched recursive-GP kernels (always fp64).
//
// One RGP per (vehicle, body axis): fixed basis points X[M], mean mu[M], covariance C[M][M]
// (reference src/gp/RGP.py:106-157).  One warp owns one (vehicle, axis) model; lanes stride the basis index so
// that rows of C and K_x^-1 are read coalesced.  C is streamed from HBM twice per update (second pass hits L2/L1):
// the update is HBM-bound, 16*M^2 algorithmic bytes per (vehicle, axis).
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace qmpc {

struct RgpArgs {
    int B, M;
    const double* X;       // [3][M]
    const double* theta;   // [3][3] (L, sigma_f, sigma_n)
    const double* Kx_inv;  // [3][M][M]
    double* mu;            // [B][3][M]
    double* C;             // [B][3][M][M]
    double* alpha;         // [B][3][M]  = Kx_inv mu   (may be null)
    const double* xt;      // [B][3]
    const double* yt;      // [B][3]
};

__device__ __forceinline__ double rbf_k(double a, double b, double iL2, double sf2)
{
    const double e = a - b;
    return sf2 * exp(-0.5 * e * iL2 * e);
}

constexpr int RGP_MAXT = 4;   // lanes stride M; supports M <= 128

// RGP.regress with one sample (RGP.py:303-330 via predict :199-208), operand order of the reference:
//   Jt = k(x,X) Kx^-1 ; m = Jt mu ; b = k(x,x) - Jt k(X,x) ; w = Jt C ; cj = C Jt^T ; s = b + w Jt^T + sn^2
//   mu += cj (1/s) (y - m) ; C -= (cj/s) w        (the row w = Jt C is used, C is never re-symmetrised)
// dynamic smem per warp: 3*M doubles (kv, Jt, cj)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) qrgp_regress_kernel(RgpArgs a)
{
    QMPC_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int model = blockIdx.x * WARPS + warp;           // (vehicle, axis)
    if (model >= a.B * 3) return;
    const int M = a.M, d = model % 3;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 3 * M;
    double *kv = sm, *Jt = sm + M, *cj = sm + 2 * M;
    const double L = a.theta[3 * d], sf = a.theta[3 * d + 1], sn = a.theta[3 * d + 2];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    const double* X = a.X + d * M;
    const double* Ki = a.Kx_inv + (size_t)d * M * M;
    double* mu = a.mu + (size_t)model * M;
    double* Cm = a.C + (size_t)model * M * M;
    const double xt = a.xt[model], yt = a.yt[model];
    if (!(xt - xt == 0.0) || !(yt - yt == 0.0)) return;    // non-finite sample = "no update for this axis" (single-axis RGP.regress; a crashed vehicle)

    for (int i = lane; i < M; i += 32) kv[i] = rbf_k(xt, X[i], iL2, sf2);
    __syncwarp();
    double mp = 0, jk = 0;
    for (int j = lane; j < M; j += 32) {
        double s = 0;
        for (int i = 0; i < M; ++i) s += kv[i] * __ldg(Ki + (size_t)i * M + j);
        Jt[j] = s;
        mp += s * mu[j];
        jk += s * kv[j];
    }
    mp = warp_sum(mp); jk = warp_sum(jk);
    __syncwarp();
    double w[RGP_MAXT], jl[RGP_MAXT];
#pragma unroll
    for (int t = 0; t < RGP_MAXT; ++t) { w[t] = 0; const int j = lane + 32 * t; jl[t] = j < M ? Jt[j] : 0.0; }
    for (int i = 0; i < M; ++i) {
        const double ji = Jt[i];
        double rs = 0;
#pragma unroll
        for (int t = 0; t < RGP_MAXT; ++t) {
            const int j = lane + 32 * t;
            if (j < M) { const double c = Cm[(size_t)i * M + j]; w[t] += ji * c; rs += c * jl[t]; }
        }
        rs = warp_sum(rs);
        if (lane == 0) cj[i] = rs;
    }
    double jcj = 0;
#pragma unroll
    for (int t = 0; t < RGP_MAXT; ++t) jcj += w[t] * jl[t];
    jcj = warp_sum(jcj);
    __syncwarp();
    const double b = rbf_k(xt, xt, iL2, sf2) - jk;
    const double sinv = 1.0 / (b + jcj + sn * sn);
    const double innov = yt - mp;
    for (int i = 0; i < M; ++i) {
        const double g = cj[i] * sinv;
#pragma unroll
        for (int t = 0; t < RGP_MAXT; ++t) {
            const int j = lane + 32 * t;
            if (j < M) Cm[(size_t)i * M + j] -= g * w[t];
        }
    }
    for (int i = lane; i < M; i += 32) { const double v = mu[i] + cj[i] * sinv * innov; mu[i] = v; kv[i] = v; }
    __syncwarp();
    if (a.alpha) {
        double* al = a.alpha + (size_t)model * M;
        for (int i = lane; i < M; i += 32) {
            double s = 0;
            for (int j = 0; j < M; ++j) s += __ldg(Ki + (size_t)i * M + j) * kv[j];
            al[i] = s;
        }
    }
}

// alpha = Kx_inv y  for y [B][3][M]   (constant part of RGP.predict_using_y, RGP.py:252-254)
__global__ void qrgp_alpha_kernel(int B, int M, const double* Kx_inv, const double* y, double* alpha)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3 * M) return;
    const int i = t % M, model = t / M, d = model % 3;
    const double* Ki = Kx_inv + (size_t)d * M * M + (size_t)i * M;
    const double* yy = y + (size_t)model * M;
    double s = 0;
    for (int j = 0; j < M; ++j) s += __ldg(Ki + j) * yy[j];
    alpha[t] = s;
}

struct RgpPredArgs {
    int B, M, m;
    const double* X; const double* theta; const double* Kx_inv;
    const double* mu;      // [B][3][M]  (or y for predict_using_y)
    const double* C;       // [B][3][M][M] (null: mean only)
    const double* xs;      // [B][3][m]
    double* mean;          // [B][3][m]
    double* var;           // [B][3][m] or null
};

// RGP.predict / predict_using_y numpy branches (RGP.py:195-210, 264-283): one warp per (vehicle, axis),
// loop over query points.  dynamic smem per warp: 2*M doubles
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) qrgp_predict_kernel(RgpPredArgs a)
{
    QMPC_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int model = blockIdx.x * WARPS + warp;
    if (model >= a.B * 3) return;
    const int M = a.M, d = model % 3;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * M;
    double *kv = sm, *Jt = sm + M;
    const double L = a.theta[3 * d], sf = a.theta[3 * d + 1];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    const double* X = a.X + d * M;
    const double* Ki = a.Kx_inv + (size_t)d * M * M;
    const double* mu = a.mu + (size_t)model * M;
    for (int q = 0; q < a.m; ++q) {
        const double xt = a.xs[(size_t)model * a.m + q];
        __syncwarp();
        for (int i = lane; i < M; i += 32) kv[i] = rbf_k(xt, X[i], iL2, sf2);
        __syncwarp();
        double mp = 0, jk = 0;
        for (int j = lane; j < M; j += 32) {
            double s = 0;
            for (int i = 0; i < M; ++i) s += kv[i] * __ldg(Ki + (size_t)i * M + j);
            Jt[j] = s;
            mp += s * mu[j];
            jk += s * kv[j];
        }
        mp = warp_sum(mp); jk = warp_sum(jk);
        __syncwarp();
        if (lane == 0) a.mean[(size_t)model * a.m + q] = mp;
        if (a.var && a.C) {
            const double* Cm = a.C + (size_t)model * M * M;
            double jcj = 0;
            for (int j = lane; j < M; j += 32) {
                double s = 0;
                for (int i = 0; i < M; ++i) s += Jt[i] * Cm[(size_t)i * M + j];
                jcj += s * Jt[j];
            }
            jcj = warp_sum(jcj);
            if (lane == 0) a.var[(size_t)model * a.m + q] = rbf_k(xt, xt, iL2, sf2) - jk + jcj;
        }
    }
}

// ------------------------------------------------------------------ shared-swarm (information form)

struct RgpSharedArgs {
    int B, M;
    const double* X; const double* theta; const double* Kx_inv;
    const double* xt;   // [B][3]
    const double* yt;   // [B][3]
    double* info;       // [3][M*M + M]  (Lambda row-major, then eta), accumulated with atomics
};

// Each warp walks a strided subset of vehicles for one axis and accumulates j^T j / r and j^T y / r into a
// warp-private shared-memory tile; one atomicAdd per entry per warp at the end.
// dynamic smem per warp: (M*M + 3*M) doubles
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) qrgp_shared_accumulate_kernel(RgpSharedArgs a)
{
    QMPC_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.y, M = a.M;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (M * M + 3 * M);
    double *acc = sm, *eta = sm + M * M, *kv = eta + M, *Jt = kv + M;
    const double L = a.theta[3 * d], sf = a.theta[3 * d + 1], sn = a.theta[3 * d + 2];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    const double* X = a.X + d * M;
    const double* Ki = a.Kx_inv + (size_t)d * M * M;
    for (int t = lane; t < M * M + M; t += 32) acc[t] = 0;
    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    for (int v = gw; v < a.B; v += nw) {
        const double xt = a.xt[(size_t)v * 3 + d], yt = a.yt[(size_t)v * 3 + d];
        if (!(xt - xt == 0.0) || !(yt - yt == 0.0)) continue;   // a vehicle with a non-finite residual must not poison the shared model
        __syncwarp();
        for (int i = lane; i < M; i += 32) kv[i] = rbf_k(xt, X[i], iL2, sf2);
        __syncwarp();
        double jk = 0;
        for (int j = lane; j < M; j += 32) {
            double s = 0;
            for (int i = 0; i < M; ++i) s += kv[i] * __ldg(Ki + (size_t)i * M + j);
            Jt[j] = s;
            jk += s * kv[j];
        }
        jk = warp_sum(jk);
        __syncwarp();
        const double rinv = 1.0 / (rbf_k(xt, xt, iL2, sf2) - jk + sn * sn);
        for (int t = lane; t < M * M; t += 32) { const int i = t / M, j = t - i * M; acc[t] += Jt[i] * Jt[j] * rinv; }
        for (int i = lane; i < M; i += 32) eta[i] += Jt[i] * yt * rinv;
    }
    __syncwarp();
    double* out = a.info + (size_t)d * (M * M + M);
    for (int t = lane; t < M * M + M; t += 32) atomicAdd(out + t, acc[t]);
}

// Posterior of the shared model after the all-reduce (one CTA per axis, Gauss-Jordan in shared memory):
//   (I + C Lambda) [C_new | mu_new] = [C | mu + C eta]
// dynamic smem: M*(2M+1) doubles
__global__ void __launch_bounds__(256) qrgp_shared_apply_kernel(int M, const double* info, double* mu, double* C,
                                                                 const double* Kx_inv, double* alpha)
{
    QMPC_DYN_SMEM(smem_raw);
    const int d = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int Wd = 2 * M + 1;
    double* A = reinterpret_cast<double*>(smem_raw);     // [M][2M+1] : (I + C Lambda | C | mu + C eta)
    const double* Lam = info + (size_t)d * (M * M + M);
    const double* eta = Lam + M * M;
    double* Cm = C + (size_t)d * M * M;
    double* mud = mu + (size_t)d * M;
    for (int t = tid; t < M * M; t += nt) {
        const int i = t / M, j = t - i * M;
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = 0; k < M; ++k) s += Cm[i * M + k] * Lam[k * M + j];
        A[i * Wd + j] = s;
        A[i * Wd + M + j] = Cm[t];
    }
    for (int i = tid; i < M; i += nt) {
        double s = mud[i];
        for (int k = 0; k < M; ++k) s += Cm[i * M + k] * eta[k];
        A[i * Wd + 2 * M] = s;
    }
    __syncthreads();
    __shared__ int piv_s;
    for (int c = 0; c < M; ++c) {
        if (tid == 0) {
            int p = c; double best = fabs(A[c * Wd + c]);
            for (int i = c + 1; i < M; ++i) { const double v = fabs(A[i * Wd + c]); if (v > best) { best = v; p = i; } }
            piv_s = p;
        }
        __syncthreads();
        const int p = piv_s;
        if (p != c) for (int j = tid; j < Wd; j += nt) { const double t = A[c * Wd + j]; A[c * Wd + j] = A[p * Wd + j]; A[p * Wd + j] = t; }
        __syncthreads();
        const double dinv = 1.0 / A[c * Wd + c];
        __syncthreads();
        for (int j = tid; j < Wd; j += nt) A[c * Wd + j] *= dinv;
        __syncthreads();
        for (int t = tid; t < M * Wd; t += nt) {
            const int i = t / Wd, j = t - i * Wd;
            if (i != c && j != c) A[t] -= A[i * Wd + c] * A[c * Wd + j];
        }
        __syncthreads();
        for (int i = tid; i < M; i += nt) if (i != c) A[i * Wd + c] = 0.0;
        __syncthreads();
    }
    for (int t = tid; t < M * M; t += nt) { const int i = t / M, j = t - i * M; Cm[t] = A[i * Wd + M + j]; }
    for (int i = tid; i < M; i += nt) mud[i] = A[i * Wd + 2 * M];
    __syncthreads();
    if (alpha) {
        const double* Ki = Kx_inv + (size_t)d * M * M;
        for (int i = tid; i < M; i += nt) {
            double s = 0;
            for (int j = 0; j < M; ++j) s += Ki[i * M + j] * A[j * Wd + 2 * M];
            alpha[(size_t)d * M + i] = s;
        }
    }
}

}  // namespace qmpc
