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

// The same update with the covariance staged through shared memory by TMA bulk copies (M even, M <= RGP_TMA_MAXM): lane 0
// starts one cp.async.bulk of the model's C (8 M^2 bytes, contiguous) onto an mbarrier, the warp computes kv / Jt meanwhile,
// forms w = Jt C (lane = column) and cj = C Jt^T (lane = row) from shared memory without a single shuffle, applies the
// rank-1 update in place and sends C back with one bulk store while mu and alpha are finished.  C crosses HBM exactly once
// in each direction (16 M^2 bytes per model, the algorithmic figure) and all 32 lanes take part in the element-wise pass.
// dynamic smem per warp: M*M + 4*M + 2 doubles (C, kv, Jt, cj, w, mbarrier)
constexpr int RGP_TMA_MAXM = 64;
__host__ __device__ constexpr int rgp_tma_reals(int M) { return M * M + 4 * M + 2; }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) qrgp_regress_tma_kernel(RgpArgs a)
{
    QMPC_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int model = blockIdx.x * WARPS + warp;           // (vehicle, axis)
    if (model >= a.B * 3) return;
    const int M = a.M, MM = M * M, d = model % 3;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * rgp_tma_reals(M);
    double *Cs = sm, *kv = sm + MM, *Jt = kv + M, *cj = Jt + M, *wv = cj + M;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(wv + M);
    const double L = a.theta[3 * d], sf = a.theta[3 * d + 1], sn = a.theta[3 * d + 2];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    const double* X = a.X + d * M;
    const double* Ki = a.Kx_inv + (size_t)d * M * M;
    double* mu = a.mu + (size_t)model * M;
    double* Cm = a.C + (size_t)model * MM;
    const double xt = a.xt[model], yt = a.yt[model];
    if (!(xt - xt == 0.0) || !(yt - yt == 0.0)) return;    // non-finite sample = "no update for this axis"
    const unsigned bytes = (unsigned)(MM * sizeof(double));
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect(bar, bytes);
        bulk_g2s(Cs, Cm, bytes, bar);
    }
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
    bulk_wait_warp(bar, 0);
    double jcj = 0;
    for (int j = lane; j < M; j += 32) {
        double w = 0, c = 0;
        for (int i = 0; i < M; ++i) w += Jt[i] * Cs[i * M + j];             // w_j  = (Jt C)_j     : row reads, conflict-free
        for (int i = 0; i < M; ++i) c += Cs[j * M + i] * Jt[i];             // cj_j = (C Jt^T)_j   : lane j walks its own row
        wv[j] = w; cj[j] = c;
        jcj += w * Jt[j];
    }
    jcj = warp_sum(jcj);
    __syncwarp();
    const double b = rbf_k(xt, xt, iL2, sf2) - jk;
    const double sinv = 1.0 / (b + jcj + sn * sn);
    const double innov = yt - mp;
    {   // C -= (cj / s) w, every lane 1/32 of the elements; (i, j) of element e = lane + 32 t advance without a division
        int i = lane / M, j = lane - i * M;
        const int di = 32 / M, dj = 32 - di * M;
        for (int e = lane; e < MM; e += 32) {
            Cs[e] -= cj[i] * sinv * wv[j];
            i += di; j += dj;
            if (j >= M) { j -= M; ++i; }
        }
    }
    fence_proxy_async();                                   // this lane's writes to Cs, before the async proxy reads them
    __syncwarp();
    if (lane == 0) bulk_s2g(Cm, Cs, bytes);
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
    if (lane == 0) bulk_s2g_wait_read();                   // the shared buffer must outlive the store's read
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

// RGP.predict(cov=True, return_Jt=True) (RGP.py:195-229): gain rows Jt [m][M] and the full posterior covariance
//   C_p = K(X*,X*) - Jt K(X,X*) + Jt C_g Jt^T   [m][m]   per (vehicle, axis).  One warp per model; off the control path.
struct RgpCovArgs {
    int B, M, m;
    const double* X; const double* theta; const double* Kx_inv;
    const double* C;       // [B][3][M][M]
    const double* xs;      // [B][3][m]
    double* Jt;            // [B][3][m][M]   (output, also the workspace of the covariance pass)
    double* cov;           // [B][3][m][m]
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) qrgp_predict_cov_kernel(RgpCovArgs a)
{
    QMPC_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int model = blockIdx.x * WARPS + warp;
    if (model >= a.B * 3) return;
    const int M = a.M, m = a.m, d = model % 3;
    double* kv = reinterpret_cast<double*>(smem_raw) + (size_t)warp * M;
    const double L = a.theta[3 * d], sf = a.theta[3 * d + 1];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    const double* X = a.X + d * M;
    const double* Ki = a.Kx_inv + (size_t)d * M * M;
    const double* Cm = a.C + (size_t)model * M * M;
    const double* xs = a.xs + (size_t)model * m;
    double* Jt = a.Jt + (size_t)model * m * M;
    for (int q = 0; q < m; ++q) {
        __syncwarp();
        for (int i = lane; i < M; i += 32) kv[i] = rbf_k(xs[q], X[i], iL2, sf2);
        __syncwarp();
        for (int j = lane; j < M; j += 32) {
            double s = 0;
            for (int i = 0; i < M; ++i) s += kv[i] * __ldg(Ki + (size_t)i * M + j);
            Jt[(size_t)q * M + j] = s;
        }
    }
    __syncwarp();
    for (int e = lane; e < m * m; e += 32) {
        const int q = e / m, r = e - q * m;
        const double* Jq = Jt + (size_t)q * M;
        const double* Jr = Jt + (size_t)r * M;
        double jk = 0, jcj = 0;
        for (int i = 0; i < M; ++i) {
            jk += Jq[i] * rbf_k(X[i], xs[r], iL2, sf2);
            double t = 0;
            for (int j = 0; j < M; ++j) t += Cm[(size_t)i * M + j] * Jr[j];
            jcj += Jq[i] * t;
        }
        a.cov[(size_t)model * m * m + e] = rbf_k(xs[q], xs[r], iL2, sf2) - jk + jcj;
    }
}

// ------------------------------------------------------------------ shared-swarm (information form)

struct RgpSharedArgs {
    int B, M;
    const double* X; const double* theta; const double* Kx_inv;
    const double* xt;   // [B][3]
    const double* yt;   // [B][3]
    double* part;       // [3][gridDim.x][M*M + M]  per-CTA partial sums (Lambda row-major, then eta)
};

// Each warp walks a strided subset of vehicles for one axis and accumulates j^T j / r and j^T y / r into a
// warp-private shared-memory tile; the CTA adds its warps' tiles in warp order and writes ONE partial block, which
// qrgp_shared_reduce_kernel sums over CTAs in CTA order: no atomics anywhere, so the result is bit-reproducible.
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
    __syncthreads();
    double* out = a.part + ((size_t)d * gridDim.x + blockIdx.x) * (M * M + M);
    const double* base = reinterpret_cast<double*>(smem_raw);
    for (int t = threadIdx.x; t < M * M + M; t += WARPS * 32) {
        double s = 0;
        for (int w = 0; w < WARPS; ++w) s += base[(size_t)w * (M * M + 3 * M) + t];
        out[t] = s;
    }
}

// info[d][t] = sum over the `ctas` partial blocks of axis d, in CTA order
__global__ void qrgp_shared_reduce_kernel(int len, int ctas, int stride_ctas, const double* __restrict__ part, double* __restrict__ info)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
    if (t >= len) return;
    const double* p = part + (size_t)d * ctas * len + t;
    double s = 0;
    for (int c = 0; c < ctas; ++c) s += p[(size_t)c * len];
    info[(size_t)d * len + t] = s;
    (void)stride_ctas;
}

// C0 = K_x for every vehicle and axis (RGP.py:144): C[b][e] = Kx[e]
__global__ void qrgp_fill_C_kernel(long long B, int per, const double* __restrict__ Kx, double* __restrict__ C)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * per) C[t] = Kx[t % per];
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

// ------------------------------------------------------------------------------------------ RGP* learning
// RGP.learn (reference src/gp/RGP.py:332-482) with its sigma points (__draw_sigma_points :485-505): joint recursive
// update of the basis-point estimate g and the hyper-parameters eta = (L, sigma_f, sigma_n) from one sample, restated
// as the reference computes it, peculiarities included: At (hence Jt) is built once at the current hyper-parameters
// and reused for every sigma point (:357-366, :392); the running mean of the cumulative sum is used inside the outer
// product of the same iteration (:403-404); C_g_eta_t is never assigned by learn (:153); the exp() transform of eta
// (:468-470) is overwritten by the plain assignment (:472-474).
// One CTA per 1-D model; the (M+4)^2 joint covariance, the M x 2M Gauss-Jordan tableau of the new K_x and the vectors
// live in shared memory, so M <= 64.
struct RgpLearnArgs {
    int n_models, M;
    const double* X;        // [M] basis points (shared grid)
    double* mu_g;           // [n][M]
    double* C_g;            // [n][M][M]
    double* mu_eta;         // [n][3]
    double* C_eta;          // [n][3][3]
    const double* C_g_eta;  // [n][M][3]   (never written by learn, like the reference)
    double* Kx_inv;         // [n][M][M]
    double* mu_z;           // [n][M+3]        return value of learn (may be null)
    double* C_z;            // [n][M+3][M+3]   (may be null)
    const double* xt;       // [n]
    const double* yt;       // [n]
    int* status;            // [n] 0 ok, 1 singular K_x
};

__host__ __device__ inline int rgp_learn_smem_doubles(int M) { return (M + 4) * (M + 4) + 3 * M * M + 14 * M + 64; }

__device__ inline void sym3_sqrt_dev(const double* A, double* S)     // principal square root, cyclic Jacobi (thread 0)
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (A[i * 3 + j] + A[j * 3 + i]);
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += v[i][k] * sqrt(a[k][k]) * v[j][k];
            S[i * 3 + j] = acc;
        }
}

// in-place Gauss-Jordan with partial pivoting on the n x 2n tableau W = [A | I] in shared memory (whole CTA)
__device__ inline int gj_inverse_cta(int n, double* W, int* piv_sh)
{
    const int tid = threadIdx.x, nt = blockDim.x, w2 = 2 * n;
    for (int c = 0; c < n; ++c) {
        if (tid == 0) {
            int piv = c;
            for (int i = c + 1; i < n; ++i) if (fabs(W[i * w2 + c]) > fabs(W[piv * w2 + c])) piv = i;
            piv_sh[0] = W[piv * w2 + c] == 0.0 ? -1 : piv;
        }
        __syncthreads();
        const int piv = piv_sh[0];
        if (piv < 0) return 1;
        if (piv != c) for (int j = tid; j < w2; j += nt) { const double t = W[c * w2 + j]; W[c * w2 + j] = W[piv * w2 + j]; W[piv * w2 + j] = t; }
        __syncthreads();
        const double d = 1.0 / W[c * w2 + c];
        __syncthreads();
        for (int j = tid; j < w2; j += nt) W[c * w2 + j] *= d;
        __syncthreads();
        // eliminate column c from the other rows; the multipliers are read before anybody overwrites column c
        double* fcol = W + n * w2;              // n spare doubles behind the tableau
        for (int i = tid; i < n; i += nt) fcol[i] = i == c ? 0.0 : W[i * w2 + c];
        __syncthreads();
        for (int e = tid; e < n * w2; e += nt) {
            const int i = e / w2, j = e - i * w2;
            const double f = fcol[i];
            if (f != 0.0) W[e] -= f * W[c * w2 + j];
        }
        __syncthreads();
    }
    return 0;
}

__global__ void __launch_bounds__(128) qrgp_learn_kernel(RgpLearnArgs a)
{
    extern __shared__ __align__(16) double lsm[];
    const int model = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, M = a.M;
    if (model >= a.n_models) return;
    const int np_ = M + 4, nu_ = M + 2, nz = M + 3;
    double* C_p = lsm;                         // (M+4)^2
    double* Cgp = C_p + np_ * np_;             // M^2
    double* W = Cgp + M * M;                   // M x 2M tableau (+ M spare)
    double* kv = W + 2 * M * M + M;
    double* Jt = kv + M; double* JC = Jt + M; double* cj = JC + M; double* vg = cj + M;
    double* mu_p = vg + M; double* mu_pi = mu_p + np_; double* Lt = mu_pi + np_;       // Lt: 2(M+2)
    double* St = Lt + 2 * nu_;                 // M x 3
    double* sc = St + 3 * M;                   // scalars: [0..8] Ssq, [9..17] Cei, 18 Bv, 19 jcj, 20 jg
    __shared__ int piv_sh[1];
    const double* X = a.X;
    double* mu_g = a.mu_g + (size_t)model * M;
    double* C_g = a.C_g + (size_t)model * M * M;
    double* mu_eta = a.mu_eta + (size_t)model * 3;
    double* C_eta = a.C_eta + (size_t)model * 9;
    const double* C_g_eta = a.C_g_eta + (size_t)model * M * 3;
    double* Kxi = a.Kx_inv + (size_t)model * M * M;
    const double xt = a.xt[model], yt = a.yt[model];
    const double L = mu_eta[0], sf = mu_eta[1];
    const double iL2 = 1.0 / (L * L), sf2 = sf * sf;
    // ---- inference step
    for (int i = tid; i < M; i += nt) kv[i] = rbf_k(xt, X[i], iL2, sf2);
    for (int e = tid; e < np_ * np_; e += nt) C_p[e] = 0.0;
    for (int i = tid; i < np_; i += nt) mu_p[i] = 0.0;
    __syncthreads();
    for (int j = tid; j < M; j += nt) { double s = 0; for (int i = 0; i < M; ++i) s += kv[i] * Kxi[i * M + j]; Jt[j] = s; }
    if (tid == 0) {
        double Ce[9], Wk[18];
        for (int i = 0; i < 9; ++i) Ce[i] = C_eta[i];
        // 3x3 inverse, Gauss-Jordan with partial pivoting
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Wk[i * 6 + j] = Ce[i * 3 + j]; Wk[i * 6 + 3 + j] = (i == j); }
        for (int c = 0; c < 3; ++c) {
            int piv = c;
            for (int i = c + 1; i < 3; ++i) if (fabs(Wk[i * 6 + c]) > fabs(Wk[piv * 6 + c])) piv = i;
            if (piv != c) for (int j = 0; j < 6; ++j) { const double t = Wk[c * 6 + j]; Wk[c * 6 + j] = Wk[piv * 6 + j]; Wk[piv * 6 + j] = t; }
            const double d = 1.0 / Wk[c * 6 + c];
            for (int j = 0; j < 6; ++j) Wk[c * 6 + j] *= d;
            for (int i = 0; i < 3; ++i) if (i != c) { const double f = Wk[i * 6 + c]; if (f != 0) for (int j = 0; j < 6; ++j) Wk[i * 6 + j] -= f * Wk[c * 6 + j]; }
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sc[9 + i * 3 + j] = Wk[i * 6 + 3 + j];
        double S6[9];
        for (int i = 0; i < 9; ++i) S6[i] = (3.0 / (1.0 - 0.5)) * Ce[i];
        sym3_sqrt_dev(S6, sc);
    }
    __syncthreads();
    if (tid == 0) {
        double jk = 0;
        for (int j = 0; j < M; ++j) jk += Jt[j] * rbf_k(X[j], xt, iL2, sf2);
        sc[18] = rbf_k(xt, xt, iL2, sf2) - jk;
    }
    for (int e = tid; e < M * 3; e += nt) {
        const int i = e / 3, j = e - i * 3;
        double s = 0;
        for (int k = 0; k < 3; ++k) s += C_g_eta[i * 3 + k] * sc[9 + k * 3 + j];
        St[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < M * M; e += nt) {
        const int i = e / M, j = e - i * M;
        double s = 0;
        for (int k = 0; k < 3; ++k) s += St[i * 3 + k] * C_g_eta[j * 3 + k];
        Cgp[e] = C_g[e] - s;
    }
    __syncthreads();
    for (int j = tid; j < M; j += nt) {
        double s = 0, t = 0;
        for (int i = 0; i < M; ++i) { s += Jt[i] * Cgp[i * M + j]; t += Cgp[j * M + i] * Jt[i]; }
        JC[j] = s; cj[j] = t;
    }
    __syncthreads();
    if (tid == 0) { double s = 0; for (int j = 0; j < M; ++j) s += JC[j] * Jt[j]; sc[19] = s; }
    __syncthreads();
    // ---- unscented transform: 7 sigma points of eta, cumulative mean / covariance exactly in the reference's order
    for (int sp = 0; sp < 7; ++sp) {
        const double w = sp == 0 ? 0.5 : (1.0 - 0.5) / 6.0;
        double eta[3];
        for (int k = 0; k < 3; ++k)
            eta[k] = sp == 0 ? mu_eta[k] : (sp <= 3 ? mu_eta[k] + sc[k * 3 + (sp - 1)] : mu_eta[k] - sc[k * 3 + (sp - 4)]);
        for (int i = tid; i < M; i += nt) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += St[i * 3 + k] * (eta[k] - mu_eta[k]);
            vg[i] = mu_g[i] + s;
        }
        __syncthreads();
        if (tid == 0) { double s = 0; for (int i = 0; i < M; ++i) s += Jt[i] * vg[i]; sc[20] = s; }
        __syncthreads();
        for (int i = tid; i < np_; i += nt) {
            const double v = i < M ? vg[i] : (i < M + 3 ? eta[i - M] : sc[20]);
            mu_pi[i] = v;
            mu_p[i] += w * v;
        }
        __syncthreads();
        for (int e = tid; e < np_ * np_; e += nt) {
            const int i = e / np_, j = e - i * np_;
            double cpi = 0;
            if (i < M && j < M) cpi = Cgp[i * M + j];
            else if (i < M && j == M + 3) cpi = cj[i];
            else if (i == M + 3 && j < M) cpi = JC[j];
            else if (i == M + 3 && j == M + 3) cpi = sc[19] + sc[18];
            C_p[e] += w * ((mu_pi[i] - mu_p[i]) * (mu_pi[j] - mu_p[j]) + cpi);
        }
        __syncthreads();
    }
    // ---- update step: o = [sigma_n, g_t] (rows M+2, M+3), u = [g, L, sigma_f]
    const int o0 = M + 2, o1 = M + 3;
    const double mo0 = mu_p[o0], mo1 = mu_p[o1];
    const double Co00 = C_p[o0 * np_ + o0], Co01 = C_p[o0 * np_ + o1], Co10 = C_p[o1 * np_ + o0], Co11 = C_p[o1 * np_ + o1];
    const double Cy = Co11 + Co00 + mo0 * mo0;
    const double G0 = Co01 / Cy, G1 = Co11 / Cy;
    const double me0 = mo0 + G0 * (yt - mo1), me1 = mo1 + G1 * (yt - mo1);
    const double Ce00 = Co00 - G0 * Cy * G0, Ce01 = Co01 - G0 * Cy * G1, Ce10 = Co10 - G1 * Cy * G0, Ce11 = Co11 - G1 * Cy * G1;
    const double det = Co00 * Co11 - Co01 * Co10;
    const double Ci00 = Co11 / det, Ci01 = -Co01 / det, Ci10 = -Co10 / det, Ci11 = Co00 / det;
    for (int i = tid; i < nu_; i += nt) {
        const double c0 = C_p[o0 * np_ + i], c1 = C_p[o1 * np_ + i];
        Lt[i * 2] = c0 * Ci00 + c1 * Ci10;
        Lt[i * 2 + 1] = c0 * Ci01 + c1 * Ci11;
    }
    __syncthreads();
    const double d0 = me0 - mo0, d1 = me1 - mo1;
    const double D00 = Ce00 - Co00, D01 = Ce01 - Co01, D10 = Ce10 - Co10, D11 = Ce11 - Co11;
    double* mu_z = a.mu_z ? a.mu_z + (size_t)model * nz : nullptr;
    double* C_z = a.C_z ? a.C_z + (size_t)model * nz * nz : nullptr;
    for (int i = tid; i < nz; i += nt) {
        const double v = i < nu_ ? mu_p[i] + Lt[i * 2] * d0 + Lt[i * 2 + 1] * d1 : me0;
        if (mu_z) mu_z[i] = v;
        if (i < M) mu_g[i] = v; else mu_eta[i - M] = v;
    }
    for (int e = tid; e < nz * nz; e += nt) {
        const int i = e / nz, j = e - i * nz;
        double v;
        if (i < nu_ && j < nu_) {
            const double t0 = Lt[i * 2] * D00 + Lt[i * 2 + 1] * D10, t1 = Lt[i * 2] * D01 + Lt[i * 2 + 1] * D11;
            v = C_p[i * np_ + j] + t0 * Lt[j * 2] + t1 * Lt[j * 2 + 1];
        } else if (i < nu_) v = Lt[i * 2] * Ce00 + Lt[i * 2 + 1] * Ce10;
        else if (j < nu_) v = Ce00 * Lt[j * 2] + Ce01 * Lt[j * 2 + 1];
        else v = Ce00;
        if (C_z) C_z[e] = v;
        if (i < M && j < M) C_g[i * M + j] = v;
        else if (i >= M && j >= M) C_eta[(i - M) * 3 + (j - M)] = v;
    }
    __syncthreads();
    // ---- the pre-computed matrices follow the new hyper-parameters (RGP.py:476-479)
    __threadfence_block();
    const double Ln = mu_eta[0], sfn = mu_eta[1], snn = mu_eta[2];
    const double iL2n = 1.0 / (Ln * Ln), sf2n = sfn * sfn;
    for (int e = tid; e < M * 2 * M; e += nt) {
        const int i = e / (2 * M), j = e - i * 2 * M;
        W[e] = j < M ? rbf_k(X[i], X[j], iL2n, sf2n) + (i == j ? snn * snn : 0.0) : (j - M == i ? 1.0 : 0.0);
    }
    __syncthreads();
    const int rc = gj_inverse_cta(M, W, piv_sh);
    if (rc == 0) for (int e = tid; e < M * M; e += nt) { const int i = e / M, j = e - i * M; Kxi[e] = W[i * 2 * M + M + j]; }
    if (tid == 0 && a.status) a.status[model] = rc;
}

}  // namespace qmpc
