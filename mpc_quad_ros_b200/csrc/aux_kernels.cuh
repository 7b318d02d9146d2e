// aux_kernels.cuh — the small per-vehicle kernels around the solve (fp64, one thread per vehicle or element).
#pragma once
#include "common.cuh"
#include "model.cuh"

namespace qmpc {

// quad_optimizer.set_reference_trajectory (reference src/quad_opt.py:295-317):
// yref[b][k] = [x_ref[b][k], u_ref[b][k] or 0.16], yref_e[b] = x_ref[b][N-1]
__global__ void set_reference_kernel(int B, int N, const double* __restrict__ x_ref, const double* __restrict__ u_ref,
                                     double u_hover, double* __restrict__ yref, double* __restrict__ yref_e)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = B * N * NY;
    if (t < total) {
        const int c = t % NY, bk = t / NY;
        yref[t] = c < NX ? x_ref[(size_t)bk * NX + c] : (u_ref ? u_ref[(size_t)bk * NU + (c - NX)] : u_hover);
    }
    if (t < B * NX) {
        const int c = t % NX, b = t / NX;
        yref_e[t] = x_ref[((size_t)b * N + (N - 1)) * NX + c];
    }
}

// A solve that ended with status != 0 (iteration limit or numerical breakdown) leaves an iterate that the next
// linearisation cannot use (the vehicle has usually departed from its reference).  Before the next solve the SQP
// iterate of such a vehicle is re-initialised on the new reference (states = reference, inputs = reference inputs)
// and its remembered active set is dropped.  The reference implementation ignores acados' status (quad_opt.py:333);
// this only acts where its behaviour is undefined.  A vehicle that keeps failing (two or more solves in a row: it has
// usually crashed) gets a bounded attempt per step (IpmArgs::max_iter_failed) until a solve succeeds again, so that it
// cannot hold up the other vehicles of its launch every step.
__global__ void reset_failed_kernel(int B, int N, const int* __restrict__ status, const double* __restrict__ yref,
                                    const double* __restrict__ yref_e, double* __restrict__ xit, double* __restrict__ uit,
                                    unsigned char* __restrict__ act, int* __restrict__ fail_streak)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * (N + 1)) return;
    const int b = t / (N + 1), k = t - b * (N + 1);
    if (k == 0) fail_streak[b] = status[b] == 0 ? 0 : fail_streak[b] + 1;      // consecutive failed solves of this vehicle
    if (status[b] == 0) return;
    double* x = xit + ((size_t)b * (N + 1) + k) * NX;
    if (k < N) {
        const double* y = yref + ((size_t)b * N + k) * NY;
        for (int c = 0; c < NX; ++c) x[c] = y[c];
        for (int c = 0; c < NU; ++c) { uit[((size_t)b * N + k) * NU + c] = y[NX + c]; act[((size_t)b * N + k) * NU + c] = 255; }
    } else {
        for (int c = 0; c < NX; ++c) x[c] = yref_e[(size_t)b * NX + c];
    }
}

// nominal RK4 step of the OCP model without GP (quad_optimizer.discrete_dynamics, quad_opt.py:353-377)
__device__ __forceinline__ void rk4_nominal(const ModelParams<double>& mp, const double* x, const double* u, double dt,
                                            double* xn)
{
    const double z3[3] = {0, 0, 0};
    double k[NX], xs[NX], acc[NX];
    EvalPoint<double> e;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double as = s == 3 ? dt : dt / 2;
        const double ws = (s == 0 || s == 3) ? 1.0 : 2.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) xs[i] = s == 0 ? x[i] : x[i] + as * k[i];
        eval_f(mp, xs, u, z3, z3, e, k);
#pragma unroll
        for (int i = 0; i < NX; ++i) acc[i] = s == 0 ? k[i] : acc[i] + ws * k[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = x[i] + dt / 6 * acc[i];    // x + dt/6 (k1 + 2k2 + 2k3 + k4)
}

__global__ void predict_nominal_kernel(ModelParams<double> mp, int B, const double* __restrict__ x,
                                       const double* __restrict__ u, double dt, int body_frame, double* __restrict__ xn)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xl[NX], ul[NU], out[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xl[i] = x[(size_t)b * NX + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) ul[i] = u[(size_t)b * NU + i];
    rk4_nominal(mp, xl, ul, dt, out);
    if (body_frame) {
        double vb[3];
        body_velocity(out, vb);
        out[7] = vb[0]; out[8] = vb[1]; out[9] = vb[2];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[(size_t)b * NX + i] = out[i];
}

// utils.compute_a_drag (utils.py:934-950)
__global__ void compute_a_drag_kernel(int B, const double* __restrict__ x_now, const double* __restrict__ x_pred,
                                      double dt, double* __restrict__ v_body, double* __restrict__ a_drag)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xa[NX], xb[NX], va[3], vp[3];
#pragma unroll
    for (int i = 0; i < NX; ++i) { xa[i] = x_now[(size_t)b * NX + i]; xb[i] = x_pred[(size_t)b * NX + i]; }
    body_velocity(xa, va);
    body_velocity(xb, vp);
#pragma unroll
    for (int i = 0; i < 3; ++i) { v_body[(size_t)b * 3 + i] = va[i]; a_drag[(size_t)b * 3 + i] = (va[i] - vp[i]) / dt; }
}

// After the solve of step i (execute_trajectory.py:212-214,251-255):
//   x_pred = nominal RK4(x_now, u0, dt);  residual of x_now against the prediction made one step earlier;
//   x_pred_prev <- x_pred.   xt/yt feed qrgp_regress_kernel.
__global__ void post_solve_kernel(ModelParams<double> mp, int B, double dt, int first_step,
                                  const double* __restrict__ x_now, const double* __restrict__ u0,
                                  double* __restrict__ x_pred_prev, double* __restrict__ xt, double* __restrict__ yt)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xa[NX], xb[NX], ul[NU], xp[NX], va[3], vp[3];
#pragma unroll
    for (int i = 0; i < NX; ++i) { xa[i] = x_now[(size_t)b * NX + i]; xb[i] = first_step ? xa[i] : x_pred_prev[(size_t)b * NX + i]; }
#pragma unroll
    for (int i = 0; i < NU; ++i) ul[i] = u0[(size_t)b * NU + i];
    rk4_nominal(mp, xa, ul, dt, xp);
    if (xt) {
        body_velocity(xa, va);
        body_velocity(xb, vp);
#pragma unroll
        for (int i = 0; i < 3; ++i) { xt[(size_t)b * 3 + i] = va[i]; yt[(size_t)b * 3 + i] = (va[i] - vp[i]) / dt; }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) x_pred_prev[(size_t)b * NX + i] = xp[i];
}

// utils.get_reference_chunk (utils.py:897-931), same idx for every vehicle
__global__ void reference_chunk_kernel(int B, int K, const double* __restrict__ traj, int idx, int N, int skip,
                                       double* __restrict__ chunk)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * N * NX) return;
    const int c = t % NX, k = (t / NX) % N, b = t / (NX * N);
    const int left = K - idx;
    int row;
    if (left > N * skip) row = idx + k * skip;
    else if (left > skip - 1) {
        const int n_left = (left + skip - 1) / skip;          // rows idx, idx+skip, ... < K
        row = k < n_left ? idx + k * skip : K - 1;
    } else row = K - 1;
    chunk[t] = traj[((size_t)b * K + row) * NX + c];
}

// ---- reference generation on the device: the chunk get_reference_chunk (utils.py:897-931) would cut out of
// a trajectory sampled at t = row * dt is evaluated analytically instead, so a closed loop never uploads references.
// q = (1,0,0,0) and r = 0 as in TrajectoryGenerator.load_trajectory (TrajectoryGenerator.py:237-238).
// par[b][REFGEN_NPAR], by kind:
//   0 sum of three sinusoids per axis (BASELINE config 2): amp[3][3], f[3][3] (Hz), phase[3][3], scale, p0[3], z0
//   1 lemniscate (config 5): phase, yaw, w, a, z0, p0x, p0y, ramp time T (angular rate ramps 0 -> w over T)
//   2 accelerating circle (config 1, TrajectoryGenerator.sample_circle_trajectory_accelerating :41-74): radius, v_max,
//     n (samples of the whole trajectory), start[3], csv flag (values rounded to 6 decimals, the reference's '%.6f' file)
constexpr int REFGEN_NPAR = 32;

__device__ __forceinline__ double csv6(double v) { return rint(v * 1e6) / 1e6; }

__global__ void reference_generate_kernel(int kind, int B, const double* __restrict__ par, int K, int idx, int N, int skip,
                                          double dt, double* __restrict__ chunk)
{
    const int t_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= B * N) return;
    const int k = t_ % N, b = t_ / N;
    const int left = K - idx;
    int row;                                                   // same row rule as reference_chunk_kernel
    if (left > N * skip) row = idx + k * skip;
    else if (left > skip - 1) {
        const int n_left = (left + skip - 1) / skip;
        row = k < n_left ? idx + k * skip : K - 1;
    } else row = K - 1;
    const double* q = par + (size_t)b * REFGEN_NPAR;
    const double t = row * dt;
    const double two_pi = 6.283185307179586;
    double p[3], v[3];
    if (kind == 0) {
        const double s = q[27];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double ps = 0, vs = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double amp = q[a * 3 + j], f = q[9 + a * 3 + j], arg = two_pi * f * t + q[18 + a * 3 + j];
                double sn, cs;
                sincos(arg, &sn, &cs);
                ps += amp * sn;
                vs += amp * two_pi * f * cs;
            }
            p[a] = ps * s - q[28 + a];
            v[a] = vs * s;
        }
        p[2] += q[31];
    } else if (kind == 1) {
        const double ph = q[0], yaw = q[1], w = q[2], a = q[3], T = q[7];
        const double th = ph + w * (t < T ? t * t / (2.0 * T) : t - 0.5 * T);
        const double dth = w * (t < T ? t / T : 1.0);
        double sth, cth, sy, cy;
        sincos(th, &sth, &cth);
        sincos(yaw, &sy, &cy);
        const double px = a * sth, py = a * sth * cth, vx = a * cth * dth, vy = a * cos(2.0 * th) * dth;
        p[0] = cy * px - sy * py - q[5]; p[1] = sy * px + cy * py - q[6]; p[2] = q[4];
        v[0] = cy * vx - sy * vy; v[1] = sy * vx + cy * vy; v[2] = 0.0;
    } else {
        const double radius = q[0], w_max = q[1] / q[0], n = q[2];
        // w_m = w_max (sin(pi k_m + 3 pi / 4) + 1) / 2 with k_m = 2 (m + 1) / n - 1; phi_row = dt * sum_{m <= row} w_m in closed form
        const double d = two_pi / n, a0 = 0.5 * two_pi * (2.0 / n - 1.0) + 0.75 * 0.5 * two_pi;
        const double S = sin(a0 + 0.5 * row * d) * sin(0.5 * (row + 1) * d) / sin(0.5 * d);
        const double phi = w_max * dt * 0.5 * ((row + 1) + S);
        const double w = w_max * 0.5 * (sin(a0 + row * d) + 1.0);
        double sp, cp;
        sincos(phi, &sp, &cp);
        p[0] = radius * cp - radius + q[3]; p[1] = radius * sp + q[4]; p[2] = q[5];
        v[0] = -radius * w * sp; v[1] = radius * w * cp; v[2] = 0.0;
        if (q[6] != 0.0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { p[a] = csv6(p[a]); v[a] = csv6(v[a]); }
        }
    }
    double* o = chunk + (size_t)t_ * NX;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    o[3] = 1.0; o[4] = 0.0; o[5] = 0.0; o[6] = 0.0;
    o[7] = v[0]; o[8] = v[1]; o[9] = v[2];
    o[10] = 0.0; o[11] = 0.0; o[12] = 0.0;
}

// Quadrotor3D.update (quad.py:234-277,305-381): RK4 of the nominal model + aero/rotor drag, inputs clipped to [0,1]
struct PlantParams { double aero, rotor[3], mass; };

__device__ __forceinline__ void plant_f(const ModelParams<double>& mp, const PlantParams& pp, const double* x,
                                        const double* u, double* f)
{
    const double z3[3] = {0, 0, 0};
    EvalPoint<double> e;
    eval_f(mp, x, u, z3, z3, e, f);
    double ad[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double vb = e.R[i] * e.v[0] + e.R[3 + i] * e.v[1] + e.R[6 + i] * e.v[2];
        const double sg = (vb > 0) - (vb < 0);
        ad[i] = -pp.aero * vb * vb * sg / pp.mass - pp.rotor[i] * vb / pp.mass;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) f[7 + i] += e.R[3 * i] * ad[0] + e.R[3 * i + 1] * ad[1] + e.R[3 * i + 2] * ad[2];
}

__global__ void plant_period_kernel(ModelParams<double> mp, PlantParams pp, int B, double* __restrict__ x,
                                    const double* __restrict__ u_in, double sim_dt, int n_sub)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xl[NX], u[NU], k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xl[i] = x[(size_t)b * NX + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) { const double v = u_in[(size_t)b * NU + i]; u[i] = v < 0 ? 0 : (v > 1 ? 1 : v); }
    for (int s = 0; s < n_sub; ++s) {
        plant_f(mp, pp, xl, u, k1);
#pragma unroll
        for (int i = 0; i < NX; ++i) xs[i] = xl[i] + sim_dt / 2 * k1[i];
        plant_f(mp, pp, xs, u, k2);
#pragma unroll
        for (int i = 0; i < NX; ++i) xs[i] = xl[i] + sim_dt / 2 * k2[i];
        plant_f(mp, pp, xs, u, k3);
#pragma unroll
        for (int i = 0; i < NX; ++i) xs[i] = xl[i] + sim_dt * k3[i];
        plant_f(mp, pp, xs, u, k4);
#pragma unroll
        for (int i = 0; i < NX; ++i) xl[i] += sim_dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) x[(size_t)b * NX + i] = xl[i];
}

// 16 independent FMA chains per thread, all in registers (roofline denominator for the FMA-bound kernels)
template <typename real>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, double* sink)
{
    real a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = real(threadIdx.x + i) * real(1e-3);
    const real b = real(0.999), c = real(1e-6);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = a[i] * b + c;
    }
    real s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == real(-1)) *sink = double(s);
}

}  // namespace qmpc
