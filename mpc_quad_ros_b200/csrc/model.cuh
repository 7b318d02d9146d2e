// model.cuh — quadrotor model of the OCP, device side.
//
// 13-state quaternion model with 4 motor inputs and optional RGP drag augmentation
// (reference src/quad_opt.py:164-262; rotation helpers src/utils/utils.py:317-340,394-440;
//  RGP mean src/gp/RGP.py:250-254 with the constant product K_x^-1 p folded into alpha).
// Point evaluation + hand-written forward-mode derivative so that the RK4 forward sensitivities are propagated column by
// column, one column per lane; what all columns share (the velocity-row Jacobian) is formed once per evaluated point.
#pragma once
#include "common.cuh"

namespace qmpc {

constexpr int NX = 13;
constexpr int NU = 4;
constexpr int NY = 17;

template <typename real>
struct ModelParams {
    real thrust_over_mass;     // T / m
    real T;                    // max thrust per rotor
    real xf[4], yf[4], zt[4];  // rotor arms / yaw-torque arms
    real invJ[3];
    real Jc[3];                // (J1-J2), (J2-J0), (J0-J1)
    real g[3];
    real sf2[3], iL2[3];       // RGP kernel: sigma_f^2, 1/L^2 per axis
    real gx0[3], gdx[3], gidx[3], gcc[3];   // equispaced basis points of an axis (linspace, GPE.fromrange): first point, spacing
                               // (0: not equispaced - every kernel value takes its own exp), 1/spacing, exp(-spacing^2 / L^2)
    int M;                     // basis points per axis (0 = nominal)
};

template <typename real> __device__ __forceinline__ real rexp(real x);
template <> __device__ __forceinline__ double rexp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float rexp<float>(float x) { return expf(x); }

// everything the JVP re-uses from one evaluation point
template <typename real>
struct EvalPoint {
    real q[4], v[3], r[3];
    real R[9];
    real ab[3];   // body-frame specific force: (mu0, mu1, aT + mu2)
    real dmu[3];  // d mu_d / d v_b,d
};

template <typename real>
__device__ __forceinline__ void rotmat(const real* q, real* R)
{
    const real w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = real(1) - real(2) * (y * y + z * z); R[1] = real(2) * (x * y - w * z); R[2] = real(2) * (x * z + w * y);
    R[3] = real(2) * (x * y + w * z); R[4] = real(1) - real(2) * (x * x + z * z); R[5] = real(2) * (y * z - w * x);
    R[6] = real(2) * (x * z - w * y); R[7] = real(2) * (y * z + w * x); R[8] = real(1) - real(2) * (x * x + y * y);
}

// Evaluate f at (x,u).  mu/dmu: GP mean and slope per body axis, already reduced over the basis points
// (the caller distributes the 3*M kernel evaluations over its lane group); pass zeros for the nominal model.
template <typename real>
__device__ __forceinline__ void eval_f(const ModelParams<real>& mp, const real* x, const real* u,
                                       const real* mu, const real* dmu, EvalPoint<real>& e, real* f)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) e.q[i] = x[3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { e.v[i] = x[7 + i]; e.r[i] = x[10 + i]; e.dmu[i] = dmu[i]; }
    rotmat(e.q, e.R);
    const real aT = mp.thrust_over_mass * (u[0] + u[1] + u[2] + u[3]);
    e.ab[0] = mu[0]; e.ab[1] = mu[1]; e.ab[2] = aT + mu[2];
    const real* q = e.q; const real* r = e.r;
    f[0] = e.v[0]; f[1] = e.v[1]; f[2] = e.v[2];
    f[3] = real(0.5) * (-r[0] * q[1] - r[1] * q[2] - r[2] * q[3]);
    f[4] = real(0.5) * (r[0] * q[0] + r[2] * q[2] - r[1] * q[3]);
    f[5] = real(0.5) * (r[1] * q[0] - r[2] * q[1] + r[0] * q[3]);
    f[6] = real(0.5) * (r[2] * q[0] + r[1] * q[1] - r[0] * q[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        f[7 + i] = e.R[3 * i] * e.ab[0] + e.R[3 * i + 1] * e.ab[1] + e.R[3 * i + 2] * e.ab[2] - mp.g[i];
    real ty = 0, tx = 0, tz = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { ty += u[i] * mp.yf[i]; tx += u[i] * mp.xf[i]; tz += u[i] * mp.zt[i]; }
    f[10] = (mp.T * ty + mp.Jc[0] * r[1] * r[2]) * mp.invJ[0];
    f[11] = (-mp.T * tx + mp.Jc[1] * r[2] * r[0]) * mp.invJ[1];
    f[12] = (mp.T * tz + mp.Jc[2] * r[0] * r[1]) * mp.invJ[2];
}

// body-frame velocity v_b = R(q)^T v (= v_dot_q(v, quaternion_inverse(q)), utils.py:317,434)
template <typename real>
__device__ __forceinline__ void body_velocity(const real* x, real* vb)
{
    real R[9];
    rotmat(x + 3, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) vb[i] = R[i] * x[7] + R[3 + i] * x[8] + R[6 + i] * x[9];
}

// Jacobian blocks of the velocity rows of f at an evaluated point; they are the same for every tangent column:
//   d(v_dot) = Cq dq + Cv dv + R[:,2] (T/m) sum(du)
//   Cq = d(R ab)/dq + R D d(R^T v)/dq   [3][4],    Cv = R D R^T   [3][3],    D = diag(dmu)  (GP slopes; 0 for the nominal model)
// R(q) is the un-normalised rotation matrix of rotmat(), so dR/dq_c has entries 0, +-2 q_k and -4 q_k.
template <typename real>
__device__ __forceinline__ void velocity_jacobian(const EvalPoint<real>& e, real* Cq, real* Cv)
{
    const real w = e.q[0], x = e.q[1], y = e.q[2], z = e.q[3];
    const real w2 = real(2) * w, x2 = real(2) * x, y2 = real(2) * y, z2 = real(2) * z, o = real(0);
    const real dRc[4][9] = {{o, -z2, y2, z2, o, -x2, -y2, x2, o},
                            {o, y2, z2, y2, -real(2) * x2, -w2, z2, w2, -real(2) * x2},
                            {-real(2) * y2, x2, w2, x2, o, z2, -w2, z2, -real(2) * y2},
                            {-real(2) * z2, -w2, x2, w2, -real(2) * z2, y2, x2, y2, o}};
    real RD[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int m = 0; m < 3; ++m) RD[3 * i + m] = e.R[3 * i + m] * e.dmu[m];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        real g[3], h[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g[i] = dRc[c][3 * i] * e.ab[0] + dRc[c][3 * i + 1] * e.ab[1] + dRc[c][3 * i + 2] * e.ab[2];      // (dR/dq_c) ab
            h[i] = dRc[c][i] * e.v[0] + dRc[c][3 + i] * e.v[1] + dRc[c][6 + i] * e.v[2];                      // (dR/dq_c)^T v
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) Cq[4 * i + c] = g[i] + RD[3 * i] * h[0] + RD[3 * i + 1] * h[1] + RD[3 * i + 2] * h[2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int m = 0; m < 3; ++m) Cv[3 * i + m] = RD[3 * i] * e.R[3 * m] + RD[3 * i + 1] * e.R[3 * m + 1] + RD[3 * i + 2] * e.R[3 * m + 2];
}

// df = (df/dx) dx + (df/du) e_j for ONE tangent column at a point whose velocity-row Jacobian is cached:
//   uz = R[:,2] T/m if the column is an input direction (else 0), tq = torque-row constants of that input (else 0),
//   kr = (J1-J2)/J0, (J2-J0)/J1, (J0-J1)/J2
template <typename real>
__device__ __forceinline__ void jvp_cached(const real* q, const real* r, const real* Cq, const real* Cv, const real* uz,
                                           const real* kr, const real* tq, const real* dx, real* df)
{
    const real* dq = dx + 3; const real* dv = dx + 7; const real* dr = dx + 10;
    df[0] = dv[0]; df[1] = dv[1]; df[2] = dv[2];
    df[3] = real(0.5) * (-dr[0] * q[1] - dr[1] * q[2] - dr[2] * q[3] - r[0] * dq[1] - r[1] * dq[2] - r[2] * dq[3]);
    df[4] = real(0.5) * (dr[0] * q[0] + dr[2] * q[2] - dr[1] * q[3] + r[0] * dq[0] + r[2] * dq[2] - r[1] * dq[3]);
    df[5] = real(0.5) * (dr[1] * q[0] - dr[2] * q[1] + dr[0] * q[3] + r[1] * dq[0] - r[2] * dq[1] + r[0] * dq[3]);
    df[6] = real(0.5) * (dr[2] * q[0] + dr[1] * q[1] - dr[0] * q[2] + r[2] * dq[0] + r[1] * dq[1] - r[0] * dq[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        df[7 + i] = uz[i] + Cq[4 * i] * dq[0] + Cq[4 * i + 1] * dq[1] + Cq[4 * i + 2] * dq[2] + Cq[4 * i + 3] * dq[3]
                  + Cv[3 * i] * dv[0] + Cv[3 * i + 1] * dv[1] + Cv[3 * i + 2] * dv[2];
    df[10] = tq[0] + kr[0] * (dr[1] * r[2] + r[1] * dr[2]);
    df[11] = tq[1] + kr[1] * (dr[2] * r[0] + r[2] * dr[0]);
    df[12] = tq[2] + kr[2] * (dr[0] * r[1] + r[0] * dr[1]);
}

}  // namespace qmpc
