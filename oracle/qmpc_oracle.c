/*
 * oracle/qmpc_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C fp64 restatement of the reference's per-control-step loop
 * (smidmatej/mpc_quad_ros).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product path (mpc_quad_ros_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py replays the reference's own
 * shipped run logs (tests/golden/*.npz, extracted by oracle/make_golden.py from
 * /root/reference/outputs/python_simulation/data/*.pkl) through these functions:
 * nominal RK4 <= 1e-13, RGP mean/cov <= 1e-9 rel, drag residual <= 1e-12,
 * RTI first control <= 1e-5 abs of the acados log (HPIPM's own tolerance),
 * RTI objective <= 2e-5 rel.
 *
 * What each function restates (paths relative to /root/reference):
 *   orc_f / f_eval            src/quad_opt.py:164-262 (setup_casadi_model), with
 *                             src/utils/utils.py:317-340 (v_dot_q, q_to_rot_mat),
 *                             :394-412 (skew_symmetric), :434-440 (quaternion_inverse);
 *                             GP term src/gp/RGP.py:250-254, :52-56
 *   orc_rk4                   src/quad_opt.py:353-377 (discrete_dynamics)
 *   orc_linearize             acados ERK (4 stages, 1 step) forward sensitivities of the
 *                             same map: src/_acados_ocp.json:2126,2138 (third party, restated)
 *   orc_rti_step              src/quad_opt.py:104-151 (cost/bounds/options), :295-350
 *                             (yref, x0 pin, one SQP_RTI iteration, read-back, get_cost);
 *                             acados SQP_RTI + HPIPM are third party and un-vendored: the QP
 *                             is strictly convex so its unique minimiser is computed here by
 *                             a Riccati Mehrotra IPM + exact active-set polish.
 *   orc_rgp_*                 src/gp/RGP.py:106-157 (prior), :168-229 (predict),
 *                             :303-330 (regress); src/gp/GPE.py:244-268
 *   orc_compute_a_drag        src/utils/utils.py:934-950
 *   orc_plant_update          src/quad.py:166-190,234-277,305-381 (Quadrotor3D.update)
 *   orc_reference_chunk       src/utils/utils.py:897-931
 *   orc_closed_loop           src/execute_trajectory.py:196-277 (loop order)
 *
 * Flat parameter conventions (all double):
 *   quad[20] = mass, max_thrust, J[3], x_f[4], y_f[4], z_l_tau[4], g[3]
 *   gp:   M basis points per axis (0 = nominal model), gpX[3*M],
 *         gpth[9] = (L, sigma_f, sigma_n) per axis, alpha[3*M] = K_x^-1 mu per axis
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 13
#define NU 4
#define NZ 17
#define MAXN 128

typedef struct {
    const double *quad;
    int M;
    const double *gpX, *gpth, *alpha;
} model_t;

/* ------------------------------------------------------------------ model */

static void rotmat(const double *q, double R[3][3])
{
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z); R[0][2] = 2 * (x * z + w * y);
    R[1][0] = 2 * (x * y + w * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
    R[2][0] = 2 * (x * z - w * y); R[2][1] = 2 * (y * z + w * x); R[2][2] = 1 - 2 * (x * x + y * y);
}

static void drotmat(const double *q, const double *dq, double dR[3][3])
{
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double dw = dq[0], dx = dq[1], dy = dq[2], dz = dq[3];
    dR[0][0] = -4 * (y * dy + z * dz);
    dR[0][1] = 2 * (dx * y + x * dy - dw * z - w * dz);
    dR[0][2] = 2 * (dx * z + x * dz + dw * y + w * dy);
    dR[1][0] = 2 * (dx * y + x * dy + dw * z + w * dz);
    dR[1][1] = -4 * (x * dx + z * dz);
    dR[1][2] = 2 * (dy * z + y * dz - dw * x - w * dx);
    dR[2][0] = 2 * (dx * z + x * dz - dw * y - w * dy);
    dR[2][1] = 2 * (dy * z + y * dz + dw * x + w * dx);
    dR[2][2] = -4 * (x * dx + y * dy);
}

/* everything about one evaluation point that the JVPs reuse */
typedef struct {
    double x[NX], f[NX];
    double R[3][3], vb[3], mu[3], dmu[3], ab[3];
} evalpt_t;

static void gp_mean(const model_t *m, const double vb[3], double mu[3], double dmu[3])
{
    for (int d = 0; d < 3; ++d) {
        mu[d] = 0; dmu[d] = 0;
        if (m->M <= 0) continue;
        double L = m->gpth[3 * d], sf = m->gpth[3 * d + 1];
        double iL2 = 1.0 / (L * L), sf2 = sf * sf;
        for (int j = 0; j < m->M; ++j) {
            double e = vb[d] - m->gpX[d * m->M + j];
            double k = sf2 * exp(-0.5 * e * iL2 * e);
            double a = m->alpha[d * m->M + j];
            mu[d] += k * a;
            dmu[d] += k * (-e * iL2) * a;
        }
    }
}

static void f_point(const model_t *m, const double *x, const double *u, evalpt_t *e)
{
    const double *Q = m->quad;
    double mass = Q[0], T = Q[1];
    const double *J = Q + 2, *xf = Q + 5, *yf = Q + 9, *zt = Q + 13, *g = Q + 17;
    const double *q = x + 3, *v = x + 7, *r = x + 10;
    memcpy(e->x, x, sizeof(double) * NX);
    rotmat(q, e->R);
    for (int i = 0; i < 3; ++i)
        e->vb[i] = e->R[0][i] * v[0] + e->R[1][i] * v[1] + e->R[2][i] * v[2]; /* R(q)^T v = R(qbar) v */
    gp_mean(m, e->vb, e->mu, e->dmu);
    double aT = T * (u[0] + u[1] + u[2] + u[3]) / mass;
    e->ab[0] = e->mu[0]; e->ab[1] = e->mu[1]; e->ab[2] = aT + e->mu[2];
    double *f = e->f;
    f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
    f[3] = 0.5 * (-r[0] * q[1] - r[1] * q[2] - r[2] * q[3]);
    f[4] = 0.5 * (r[0] * q[0] + r[2] * q[2] - r[1] * q[3]);
    f[5] = 0.5 * (r[1] * q[0] - r[2] * q[1] + r[0] * q[3]);
    f[6] = 0.5 * (r[2] * q[0] + r[1] * q[1] - r[0] * q[2]);
    for (int i = 0; i < 3; ++i)
        f[7 + i] = e->R[i][0] * e->ab[0] + e->R[i][1] * e->ab[1] + e->R[i][2] * e->ab[2] - g[i];
    double ty = 0, tx = 0, tz = 0;
    for (int i = 0; i < 4; ++i) { ty += T * u[i] * yf[i]; tx += T * u[i] * xf[i]; tz += T * u[i] * zt[i]; }
    f[10] = (ty + (J[1] - J[2]) * r[1] * r[2]) / J[0];
    f[11] = (-tx + (J[2] - J[0]) * r[2] * r[0]) / J[1];
    f[12] = (tz + (J[0] - J[1]) * r[0] * r[1]) / J[2];
}

/* directional derivative of f at a cached point along (dx,du) */
static void f_jvp(const model_t *m, const evalpt_t *e, const double *dx, const double *du, double *df)
{
    const double *Q = m->quad;
    double mass = Q[0], T = Q[1];
    const double *J = Q + 2, *xf = Q + 5, *yf = Q + 9, *zt = Q + 13;
    const double *q = e->x + 3, *v = e->x + 7, *r = e->x + 10;
    const double *dq = dx + 3, *dv = dx + 7, *dr = dx + 10;
    double dR[3][3];
    drotmat(q, dq, dR);
    df[0] = dv[0]; df[1] = dv[1]; df[2] = dv[2];
    df[3] = 0.5 * (-dr[0] * q[1] - dr[1] * q[2] - dr[2] * q[3] - r[0] * dq[1] - r[1] * dq[2] - r[2] * dq[3]);
    df[4] = 0.5 * (dr[0] * q[0] + dr[2] * q[2] - dr[1] * q[3] + r[0] * dq[0] + r[2] * dq[2] - r[1] * dq[3]);
    df[5] = 0.5 * (dr[1] * q[0] - dr[2] * q[1] + dr[0] * q[3] + r[1] * dq[0] - r[2] * dq[1] + r[0] * dq[3]);
    df[6] = 0.5 * (dr[2] * q[0] + dr[1] * q[1] - dr[0] * q[2] + r[2] * dq[0] + r[1] * dq[1] - r[0] * dq[2]);
    double dab[3];
    for (int i = 0; i < 3; ++i) {
        double dvb = dR[0][i] * v[0] + dR[1][i] * v[1] + dR[2][i] * v[2]
                   + e->R[0][i] * dv[0] + e->R[1][i] * dv[1] + e->R[2][i] * dv[2];
        dab[i] = e->dmu[i] * dvb;
    }
    dab[2] += T * (du[0] + du[1] + du[2] + du[3]) / mass;
    for (int i = 0; i < 3; ++i)
        df[7 + i] = dR[i][0] * e->ab[0] + dR[i][1] * e->ab[1] + dR[i][2] * e->ab[2]
                  + e->R[i][0] * dab[0] + e->R[i][1] * dab[1] + e->R[i][2] * dab[2];
    double ty = 0, tx = 0, tz = 0;
    for (int i = 0; i < 4; ++i) { ty += T * du[i] * yf[i]; tx += T * du[i] * xf[i]; tz += T * du[i] * zt[i]; }
    df[10] = (ty + (J[1] - J[2]) * (dr[1] * r[2] + r[1] * dr[2])) / J[0];
    df[11] = (-tx + (J[2] - J[0]) * (dr[2] * r[0] + r[2] * dr[0])) / J[1];
    df[12] = (tz + (J[0] - J[1]) * (dr[0] * r[1] + r[0] * dr[1])) / J[2];
}

static model_t mk_model(const double *quad, int M, const double *gpX, const double *gpth, const double *alpha)
{
    model_t m; m.quad = quad; m.M = (alpha && gpX) ? M : 0; m.gpX = gpX; m.gpth = gpth; m.alpha = alpha;
    return m;
}

void orc_f(const double *quad, int M, const double *gpX, const double *gpth, const double *alpha,
           const double *x, const double *u, double *f)
{
    model_t m = mk_model(quad, M, gpX, gpth, alpha);
    evalpt_t e; f_point(&m, x, u, &e);
    memcpy(f, e.f, sizeof(double) * NX);
}

/* classic RK4, one step (quad_opt.py:363-367) */
static void rk4_points(const model_t *m, const double *x, const double *u, double dt, evalpt_t e[4], double *xn)
{
    double xs[NX];
    f_point(m, x, u, &e[0]);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt / 2 * e[0].f[i];
    f_point(m, xs, u, &e[1]);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt / 2 * e[1].f[i];
    f_point(m, xs, u, &e[2]);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt * e[2].f[i];
    f_point(m, xs, u, &e[3]);
    for (int i = 0; i < NX; ++i)
        xn[i] = x[i] + dt / 6 * (e[0].f[i] + 2 * e[1].f[i] + 2 * e[2].f[i] + e[3].f[i]);
}

void orc_rk4(const double *quad, int M, const double *gpX, const double *gpth, const double *alpha,
             const double *x, const double *u, double dt, double *xn)
{
    model_t m = mk_model(quad, M, gpX, gpth, alpha);
    evalpt_t e[4];
    rk4_points(&m, x, u, dt, e, xn);
}

/* Phi, A = dPhi/dx [13x13 row-major], B = dPhi/du [13x4 row-major] */
static void linearize(const model_t *m, const double *x, const double *u, double dt,
                      double *Phi, double *A, double *B)
{
    evalpt_t e[4];
    rk4_points(m, x, u, dt, e, Phi);
    for (int c = 0; c < NZ; ++c) {
        double ex[NX] = {0}, eu[NU] = {0}, d1[NX], d2[NX], d3[NX], d4[NX], s[NX];
        if (c < NX) ex[c] = 1; else eu[c - NX] = 1;
        f_jvp(m, &e[0], ex, eu, d1);
        for (int i = 0; i < NX; ++i) s[i] = ex[i] + dt / 2 * d1[i];
        f_jvp(m, &e[1], s, eu, d2);
        for (int i = 0; i < NX; ++i) s[i] = ex[i] + dt / 2 * d2[i];
        f_jvp(m, &e[2], s, eu, d3);
        for (int i = 0; i < NX; ++i) s[i] = ex[i] + dt * d3[i];
        f_jvp(m, &e[3], s, eu, d4);
        for (int i = 0; i < NX; ++i) {
            double col = ex[i] + dt / 6 * (d1[i] + 2 * d2[i] + 2 * d3[i] + d4[i]);
            if (c < NX) A[i * NX + c] = col; else B[i * NU + (c - NX)] = col;
        }
    }
}

void orc_linearize(const double *quad, int M, const double *gpX, const double *gpth, const double *alpha,
                   const double *x, const double *u, double dt, double *Phi, double *A, double *B)
{
    model_t m = mk_model(quad, M, gpX, gpth, alpha);
    linearize(&m, x, u, dt, Phi, A, B);
}

/* ------------------------------------------------------- box-QP (Riccati) */

typedef struct {
    int N;
    const double *A, *B, *c;     /* [N][169], [N][52], [N][13] */
    double Qd[NX], QNd[NX], Rd[NU];
    const double *q;             /* [(N+1)][13] linear state cost */
    const double *r;             /* [N][4] linear input cost */
    const double *x0;
    double lb, ub;
} qp_t;

typedef struct { double Lam[NU][NU]; double Lx[NU][NX]; } fac_t;

/* Solve the LQR  min sum 1/2 u'(R+diag(dR))u + rt'u + 1/2 x'Qx + q'x  s.t. dynamics.
   Bmask (optional, [N][4]): 0 => that input is removed (column zeroed); its value is taken from ufix. */
static int riccati(const qp_t *qp, const double *dR, const double *rt,
                   const unsigned char *fixed, const double *ufix, double *xs, double *us)
{
    int N = qp->N;
    double P[NX][NX], p[NX];
    fac_t *fac = (fac_t *)malloc(sizeof(fac_t) * N);
    double(*lv)[NU] = (double(*)[NU])malloc(sizeof(double) * NU * N);
    double *ceff = (double *)malloc(sizeof(double) * NX * N);
    double *Beff = (double *)malloc(sizeof(double) * NX * NU * N);
    int ok = 1;
    memset(P, 0, sizeof(P));
    for (int i = 0; i < NX; ++i) { P[i][i] = qp->QNd[i]; p[i] = qp->q[N * NX + i]; }
    for (int k = N - 1; k >= 0; --k) {
        const double *A = qp->A + k * NX * NX;
        double *B = Beff + k * NX * NU, *c = ceff + k * NX;
        memcpy(B, qp->B + k * NX * NU, sizeof(double) * NX * NU);
        memcpy(c, qp->c + k * NX, sizeof(double) * NX);
        double Rk[NU], rk[NU];
        for (int j = 0; j < NU; ++j) {
            Rk[j] = qp->Rd[j] + (dR ? dR[k * NU + j] : 0.0);
            rk[j] = rt[k * NU + j];
            if (fixed && fixed[k * NU + j]) {
                for (int i = 0; i < NX; ++i) { c[i] += B[i * NU + j] * ufix[k * NU + j]; B[i * NU + j] = 0; }
                Rk[j] = 1.0; rk[j] = 0.0;
            }
        }
        double PA[NX][NX], PB[NX][NU], h[NX];
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) { double s = 0; for (int l = 0; l < NX; ++l) s += P[i][l] * A[l * NX + j]; PA[i][j] = s; }
            for (int j = 0; j < NU; ++j) { double s = 0; for (int l = 0; l < NX; ++l) s += P[i][l] * B[l * NU + j]; PB[i][j] = s; }
            double s = p[i]; for (int l = 0; l < NX; ++l) s += P[i][l] * c[l]; h[i] = s;
        }
        double Muu[NU][NU], Mux[NU][NX], Mxx[NX][NX], gu[NU], gx[NX];
        for (int a = 0; a < NU; ++a) {
            for (int b = 0; b < NU; ++b) { double s = 0; for (int l = 0; l < NX; ++l) s += B[l * NU + a] * PB[l][b]; Muu[a][b] = s; }
            Muu[a][a] += Rk[a];
            for (int j = 0; j < NX; ++j) { double s = 0; for (int l = 0; l < NX; ++l) s += B[l * NU + a] * PA[l][j]; Mux[a][j] = s; }
            double s = rk[a]; for (int l = 0; l < NX; ++l) s += B[l * NU + a] * h[l]; gu[a] = s;
        }
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) { double s = 0; for (int l = 0; l < NX; ++l) s += A[l * NX + i] * PA[l][j]; Mxx[i][j] = s; }
            Mxx[i][i] += qp->Qd[i];
            double s = qp->q[k * NX + i]; for (int l = 0; l < NX; ++l) s += A[l * NX + i] * h[l]; gx[i] = s;
        }
        /* Cholesky of Muu */
        fac_t *F = &fac[k];
        memset(F->Lam, 0, sizeof(F->Lam));
        for (int j = 0; j < NU; ++j) {
            double d = Muu[j][j];
            for (int l = 0; l < j; ++l) d -= F->Lam[j][l] * F->Lam[j][l];
            if (!(d > 0)) { ok = 0; d = 1e-300; }
            d = sqrt(d); F->Lam[j][j] = d;
            for (int i = j + 1; i < NU; ++i) {
                double s = Muu[i][j];
                for (int l = 0; l < j; ++l) s -= F->Lam[i][l] * F->Lam[j][l];
                F->Lam[i][j] = s / d;
            }
        }
        for (int j = 0; j < NX; ++j)
            for (int a = 0; a < NU; ++a) {
                double s = Mux[a][j];
                for (int l = 0; l < a; ++l) s -= F->Lam[a][l] * F->Lx[l][j];
                F->Lx[a][j] = s / F->Lam[a][a];
            }
        for (int a = 0; a < NU; ++a) {
            double s = gu[a];
            for (int l = 0; l < a; ++l) s -= F->Lam[a][l] * lv[k][l];
            lv[k][a] = s / F->Lam[a][a];
        }
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) {
                double s = Mxx[i][j];
                for (int a = 0; a < NU; ++a) s -= F->Lx[a][i] * F->Lx[a][j];
                P[i][j] = s;
            }
            double s = gx[i];
            for (int a = 0; a < NU; ++a) s -= F->Lx[a][i] * lv[k][a];
            p[i] = s;
        }
        for (int i = 0; i < NX; ++i)
            for (int j = i + 1; j < NX; ++j) { double s = 0.5 * (P[i][j] + P[j][i]); P[i][j] = s; P[j][i] = s; }
    }
    memcpy(xs, qp->x0, sizeof(double) * NX);
    for (int k = 0; k < N; ++k) {
        const double *A = qp->A + k * NX * NX;
        const double *B = Beff + k * NX * NU, *c = ceff + k * NX;
        const fac_t *F = &fac[k];
        const double *x = xs + k * NX;
        double *u = us + k * NU, *xn = xs + (k + 1) * NX;
        double w[NU];
        for (int a = 0; a < NU; ++a) { double s = lv[k][a]; for (int j = 0; j < NX; ++j) s += F->Lx[a][j] * x[j]; w[a] = -s; }
        for (int a = NU - 1; a >= 0; --a) {
            double s = w[a];
            for (int l = a + 1; l < NU; ++l) s -= F->Lam[l][a] * u[l];
            u[a] = s / F->Lam[a][a];
        }
        if (fixed) for (int a = 0; a < NU; ++a) if (fixed[k * NU + a]) u[a] = 0.0; /* decoupled dummy */
        for (int i = 0; i < NX; ++i) {
            double s = c[i];
            for (int j = 0; j < NX; ++j) s += A[i * NX + j] * x[j];
            for (int a = 0; a < NU; ++a) s += B[i * NU + a] * u[a];
            xn[i] = s;
        }
        if (fixed) for (int a = 0; a < NU; ++a) if (fixed[k * NU + a]) u[a] = ufix[k * NU + a];
    }
    free(fac); free(lv); free(ceff); free(Beff);
    return ok;
}

/* reduced gradient wrt u through the adjoint; xs is re-simulated from us so the pair is dynamics-exact */
static void qp_gradient(const qp_t *qp, const double *us, double *xs, double *gu)
{
    int N = qp->N;
    memcpy(xs, qp->x0, sizeof(double) * NX);
    for (int k = 0; k < N; ++k) {
        const double *A = qp->A + k * NX * NX, *B = qp->B + k * NX * NU, *c = qp->c + k * NX;
        for (int i = 0; i < NX; ++i) {
            double s = c[i];
            for (int j = 0; j < NX; ++j) s += A[i * NX + j] * xs[k * NX + j];
            for (int a = 0; a < NU; ++a) s += B[i * NU + a] * us[k * NU + a];
            xs[(k + 1) * NX + i] = s;
        }
    }
    double pi[NX], pn[NX];
    for (int i = 0; i < NX; ++i) pi[i] = qp->QNd[i] * xs[N * NX + i] + qp->q[N * NX + i];
    for (int k = N - 1; k >= 0; --k) {
        const double *A = qp->A + k * NX * NX, *B = qp->B + k * NX * NU;
        for (int a = 0; a < NU; ++a) {
            double s = qp->Rd[a] * us[k * NU + a] + qp->r[k * NU + a];
            for (int l = 0; l < NX; ++l) s += B[l * NU + a] * pi[l];
            gu[k * NU + a] = s;
        }
        for (int i = 0; i < NX; ++i) {
            double s = qp->Qd[i] * xs[k * NX + i] + qp->q[k * NX + i];
            for (int l = 0; l < NX; ++l) s += A[l * NX + i] * pi[l];
            pn[i] = s;
        }
        memcpy(pi, pn, sizeof(pi));
    }
}

int orc_debug = 0;         /* >0: print the IPM iteration log (diagnostic) */
int orc_last_rounds = 0;   /* refinement rounds of the most recent boxqp_solve (diagnostic, not thread safe) */

/* Mehrotra predictor-corrector on the box-QP, then exact active-set refinement.
   The slacks t_l = u - lb, t_u = ub - u are carried as independent positive variables (updated with the step, never
   recomputed from u) so that complementarity can be driven far below the spacing of doubles around the bounds.
   returns status: 0 exact (active set verified), 1 IPM converged but refinement rejected, 2 max iterations, 3 NaN */
static int boxqp_solve(const qp_t *qp, double *xs, double *us, int *iters_out, double *kkt_out,
                       double mu_tol, int max_iter, int do_polish)
{
    int N = qp->N, n = N * NU;
    double *buf = (double *)malloc(sizeof(double) * (11 * n + (N + 1) * NX));
    double *u = buf, *ll = buf + n, *lu = buf + 2 * n, *dR = buf + 3 * n, *rt = buf + 4 * n, *ua = buf + 5 * n;
    double *uc = buf + 6 * n, *dla = buf + 7 * n, *dua = buf + 8 * n, *tl = buf + 9 * n, *tu = buf + 10 * n;
    double *xw = buf + 11 * n;
    double lb = qp->lb, ub = qp->ub;
    /* initial multipliers scaled with the problem: clip(0.02 * mean |dJ/du| at the box centre, 0.1, 1e5) */
    double lam0 = 0.1;
    {
        double *g0 = (double *)malloc(sizeof(double) * n);
        for (int i = 0; i < n; ++i) u[i] = 0.5 * (lb + ub);
        qp_gradient(qp, u, xw, g0);
        double gs = 0;
        for (int i = 0; i < n; ++i) gs += fabs(g0[i]);
        gs /= n;
        if (gs == gs) lam0 = fmin(fmax(0.02 * gs, 0.1), 1e5);
        free(g0);
    }
    for (int i = 0; i < n; ++i) { u[i] = 0.5 * (lb + ub); tl[i] = u[i] - lb; tu[i] = ub - u[i]; ll[i] = lam0; lu[i] = lam0; }
    int it = 0, status = 2;
    double mu = 0, resfac = 1.0;   /* resfac: fraction of the initial stationarity residual still present */
    for (it = 0; it < max_iter; ++it) {
        mu = 0;
        for (int i = 0; i < n; ++i) mu += ll[i] * tl[i] + lu[i] * tu[i];
        mu /= (2.0 * n);
        if (!(mu == mu) || mu > 1e300) { status = 3; break; }
        if (mu < mu_tol && resfac < 1e-3) { status = 1; break; }
        /* predictor */
        for (int i = 0; i < n; ++i) {
            dR[i] = ll[i] / tl[i] + lu[i] / tu[i];
            rt[i] = qp->r[i] - dR[i] * u[i];
        }
        riccati(qp, dR, rt, NULL, NULL, xw, ua);
        double ap = 1.0, ad = 1.0;     /* primal (slack) and dual (multiplier) step lengths are separate */
        for (int i = 0; i < n; ++i) {
            double du = ua[i] - u[i];
            dla[i] = -ll[i] - ll[i] / tl[i] * du;
            dua[i] = -lu[i] + lu[i] / tu[i] * du;
            if (du < 0) ap = fmin(ap, -tl[i] / du);
            if (du > 0) ap = fmin(ap, tu[i] / du);
            if (dla[i] < 0) ad = fmin(ad, -ll[i] / dla[i]);
            if (dua[i] < 0) ad = fmin(ad, -lu[i] / dua[i]);
        }
        double mua = 0;
        for (int i = 0; i < n; ++i) {
            double du = ua[i] - u[i];
            mua += (ll[i] + ad * dla[i]) * (tl[i] + ap * du) + (lu[i] + ad * dua[i]) * (tu[i] - ap * du);
        }
        mua /= (2.0 * n);
        double sigma = mua / mu; sigma = sigma * sigma * sigma;
        /* corrector; safeguard: a blocked Mehrotra step (< 1/2) is recomputed once as a centring step without the
           second-order term (Mehrotra's heuristic can otherwise cycle on badly centred iterates) */
        double so = 1.0;
        double *cl = uc + 0;   /* uc is re-used below, keep the second-order terms in their own buffers */
        (void)cl;
        double *cl2 = (double *)malloc(sizeof(double) * 2 * n), *cu2 = cl2 + n;
        for (int i = 0; i < n; ++i) { double du = ua[i] - u[i]; cl2[i] = du * dla[i]; cu2[i] = -du * dua[i]; }
        for (int pass = 0; pass < 2; ++pass) {
            double smu = sigma * mu;
            for (int i = 0; i < n; ++i)
                rt[i] = qp->r[i] - dR[i] * u[i] - (smu - so * cl2[i]) / tl[i] + (smu - so * cu2[i]) / tu[i];
            riccati(qp, dR, rt, NULL, NULL, xw, uc);
            ap = 1e300; ad = 1e300;
            for (int i = 0; i < n; ++i) {
                double du = uc[i] - u[i];
                double dl = (smu - so * cl2[i]) / tl[i] - ll[i] - ll[i] / tl[i] * du;
                double dv = (smu - so * cu2[i]) / tu[i] - lu[i] + lu[i] / tu[i] * du;
                dla[i] = dl; dua[i] = dv;
                if (du < 0) ap = fmin(ap, -tl[i] / du);
                if (du > 0) ap = fmin(ap, tu[i] / du);
                if (dl < 0) ad = fmin(ad, -ll[i] / dl);
                if (dv < 0) ad = fmin(ad, -lu[i] / dv);
            }
            if (pass == 1 || fmin(ap, ad) >= 0.5) break;
            so = 0.0; sigma = fmax(sigma, 0.5);
        }
        free(cl2);
        ap = fmin(1.0, 0.995 * ap); ad = fmin(1.0, 0.995 * ad);
        if (orc_debug) printf("it %2d mu %.3e mu_aff %.3e sigma %.3e alpha %.4f %.4f\n", it, mu, mua, sigma, ap, ad);
        for (int i = 0; i < n; ++i) {
            double du = uc[i] - u[i];
            u[i] += ap * du; tl[i] += ap * du; tu[i] -= ap * du;
            ll[i] += ad * dla[i];
            lu[i] += ad * dua[i];
        }
        resfac *= 1.0 - fmin(ap, ad);
    }
    double *gu = (double *)malloc(sizeof(double) * n);
    /* IPM answer (clipped into the box: the slacks, not u, carry the sub-ulp distance to a bound) and its KKT residual */
    for (int i = 0; i < n; ++i) u[i] = fmin(fmax(u[i], lb), ub);
    qp_gradient(qp, u, xw, gu);
    double kkt = 0;
    for (int i = 0; i < n; ++i) kkt = fmax(kkt, fabs(gu[i] - ll[i] + lu[i]));
    kkt = fmax(kkt, mu);
    memcpy(us, u, sizeof(double) * n);
    memcpy(xs, xw, sizeof(double) * (N + 1) * NX);
    if (do_polish && status == 1) {
        /* primal-dual active-set refinement started from the IPM's guess: solve the equality-constrained LQR with the
           active inputs pinned; move violated free inputs onto their bound, release pinned inputs whose multiplier has
           the wrong sign; stop when nothing changes => exact KKT point of the strictly convex QP. */
        unsigned char *fixed = (unsigned char *)calloc(n, 1);
        double *ufix = (double *)calloc(n, sizeof(double));
        for (int i = 0; i < n; ++i) {
            if (tl[i] < ll[i]) { fixed[i] = 1; ufix[i] = lb; }
            else if (tu[i] < lu[i]) { fixed[i] = 2; ufix[i] = ub; }
        }
        for (int round = 0; round < 20; ++round) {
            orc_last_rounds = round + 1;
            riccati(qp, NULL, qp->r, fixed, ufix, xw, ua);
            qp_gradient(qp, ua, xw, gu);
            int changed = 0;
            double viol = 0;
            for (int i = 0; i < n; ++i) {
                if (fixed[i] == 1) { if (gu[i] < -1e-12) { fixed[i] = 0; changed = 1; } }
                else if (fixed[i] == 2) { if (gu[i] > 1e-12) { fixed[i] = 0; changed = 1; } }
                else {
                    viol = fmax(viol, fabs(gu[i]));
                    if (ua[i] < lb) { fixed[i] = 1; ufix[i] = lb; changed = 1; }
                    else if (ua[i] > ub) { fixed[i] = 2; ufix[i] = ub; changed = 1; }
                }
            }
            if (!changed) {
                if (viol < 1e-9) {
                    status = 0; kkt = viol;
                    memcpy(us, ua, sizeof(double) * n);
                    memcpy(xs, xw, sizeof(double) * (N + 1) * NX);
                }
                break;
            }
        }
        free(fixed); free(ufix);
    }
    *iters_out = it; *kkt_out = kkt;
    free(buf); free(gu);
    return status;
}

/* ------------------------------------------------------------ RTI step */

/* One SQP-RTI iteration.  xit[(N+1)*13], uit[N*4] hold the persistent iterate (in/out).
   Wd[17] stage weights (multiplied by dt inside), Wed[13] terminal weights.
   qp_out (optional): Phi/A/B dump for tests: [N][13+169+52]. */
int orc_rti_step(const double *quad, double dt, int N,
                 int M, const double *gpX, const double *gpth, const double *alpha,
                 const double *Wd, const double *Wed, double lbu, double ubu,
                 const double *x0, const double *yref, const double *yref_e,
                 double *xit, double *uit, double *cost, int *iters, double *kkt,
                 double mu_tol, int max_iter, int do_polish, double *lin_out)
{
    model_t m = mk_model(quad, M, gpX, gpth, alpha);
    double *A = (double *)malloc(sizeof(double) * N * NX * NX);
    double *B = (double *)malloc(sizeof(double) * N * NX * NU);
    double *c = (double *)malloc(sizeof(double) * N * NX);
    double *q = (double *)malloc(sizeof(double) * (N + 1) * NX);
    double *r = (double *)malloc(sizeof(double) * N * NU);
    for (int k = 0; k < N; ++k) {
        double Phi[NX];
        const double *x = xit + k * NX, *u = uit + k * NU;
        double *Ak = A + k * NX * NX, *Bk = B + k * NX * NU;
        linearize(&m, x, u, dt, Phi, Ak, Bk);
        for (int i = 0; i < NX; ++i) {
            double s = Phi[i];
            for (int j = 0; j < NX; ++j) s -= Ak[i * NX + j] * x[j];
            for (int a = 0; a < NU; ++a) s -= Bk[i * NU + a] * u[a];
            c[k * NX + i] = s;
        }
        if (lin_out) {
            double *o = lin_out + k * (NX + NX * NX + NX * NU);
            memcpy(o, Phi, sizeof(Phi)); memcpy(o + NX, Ak, sizeof(double) * NX * NX);
            memcpy(o + NX + NX * NX, Bk, sizeof(double) * NX * NU);
        }
        for (int i = 0; i < NX; ++i) q[k * NX + i] = -dt * Wd[i] * yref[k * NZ + i];
        for (int a = 0; a < NU; ++a) r[k * NU + a] = -dt * Wd[NX + a] * yref[k * NZ + NX + a];
    }
    for (int i = 0; i < NX; ++i) q[N * NX + i] = -Wed[i] * yref_e[i];
    qp_t qp; qp.N = N; qp.A = A; qp.B = B; qp.c = c; qp.q = q; qp.r = r; qp.x0 = x0; qp.lb = lbu; qp.ub = ubu;
    for (int i = 0; i < NX; ++i) { qp.Qd[i] = dt * Wd[i]; qp.QNd[i] = Wed[i]; }
    for (int a = 0; a < NU; ++a) qp.Rd[a] = dt * Wd[NX + a];
    int status = boxqp_solve(&qp, xit, uit, iters, kkt, mu_tol, max_iter, do_polish);
    double J = 0;
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < NX; ++i) { double e = xit[k * NX + i] - yref[k * NZ + i]; J += 0.5 * dt * Wd[i] * e * e; }
        for (int a = 0; a < NU; ++a) { double e = uit[k * NU + a] - yref[k * NZ + NX + a]; J += 0.5 * dt * Wd[NX + a] * e * e; }
    }
    for (int i = 0; i < NX; ++i) { double e = xit[N * NX + i] - yref_e[i]; J += 0.5 * Wed[i] * e * e; }
    *cost = J;
    free(A); free(B); free(c); free(q); free(r);
    return status;
}

/* ---------------------------------------------------------------- RGP */

static double rbf(double a, double b, double L, double sf)
{
    double e = a - b;
    return sf * sf * exp(-0.5 * e * (1.0 / (L * L)) * e);
}

/* K_x = K(X,X)+sn^2 I and its inverse (Gauss-Jordan, partial pivoting). returns 0 ok */
int orc_rgp_prior(int M, const double *X, const double *th, double *Kx, double *Kx_inv)
{
    double *W = (double *)malloc(sizeof(double) * M * 2 * M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            double k = rbf(X[i], X[j], th[0], th[1]) + (i == j ? th[2] * th[2] : 0.0);
            Kx[i * M + j] = k; W[i * 2 * M + j] = k; W[i * 2 * M + M + j] = (i == j);
        }
    for (int c = 0; c < M; ++c) {
        int piv = c;
        for (int i = c + 1; i < M; ++i) if (fabs(W[i * 2 * M + c]) > fabs(W[piv * 2 * M + c])) piv = i;
        if (W[piv * 2 * M + c] == 0) { free(W); return 1; }
        if (piv != c) for (int j = 0; j < 2 * M; ++j) { double t = W[c * 2 * M + j]; W[c * 2 * M + j] = W[piv * 2 * M + j]; W[piv * 2 * M + j] = t; }
        double d = 1.0 / W[c * 2 * M + c];
        for (int j = 0; j < 2 * M; ++j) W[c * 2 * M + j] *= d;
        for (int i = 0; i < M; ++i) if (i != c) {
            double f = W[i * 2 * M + c];
            if (f != 0) for (int j = 0; j < 2 * M; ++j) W[i * 2 * M + j] -= f * W[c * 2 * M + j];
        }
    }
    for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) Kx_inv[i * M + j] = W[i * 2 * M + M + j];
    free(W);
    return 0;
}

/* RGP.regress with one sample (RGP.py:303-330, predict :199-208).  mu[M], C[M*M] updated in place. */
void orc_rgp_regress(int M, const double *X, const double *th, const double *Kx_inv,
                     double *mu, double *C, double xt, double yt)
{
    double *kv = (double *)malloc(sizeof(double) * M * 4);
    double *Jt = kv + M, *w = kv + 2 * M, *cj = kv + 3 * M;
    for (int i = 0; i < M; ++i) kv[i] = rbf(xt, X[i], th[0], th[1]);
    for (int j = 0; j < M; ++j) { double s = 0; for (int i = 0; i < M; ++i) s += kv[i] * Kx_inv[i * M + j]; Jt[j] = s; }
    double mp = 0, jk = 0;
    for (int j = 0; j < M; ++j) { mp += Jt[j] * mu[j]; jk += Jt[j] * rbf(X[j], xt, th[0], th[1]); }
    double b = rbf(xt, xt, th[0], th[1]) - jk;
    for (int j = 0; j < M; ++j) { double s = 0; for (int i = 0; i < M; ++i) s += Jt[i] * C[i * M + j]; w[j] = s; }   /* Jt C   */
    for (int i = 0; i < M; ++i) { double s = 0; for (int j = 0; j < M; ++j) s += C[i * M + j] * Jt[j]; cj[i] = s; } /* C Jt^T */
    double jcj = 0; for (int j = 0; j < M; ++j) jcj += w[j] * Jt[j];
    double cp = b + jcj;
    double sinv = 1.0 / (cp + th[2] * th[2]);
    double innov = yt - mp;
    for (int i = 0; i < M; ++i) {
        double g = cj[i] * sinv;
        mu[i] += g * innov;
        for (int j = 0; j < M; ++j) C[i * M + j] -= g * w[j];
    }
    free(kv);
}

/* ---- RGP* hyper-parameter learning: RGP.learn (reference src/gp/RGP.py:332-482) with its unscented transform
 * (__draw_sigma_points, :485-505), restated as written, including the parts that look unintended:
 *   - At (hence Jt) is built once at the current hyper-parameters and reused for every sigma point (:357-366, :392);
 *   - the running mean mu_p_t of the cumulative sum is used inside the outer product of the same iteration (:403-404);
 *   - C_g_eta_t is never assigned by learn, it keeps the value it came in with (zero after __init__, :153);
 *   - the exp() transform of eta (:468-470) is overwritten by the plain assignment (:472-474).
 * State of one 1-D model: X[M], mu_g[M], C_g[M*M], mu_eta[3]=(L,sigma_f,sigma_n), C_eta[9], C_g_eta[M*3], Kx_inv[M*M];
 * all updated in place for the sample (xt, yt).  mu_z[M+3] / C_z[(M+3)^2] (the return value of learn) are optional. */
static void sym3_sqrt(const double *A, double *S)      /* principal square root of a symmetric positive definite 3x3 */
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (A[i * 3 + j] + A[j * 3 + i]);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += v[i][k] * sqrt(a[k][k]) * v[j][k];
            S[i * 3 + j] = acc;
        }
}

static int gj_inverse(int n, const double *A, double *Ainv)   /* Gauss-Jordan with partial pivoting */
{
    double *W = (double *)malloc(sizeof(double) * n * 2 * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { W[i * 2 * n + j] = A[i * n + j]; W[i * 2 * n + n + j] = (i == j); }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int i = c + 1; i < n; ++i) if (fabs(W[i * 2 * n + c]) > fabs(W[piv * 2 * n + c])) piv = i;
        if (W[piv * 2 * n + c] == 0) { free(W); return 1; }
        if (piv != c) for (int j = 0; j < 2 * n; ++j) { double t = W[c * 2 * n + j]; W[c * 2 * n + j] = W[piv * 2 * n + j]; W[piv * 2 * n + j] = t; }
        double d = 1.0 / W[c * 2 * n + c];
        for (int j = 0; j < 2 * n; ++j) W[c * 2 * n + j] *= d;
        for (int i = 0; i < n; ++i) if (i != c) {
            double f = W[i * 2 * n + c];
            if (f != 0) for (int j = 0; j < 2 * n; ++j) W[i * 2 * n + j] -= f * W[c * 2 * n + j];
        }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Ainv[i * n + j] = W[i * 2 * n + n + j];
    free(W);
    return 0;
}

int orc_rgp_learn(int M, const double *X, double *mu_g, double *C_g, double *mu_eta, double *C_eta,
                  const double *C_g_eta, double *Kx_inv, double xt, double yt, double *mu_z_out, double *C_z_out)
{
    const int ne = 3, np_ = M + 4, nu_ = M + 2, nz = M + 3;
    const double L = mu_eta[0], sf = mu_eta[1];
    double *buf = (double *)calloc((size_t)(4 * M + M * 3 + M * M + 2 * np_ + 2 * np_ * np_ + nz * nz + 2 * nu_), sizeof(double));
    double *kv = buf, *Jt = kv + M, *vg = Jt + M, *JC = vg + M, *St = JC + M, *Cgp = St + M * 3;
    double *mu_p = Cgp + M * M, *mu_pi = mu_p + np_, *C_pi = mu_pi + np_, *C_p = C_pi + np_ * np_;
    double *C_z = C_p + np_ * np_, *Lt = C_z + nz * nz;
    /* inference step (:357-366) */
    for (int i = 0; i < M; ++i) kv[i] = rbf(xt, X[i], L, sf);
    for (int j = 0; j < M; ++j) { double a = 0; for (int i = 0; i < M; ++i) a += kv[i] * Kx_inv[i * M + j]; Jt[j] = a; }
    double jk = 0;
    for (int j = 0; j < M; ++j) jk += Jt[j] * rbf(X[j], xt, L, sf);
    const double Bv = rbf(xt, xt, L, sf) - jk;
    double Cei[9];
    if (gj_inverse(3, C_eta, Cei)) { free(buf); return 1; }
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < ne; ++j) { double a = 0; for (int k = 0; k < ne; ++k) a += C_g_eta[i * ne + k] * Cei[k * ne + j]; St[i * ne + j] = a; }
    /* C_p_i = At [C_g - St C_g_eta^T, 0; 0, 0] At^T + C_w is the same for every sigma point (:398-399) */
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) { double a = 0; for (int k = 0; k < ne; ++k) a += St[i * ne + k] * C_g_eta[j * ne + k]; Cgp[i * M + j] = C_g[i * M + j] - a; }
    for (int j = 0; j < M; ++j) { double a = 0; for (int i = 0; i < M; ++i) a += Jt[i] * Cgp[i * M + j]; JC[j] = a; }       /* Jt Cg' */
    double *CJ = vg;                                                                                                 /* Cg' Jt^T (vg reused later) */
    double *cj = (double *)malloc(sizeof(double) * M);
    for (int i = 0; i < M; ++i) { double a = 0; for (int j = 0; j < M; ++j) a += Cgp[i * M + j] * Jt[j]; cj[i] = a; }
    double jcj = 0; for (int j = 0; j < M; ++j) jcj += JC[j] * Jt[j];
    for (int i = 0; i < M; ++i) {
        for (int j = 0; j < M; ++j) C_pi[i * np_ + j] = Cgp[i * M + j];
        C_pi[i * np_ + (M + 3)] = cj[i];
        C_pi[(M + 3) * np_ + i] = JC[i];
    }
    C_pi[(M + 3) * np_ + (M + 3)] = jcj + Bv;
    (void)CJ;
    /* unscented transform (:379-404, :485-505) */
    double S6[9], Ssq[9];
    for (int i = 0; i < 9; ++i) S6[i] = (3.0 / (1.0 - 0.5)) * C_eta[i];
    sym3_sqrt(S6, Ssq);
    for (int sp = 0; sp < 7; ++sp) {
        double eta[3], w = sp == 0 ? 0.5 : (1.0 - 0.5) / 6.0;
        for (int k = 0; k < 3; ++k)
            eta[k] = sp == 0 ? mu_eta[k] : (sp <= 3 ? mu_eta[k] + Ssq[k * 3 + (sp - 1)] : mu_eta[k] - Ssq[k * 3 + (sp - 4)]);
        double jg = 0;
        for (int i = 0; i < M; ++i) {
            double a = 0;
            for (int k = 0; k < ne; ++k) a += St[i * ne + k] * (eta[k] - mu_eta[k]);
            vg[i] = mu_g[i] + a;
            jg += Jt[i] * vg[i];
        }
        for (int i = 0; i < M; ++i) mu_pi[i] = vg[i];
        for (int k = 0; k < 3; ++k) mu_pi[M + k] = eta[k];
        mu_pi[M + 3] = jg;
        for (int i = 0; i < np_; ++i) mu_p[i] += w * mu_pi[i];
        for (int i = 0; i < np_; ++i)
            for (int j = 0; j < np_; ++j)
                C_p[i * np_ + j] += w * ((mu_pi[i] - mu_p[i]) * (mu_pi[j] - mu_p[j]) + C_pi[i * np_ + j]);
    }
    /* update step (:409-455): o = [sigma_n, g_t] at rows M+2, M+3; u = [g, L, sigma_f] */
    const int o0 = M + 2, o1 = M + 3;
    const double mo0 = mu_p[o0], mo1 = mu_p[o1];
    const double Co00 = C_p[o0 * np_ + o0], Co01 = C_p[o0 * np_ + o1], Co10 = C_p[o1 * np_ + o0], Co11 = C_p[o1 * np_ + o1];
    const double Cy = Co11 + Co00 + mo0 * mo0;
    const double G0 = Co01 / Cy, G1 = Co11 / Cy;
    const double me0 = mo0 + G0 * (yt - mo1), me1 = mo1 + G1 * (yt - mo1);
    const double Ce00 = Co00 - G0 * Cy * G0, Ce01 = Co01 - G0 * Cy * G1, Ce10 = Co10 - G1 * Cy * G0, Ce11 = Co11 - G1 * Cy * G1;
    const double det = Co00 * Co11 - Co01 * Co10;
    const double Ci00 = Co11 / det, Ci01 = -Co01 / det, Ci10 = -Co10 / det, Ci11 = Co00 / det;
    for (int i = 0; i < nu_; ++i) {                      /* Lt = C_ou^T inv(C_o) */
        const double c0 = C_p[o0 * np_ + i], c1 = C_p[o1 * np_ + i];
        Lt[i * 2] = c0 * Ci00 + c1 * Ci10;
        Lt[i * 2 + 1] = c0 * Ci01 + c1 * Ci11;
    }
    const double d0 = me0 - mo0, d1 = me1 - mo1;
    const double D00 = Ce00 - Co00, D01 = Ce01 - Co01, D10 = Ce10 - Co10, D11 = Ce11 - Co11;
    double *mu_z = mu_pi;                                 /* reuse */
    for (int i = 0; i < nu_; ++i) mu_z[i] = mu_p[i] + Lt[i * 2] * d0 + Lt[i * 2 + 1] * d1;
    mu_z[nu_] = me0;
    for (int i = 0; i < nu_; ++i) {
        const double t0 = Lt[i * 2] * D00 + Lt[i * 2 + 1] * D10, t1 = Lt[i * 2] * D01 + Lt[i * 2 + 1] * D11;
        for (int j = 0; j < nu_; ++j) C_z[i * nz + j] = C_p[i * np_ + j] + t0 * Lt[j * 2] + t1 * Lt[j * 2 + 1];
        const double lc = Lt[i * 2] * Ce00 + Lt[i * 2 + 1] * Ce10;      /* Lt C_e h^T */
        C_z[i * nz + nu_] = lc;
        C_z[nu_ * nz + i] = Ce00 * Lt[i * 2] + Ce01 * Lt[i * 2 + 1];     /* h C_e Lt^T */
    }
    C_z[nu_ * nz + nu_] = Ce00;
    /* write back (:459-479) */
    for (int i = 0; i < M; ++i) { mu_g[i] = mu_z[i]; for (int j = 0; j < M; ++j) C_g[i * M + j] = C_z[i * nz + j]; }
    for (int k = 0; k < 3; ++k) { mu_eta[k] = mu_z[M + k]; for (int l = 0; l < 3; ++l) C_eta[k * 3 + l] = C_z[(M + k) * nz + (M + l)]; }
    if (mu_z_out) memcpy(mu_z_out, mu_z, sizeof(double) * nz);
    if (C_z_out) memcpy(C_z_out, C_z, sizeof(double) * nz * nz);
    double th[3] = {mu_eta[0], mu_eta[1], mu_eta[2]};
    double *Kx = (double *)malloc(sizeof(double) * M * M);
    int rc = orc_rgp_prior(M, X, th, Kx, Kx_inv);
    free(Kx); free(cj); free(buf);
    return rc;
}

void orc_rgp_alpha(int M, const double *Kx_inv, const double *mu, double *alpha)
{
    for (int i = 0; i < M; ++i) { double s = 0; for (int j = 0; j < M; ++j) s += Kx_inv[i * M + j] * mu[j]; alpha[i] = s; }
}

/* posterior mean and variance at m query points (RGP.py:195-210) */
void orc_rgp_predict(int M, const double *X, const double *th, const double *Kx_inv,
                     const double *mu, const double *C, int m, const double *xs, double *mean, double *var)
{
    double *kv = (double *)malloc(sizeof(double) * 2 * M), *Jt = kv + M;
    for (int t = 0; t < m; ++t) {
        for (int i = 0; i < M; ++i) kv[i] = rbf(xs[t], X[i], th[0], th[1]);
        for (int j = 0; j < M; ++j) { double s = 0; for (int i = 0; i < M; ++i) s += kv[i] * Kx_inv[i * M + j]; Jt[j] = s; }
        double mp = 0, jk = 0, jcj = 0;
        for (int j = 0; j < M; ++j) { mp += Jt[j] * mu[j]; jk += Jt[j] * rbf(X[j], xs[t], th[0], th[1]); }
        for (int j = 0; j < M; ++j) { double s = 0; for (int i = 0; i < M; ++i) s += Jt[i] * C[i * M + j]; jcj += s * Jt[j]; }
        mean[t] = mp;
        if (var) var[t] = rbf(xs[t], xs[t], th[0], th[1]) - jk + jcj;
    }
    free(kv);
}

/* ------------------------------------------------ residual, plant, chunk */

void orc_compute_a_drag(const double *x_now, const double *x_pred, double dt, double *v_body, double *a_drag)
{
    double R[3][3], Rp[3][3];
    rotmat(x_now + 3, R); rotmat(x_pred + 3, Rp);
    for (int i = 0; i < 3; ++i) {
        double vb = R[0][i] * x_now[7] + R[1][i] * x_now[8] + R[2][i] * x_now[9];
        double vp = Rp[0][i] * x_pred[7] + Rp[1][i] * x_pred[8] + Rp[2][i] * x_pred[9];
        v_body[i] = vb; a_drag[i] = (vb - vp) / dt;
    }
}

/* plant[4] = aero_drag, rotor_drag_x, rotor_drag_y, rotor_drag_z (quad.py:79-88, drag=True, payload=False) */
static void plant_f(const double *quad, const double *plant, const double *x, const double *u, double *f)
{
    model_t m = mk_model(quad, 0, NULL, NULL, NULL);
    evalpt_t e; f_point(&m, x, u, &e);
    memcpy(f, e.f, sizeof(double) * NX);
    double ad[3];
    for (int i = 0; i < 3; ++i) {
        double vb = e.vb[i], sg = (vb > 0) - (vb < 0);
        ad[i] = -plant[0] * vb * vb * sg / quad[0] - plant[1 + i] * vb / quad[0];
    }
    for (int i = 0; i < 3; ++i) f[7 + i] += e.R[i][0] * ad[0] + e.R[i][1] * ad[1] + e.R[i][2] * ad[2];
}

void orc_plant_update(const double *quad, const double *plant, double *x, const double *u_in, double dt)
{
    double u[NU], k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
    for (int i = 0; i < NU; ++i) u[i] = u_in[i] < 0 ? 0 : (u_in[i] > 1 ? 1 : u_in[i]);
    plant_f(quad, plant, x, u, k1);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt / 2 * k1[i];
    plant_f(quad, plant, xs, u, k2);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt / 2 * k2[i];
    plant_f(quad, plant, xs, u, k3);
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt * k3[i];
    plant_f(quad, plant, xs, u, k4);
    for (int i = 0; i < NX; ++i) x[i] += dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* plant advance over one control period: `while control_time < dt: update(w, sim_dt)`
   (execute_trajectory.py:232-243; float accumulation decides the sub-step count). returns sub-steps */
int orc_plant_period(const double *quad, const double *plant, double *x, const double *u, double dt, double sim_dt)
{
    double t = 0; int n = 0;
    while (t < dt) { orc_plant_update(quad, plant, x, u, sim_dt); t += sim_dt; ++n; }
    return n;
}

/* utils.get_reference_chunk (utils.py:897-931): traj [K][13] -> out [N][13] */
void orc_reference_chunk(const double *traj, int K, int idx, int N, int skip, double *out)
{
    int left = K - idx;
    if (left > N * skip) {
        for (int j = 0; j < N; ++j) memcpy(out + j * NX, traj + (size_t)(idx + j * skip) * NX, sizeof(double) * NX);
    } else if (left > skip - 1) {
        int n = 0;
        for (int s = idx; s < idx + left * skip && s < K && n < N; s += skip, ++n)
            memcpy(out + n * NX, traj + (size_t)s * NX, sizeof(double) * NX);
        for (; n < N; ++n) memcpy(out + n * NX, traj + (size_t)(K - 1) * NX, sizeof(double) * NX);
    } else {
        for (int j = 0; j < N; ++j) memcpy(out + j * NX, traj + (size_t)(K - 1) * NX, sizeof(double) * NX);
    }
}

/* ------------------------------------------------------------ closed loop */

/* One vehicle's persistent controller state */
typedef struct {
    double *xit, *uit;          /* RTI iterate */
    double *mu, *C, *alpha;     /* RGP state [3][M], [3][M][M], [3][M] */
    double x_pred_prev[NX];
    int have_pred;
} veh_t;

/*
 * Closed loop for B independent vehicles over `steps` control steps, in the order of
 * execute_trajectory.py:196-277:  chunk -> RTI solve (params of the previous step) -> u0 ->
 * nominal RK4 prediction -> plant period -> drag residual -> RGP regress x3 -> alpha.
 *   traj [B][K][13], x [B][13] in/out (plant state),
 *   xit [B][(N+1)*13], uit [B][N*4], mu [B][3*M], C [B][3*M*M] persistent (in/out),
 *   xpred_prev [B][13] in/out with have_pred flag [B] (int).
 *   u0_log (optional) [steps][B][4], x_log (optional) [steps][B][13] (state seen by the controller)
 * use_gp: 0 nominal model, 1 RGP in the loop.
 */
int orc_closed_loop(const double *quad, const double *plant, double dt, double sim_dt, int N,
                    int M, const double *gpX, const double *gpth, const double *Kx_inv, int use_gp,
                    const double *Wd, const double *Wed, double u_ref,
                    int B, int K, const double *traj, int step0, int steps,
                    double *x, double *xit, double *uit, double *mu, double *C,
                    double *xpred_prev, int *have_pred,
                    double *u0_log, double *x_log, double *cost_log, int *iters_log,
                    double mu_tol, int max_iter, int do_polish, int nthreads, int reset_on_fail)
{
    /* reset_on_fail = 0: the reference's behaviour, the iterate is whatever the last solve left (quad_opt.py:333 ignores
     * the solver status).  reset_on_fail = 1 mirrors the library's per-handle option (include/qmpc.h reset_on_fail). */
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int b = 0; b < B; ++b) {
        double *yref = (double *)malloc(sizeof(double) * (N * NZ + NX));
        double *chunk = (double *)malloc(sizeof(double) * N * NX);
        double *alpha = (double *)calloc(3 * (M > 0 ? M : 1), sizeof(double));
        double *xb = x + (size_t)b * NX;
        double *xi = xit + (size_t)b * (N + 1) * NX, *ui = uit + (size_t)b * N * NU;
        double *mub = mu ? mu + (size_t)b * 3 * M : NULL, *Cb = C ? C + (size_t)b * 3 * M * M : NULL;
        for (int s = 0; s < steps; ++s) {
            int i = step0 + s;
            orc_reference_chunk(traj + (size_t)b * K * NX, K, i, N, 1, chunk);
            for (int k = 0; k < N; ++k) {
                memcpy(yref + k * NZ, chunk + k * NX, sizeof(double) * NX);
                for (int a = 0; a < NU; ++a) yref[k * NZ + NX + a] = u_ref;
            }
            const double *yref_e = chunk + (N - 1) * NX;
            if (use_gp) for (int d = 0; d < 3; ++d) orc_rgp_alpha(M, Kx_inv + d * M * M, mub + d * M, alpha + d * M);
            if (reset_on_fail && have_pred[b] >= 2) {      /* previous solve failed: restart the SQP iterate on the new reference */
                for (int k = 0; k < N; ++k) {
                    memcpy(xi + k * NX, chunk + k * NX, sizeof(double) * NX);
                    for (int a = 0; a < NU; ++a) ui[k * NU + a] = u_ref;
                }
                memcpy(xi + N * NX, yref_e, sizeof(double) * NX);
            }
            double xnow[NX], cost, kkt; int iters;
            memcpy(xnow, xb, sizeof(xnow));
            int st = orc_rti_step(quad, dt, N, use_gp ? M : 0, gpX, gpth, use_gp ? alpha : NULL, Wd, Wed, 0.0, 1.0,
                                  xnow, yref, yref_e, xi, ui, &cost, &iters, &kkt, mu_tol,
                                  (reset_on_fail && have_pred[b] >= 3 && ((have_pred[b] - 1) & 7) != 0 && max_iter > 20) ? 20 : max_iter,   /* two failures in a row: bounded attempt, full again every 8th */
                                  do_polish, NULL);
            if (st > 1) bad += 1;
            double u0[NU], xpred[NX];
            memcpy(u0, ui, sizeof(u0));
            orc_rk4(quad, 0, NULL, NULL, NULL, xnow, u0, dt, xpred);
            if (u0_log) memcpy(u0_log + ((size_t)s * B + b) * NU, u0, sizeof(u0));
            if (x_log) memcpy(x_log + ((size_t)s * B + b) * NX, xnow, sizeof(xnow));
            if (cost_log) cost_log[(size_t)s * B + b] = cost;
            if (iters_log) iters_log[(size_t)s * B + b] = iters;
            orc_plant_period(quad, plant, xb, u0, dt, sim_dt);
            if (use_gp) {
                double vb[3], ad[3];
                const double *xpm = have_pred[b] ? xpred_prev + (size_t)b * NX : xnow;
                orc_compute_a_drag(xnow, xpm, dt, vb, ad);
                for (int d = 0; d < 3; ++d)
                    orc_rgp_regress(M, gpX + d * M, gpth + 3 * d, Kx_inv + d * M * M, mub + d * M, Cb + d * M * M, vb[d], ad[d]);
            }
            memcpy(xpred_prev + (size_t)b * NX, xpred, sizeof(xpred));
            have_pred[b] = st > 1 ? (have_pred[b] >= 2 ? have_pred[b] + 1 : 2) : 1;   /* 1 ok, 1 + number of consecutive failures */
        }
        free(yref); free(chunk); free(alpha);
    }
    return bad;
}

/* B independent RTI steps (one per vehicle, OpenMP over vehicles): the per-step parity reference for a batch with
 * injected iterates.  Arrays are vehicle-major like the C-ABI of the library; alpha may be NULL (M == 0); alpha_stride
 * doubles between vehicles (0 = one shared model).  status/iters/cost per vehicle. */
int orc_rti_step_batch(const double *quad, double dt, int N, int M, const double *gpX, const double *gpth,
                       const double *alpha, int alpha_stride, const double *Wd, const double *Wed, double lbu, double ubu,
                       int B, const double *x0, const double *yref, const double *yref_e, double *xit, double *uit,
                       double *cost, int *iters, int *status, double mu_tol, int max_iter, int do_polish, int nthreads)
{
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int b = 0; b < B; ++b) {
        double kkt;
        int st = orc_rti_step(quad, dt, N, M, gpX, gpth, alpha ? alpha + (size_t)b * alpha_stride : NULL, Wd, Wed, lbu, ubu,
                              x0 + (size_t)b * NX, yref + (size_t)b * N * NZ, yref_e + (size_t)b * NX,
                              xit + (size_t)b * (N + 1) * NX, uit + (size_t)b * N * NU, cost + b, iters + b, &kkt,
                              mu_tol, max_iter, do_polish, NULL);
        status[b] = st;
        if (st > 1) bad += 1;
    }
    return bad;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
