// mpc_device.cuh -- per-problem Levenberg-Marquardt solve of DR-MPC's path-tracking MPC in fp64.
//
// What it computes (reference environment/path_tracking/utils/mpc.py:58-206 and
// pysteam_augmented/lev_marq_gauss_newton_custom_solver.py:30-171): K planar poses + K (v,w) controls,
// pose-tracking / unicycle-kinematics / 4 clamped-log-barrier residuals per step, damped Gauss-Newton
// with whole-step box-feasibility scaling and two-trial backtracking, STEAM's termination order,
// pose_error_scaling*=5 retry, un-shifted-pose / shifted-control warm start.
//
// How: the reference's 8K-dim SE(3) problem is solved as the identical 5K-dim SE(2) problem
// (in-plane / out-of-plane coordinates decouple exactly for planar poses). Unknowns are ordered
// z_j = (u_{j-1}, delta_j), j = 1..K, which makes the normal equations block-tridiagonal with 5x5 blocks
// (packed symmetric diagonal blocks, 5x3 coupling blocks), factorised by a block Cholesky in
// registers / local memory. One thread owns one problem; no tensor cores: the systems are tiny,
// and every problem takes its own data-dependent path through the LM loop.
#pragma once
#include "dre_common.cuh"

// horizon loops are fully unrolled only for the small fixed horizons
#define DRE_UNROLL_K _Pragma("unroll (KC <= 4 ? KC : 1)")

struct Se2 { double c, s, x, y; };   // [[c,-s,x],[s,c,y],[0,0,1]]

DRE_HD Se2 se2_mul(const Se2 &a, const Se2 &b)
{
    Se2 r;
    r.c = a.c * b.c - a.s * b.s;
    r.s = a.s * b.c + a.c * b.s;
    r.x = a.c * b.x - a.s * b.y + a.x;
    r.y = a.s * b.x + a.c * b.y + a.y;
    return r;
}
DRE_HD Se2 se2_inv(const Se2 &a)
{
    Se2 r;
    r.c = a.c; r.s = -a.s;
    r.x = -(a.c * a.x + a.s * a.y);
    r.y = a.s * a.x - a.c * a.y;
    return r;
}
// a * b^-1
DRE_HD Se2 se2_mul_inv(const Se2 &a, const Se2 &b)
{
    Se2 r;
    r.c = a.c * b.c + a.s * b.s;
    r.s = a.s * b.c - a.c * b.s;
    r.x = a.x - (r.c * b.x - r.s * b.y);
    r.y = a.y - (r.s * b.x + r.c * b.y);
    return r;
}
DRE_HD Se2 se2_from_pose_inv(double x, double y, double th)
{
    double sn, cs;
    sincos(th, &sn, &cs);
    Se2 r;
    r.c = cs; r.s = -sn;
    r.x = -(cs * x + sn * y);
    r.y = sn * x - cs * y;
    return r;
}

// coefficients of the SE(2) exponential / left Jacobian for angle p with sin/cos already known:
//   A = sin p / p, B = (1 - cos p)/p, beta = (1 - cos p)/p^2, alpha = (p - sin p)/p^2
struct PhiCoef { double A, B, alpha, beta; };
DRE_HD PhiCoef phi_coef(double p, double sn, double cs)
{
    PhiCoef q;
    if (fabs(p) < 1e-3) {
        const double p2 = p * p;
        q.A = 1.0 - p2 * (1.0 / 6.0) + p2 * p2 * (1.0 / 120.0);
        q.beta = 0.5 - p2 * (1.0 / 24.0) + p2 * p2 * (1.0 / 720.0);
        q.B = p * q.beta;
        q.alpha = p * (1.0 / 6.0 - p2 * (1.0 / 120.0) + p2 * p2 * (1.0 / 5040.0));
    } else {
        const double ip = 1.0 / p;
        q.A = sn * ip;
        q.B = (1.0 - cs) * ip;
        q.beta = q.B * ip;
        q.alpha = (p - sn) * ip * ip;
    }
    return q;
}
// (p/2) cot(p/2) from sin/cos of p
DRE_HD double half_cot(double p, double sn, double cs)
{
    if (fabs(p) < 1e-3) { const double p2 = p * p; return 1.0 - p2 * (1.0 / 12.0) - p2 * p2 * (1.0 / 720.0); }
    return cs >= 0.0 ? 0.5 * p * (1.0 + cs) / sn : 0.5 * p * sn / (1.0 - cs);
}
// exp of (rx, 0, p): the only twist shape the kinematic term needs (mpc.py:32,108-109)
DRE_HD Se2 se2_exp_x(double rx, double p, double sn, double cs, const PhiCoef &q)
{
    Se2 r;
    r.c = cs; r.s = sn; r.x = q.A * rx; r.y = q.B * rx;
    return r;
}
DRE_HD Se2 se2_exp(double rx, double ry, double p)
{
    double sn, cs;
    sincos(p, &sn, &cs);
    const PhiCoef q = phi_coef(p, sn, cs);
    Se2 r;
    r.c = cs; r.s = sn;
    r.x = q.A * rx - q.B * ry;
    r.y = q.B * rx + q.A * ry;
    return r;
}
// log map; also returns (p/2)cot(p/2) so Jacobians can reuse it
DRE_HD void se2_log(const Se2 &T, double e[3], double &ct)
{
    const double p = atan2(T.s, T.c);
    ct = half_cot(p, T.s, T.c);
    const double h = 0.5 * p;
    e[0] = ct * T.x + h * T.y;
    e[1] = ct * T.y - h * T.x;
    e[2] = p;
}

// 2x3 "affine" matrices: rows 0,1 of a 3x3 whose last row is (0,0,1)
struct Aff { double m[6]; };   // m[0..2] row 0, m[3..5] row 1
DRE_HD Aff aff_mul(const Aff &a, const Aff &b)
{
    Aff c;
    c.m[0] = a.m[0] * b.m[0] + a.m[1] * b.m[3];
    c.m[1] = a.m[0] * b.m[1] + a.m[1] * b.m[4];
    c.m[2] = a.m[0] * b.m[2] + a.m[1] * b.m[5] + a.m[2];
    c.m[3] = a.m[3] * b.m[0] + a.m[4] * b.m[3];
    c.m[4] = a.m[3] * b.m[1] + a.m[4] * b.m[4];
    c.m[5] = a.m[3] * b.m[2] + a.m[4] * b.m[5] + a.m[5];
    return c;
}
DRE_HD Aff se2_Ad(const Se2 &T)
{
    Aff a;
    a.m[0] = T.c; a.m[1] = -T.s; a.m[2] = T.y;
    a.m[3] = T.s; a.m[4] = T.c;  a.m[5] = -T.x;
    return a;
}
// J_l(e)^-1 for e with rotation p whose sin/cos are (sn,cs); ct = (p/2)cot(p/2)
DRE_HD Aff se2_Jlinv(const double e[3], double sn, double cs, double ct)
{
    const PhiCoef q = phi_coef(e[2], sn, cs);
    const double h = 0.5 * e[2];
    const double qx = q.alpha * e[0] + q.beta * e[1], qy = q.alpha * e[1] - q.beta * e[0];
    Aff a;
    a.m[0] = ct; a.m[1] = h;  a.m[2] = -(ct * qx + h * qy);
    a.m[3] = -h; a.m[4] = ct; a.m[5] = h * qx - ct * qy;
    return a;
}

DRE_HD double barrier_val(double x) { return x <= 0.0 ? -500.0 : (x > 1.0 ? 500.0 : log(x)); }
DRE_HD double barrier_grad(double x) { return (x <= 0.0 || x > 1.0) ? 10.0 : 1.0 / x; }

DRE_HD int sym_idx(int a, int c) { return a >= c ? a * (a + 1) / 2 + c : c * (c + 1) / 2 + a; }

template <int KC>
struct MpcProblem {
    Se2 T[KC + 1];      // T[0] = locked current (inverse) pose
    double u[KC][2];
    Se2 M[KC];          // inverse reference poses
};

template <int KC>
struct MpcSystem {
    double D[KC][15];   // packed lower-symmetric 5x5 diagonal blocks
    double O[KC][15];   // coupling block rows z_j (5) x cols delta_{j-1} (3); O[0] unused
    double b[KC][5];
    double Lf[KC][15];  // Cholesky factors of the Schur complements (packed lower)
    double Bf[KC][15];  // O_j L_{j-1}^{-T}, 5x3 (first two columns are identically zero)
};

struct MpcConst {
    int K;
    double dt, vmax, wmax, inv_vmax, inv_2wmax;
    int pose_jac_exact;
};

// cost = sum over terms of 1/2 |whitened error|^2, in the reference's term order (mpc.py:99-142)
template <int KC>
DRE_HD double mpc_cost(const MpcConst &C, const MpcProblem<KC> &P, double s)
{
    double cost = 0.0;
    const double inv_s = 1.0 / s;
DRE_UNROLL_K
    for (int k = 0; k < KC; ++k) {
        if (k >= C.K) break;
        double e[3], ct;
        se2_log(se2_mul_inv(P.M[k], P.T[k + 1]), e, ct);
        cost += 0.5 * inv_s * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        const double p = C.dt * P.u[k][1];
        double sn, cs;
        sincos(p, &sn, &cs);
        const PhiCoef q = phi_coef(p, sn, cs);
        const Se2 Xinv = se2_exp_x(C.dt * P.u[k][0], p, sn, cs, q);   // exp(xi)^-1 = exp(-xi), xi = -dt (v,0,w)
        const Se2 E = se2_mul(se2_mul_inv(P.T[k + 1], P.T[k]), Xinv);
        se2_log(E, e, ct);
        cost += 0.5 * 1000.0 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        const double v = P.u[k][0], w = P.u[k][1];
        const double b0 = barrier_val((C.vmax - v) * C.inv_vmax), b1 = barrier_val(v * C.inv_vmax);
        const double b2 = barrier_val((C.wmax - w) * C.inv_2wmax), b3 = barrier_val((w + C.wmax) * C.inv_2wmax);
        cost += 0.5 * 0.1 * (1.0 / 250000.0) * (b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
    }
    return cost;
}

// accumulate w * J^T J (packed) and -w * J^T e for a 3-row Jacobian whose rows 0,1 are `r0`,`r1` (n cols) and whose
// row 2 is `r2`, into diagonal-block rows/cols [off, off+n)
template <int N>
DRE_HD void acc_block(double *D, double *b, int off, const double *r0, const double *r1, const double *r2,
                      const double e[3], double w)
{
#pragma unroll
    for (int a = 0; a < N; ++a) {
        b[off + a] -= w * (r0[a] * e[0] + r1[a] * e[1] + r2[a] * e[2]);
#pragma unroll
        for (int c = 0; c <= a; ++c)
            D[sym_idx(off + a, off + c)] += w * (r0[a] * r0[c] + r1[a] * r1[c] + r2[a] * r2[c]);
    }
}

template <int KC>
DRE_HD void mpc_build(const MpcConst &C, const MpcProblem<KC> &P, double s, MpcSystem<KC> &S)
{
DRE_UNROLL_K
    for (int k = 0; k < KC; ++k) {
        if (k >= C.K) break;
#pragma unroll
        for (int i = 0; i < 15; ++i) { S.D[k][i] = 0.0; S.O[k][i] = 0.0; }
#pragma unroll
        for (int i = 0; i < 5; ++i) S.b[k][i] = 0.0;
    }
DRE_UNROLL_K
    for (int k = 0; k < KC; ++k) {
        if (k >= C.K) break;
        // ---- pose term: e = ln(M_k T_{k+1}^-1), J = -J_l^-1(e) [STEAM] or -J_l^-1(e) Ad(exp e) [exact] ----
        {
            double e[3], ct;
            const Se2 E = se2_mul_inv(P.M[k], P.T[k + 1]);
            se2_log(E, e, ct);
            Aff G = se2_Jlinv(e, E.s, E.c, ct);
            if (C.pose_jac_exact) G = aff_mul(G, se2_Ad(E));
            const double r0[3] = {-G.m[0], -G.m[1], -G.m[2]}, r1[3] = {-G.m[3], -G.m[4], -G.m[5]}, r2[3] = {0.0, 0.0, -1.0};
            acc_block<3>(S.D[k], S.b[k], 2, r0, r1, r2, e, 3.0 / s);
        }
        // ---- kinematic term: e = ln(T_{k+1} T_k^-1 exp(xi)^-1), xi = -dt (v, 0, w) ----
        {
            const double p = C.dt * P.u[k][1];
            double sn, cs;
            sincos(p, &sn, &cs);
            const PhiCoef q = phi_coef(p, sn, cs);
            const double rx = C.dt * P.u[k][0];
            const Se2 Lh = se2_mul_inv(P.T[k + 1], P.T[k]);
            const Se2 E = se2_mul(Lh, se2_exp_x(rx, p, sn, cs, q));
            double e[3], ct;
            se2_log(E, e, ct);
            const Aff G = se2_Jlinv(e, E.s, E.c, ct);
            // J_l(xi) for xi = (-rx, 0, -p): A(-p)=A, B(-p)=-B, alpha(-p)=-alpha, beta(-p)=beta
            Aff Jl;
            Jl.m[0] = q.A; Jl.m[1] = q.B;  Jl.m[2] = q.alpha * rx;      // alpha(-p) * (-rx)
            Jl.m[3] = -q.B; Jl.m[4] = q.A; Jl.m[5] = q.beta * rx;       // -beta * (-rx)
            const Aff GA = aff_mul(G, se2_Ad(E));
            const Aff Ju = aff_mul(GA, Jl);
            // d e / d u = -G Ad(E) J_l(xi) dt P, P = -[e1 e6]  ->  +dt * columns (0, 2)
            const double r0[5] = {C.dt * Ju.m[0], C.dt * Ju.m[2], G.m[0], G.m[1], G.m[2]};
            const double r1[5] = {C.dt * Ju.m[3], C.dt * Ju.m[5], G.m[3], G.m[4], G.m[5]};
            const double r2[5] = {0.0, C.dt, 0.0, 0.0, 1.0};
            acc_block<5>(S.D[k], S.b[k], 0, r0, r1, r2, e, 1000.0);
            if (k >= 1) {
                const Aff GL = aff_mul(G, se2_Ad(Lh));
                const double q0[3] = {-GL.m[0], -GL.m[1], -GL.m[2]}, q1[3] = {-GL.m[3], -GL.m[4], -GL.m[5]}, q2[3] = {0.0, 0.0, -1.0};
                acc_block<3>(S.D[k - 1], S.b[k - 1], 2, q0, q1, q2, e, 1000.0);
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        S.O[k][3 * a + c] = 1000.0 * (r0[a] * q0[c] + r1[a] * q1[c] + r2[a] * q2[c]);
            }
        }
        // ---- four clamped log barriers on (v, w) (mpc.py:114-142, extra_se3_evaluators.py:20-44) ----
        {
            const double v = P.u[k][0], w = P.u[k][1];
            const double x0 = (C.vmax - v) * C.inv_vmax, x1 = v * C.inv_vmax;
            const double x2 = (C.wmax - w) * C.inv_2wmax, x3 = (w + C.wmax) * C.inv_2wmax;
            const double k500 = 1.0 / 500.0;
            const double e0 = barrier_val(x0) * k500, e1 = barrier_val(x1) * k500;
            const double e2 = barrier_val(x2) * k500, e3 = barrier_val(x3) * k500;
            const double j0 = -k500 * barrier_grad(x0) * C.inv_vmax, j1 = k500 * barrier_grad(x1) * C.inv_vmax;
            const double j2 = -k500 * barrier_grad(x2) * C.inv_2wmax, j3 = k500 * barrier_grad(x3) * C.inv_2wmax;
            S.b[k][0] -= 0.1 * (j0 * e0 + j1 * e1);
            S.b[k][1] -= 0.1 * (j2 * e2 + j3 * e3);
            S.D[k][sym_idx(0, 0)] += 0.1 * (j0 * j0 + j1 * j1);
            S.D[k][sym_idx(1, 1)] += 0.1 * (j2 * j2 + j3 * j3);
        }
    }
}

// Damped block-tridiagonal Cholesky: (A with diag*(1+lambda)) d = b. false on a non-positive pivot.
template <int KC>
DRE_HD bool mpc_solve(const MpcConst &C, MpcSystem<KC> &S, double lambda, double d[KC][5])
{
    double y[KC][5];
    const double damp = 1.0 + lambda;
DRE_UNROLL_K
    for (int j = 0; j < KC; ++j) {
        if (j >= C.K) break;
        double Sj[15], rhs[5];
#pragma unroll
        for (int i = 0; i < 15; ++i) Sj[i] = S.D[j][i];
#pragma unroll
        for (int a = 0; a < 5; ++a) { Sj[sym_idx(a, a)] = S.D[j][sym_idx(a, a)] * damp; rhs[a] = S.b[j][a]; }
        if (j > 0) {
            const double *Lp = S.Lf[j - 1];
            // B L^T = O with O's first two columns zero -> B's first two columns zero
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const double b2 = S.O[j][3 * a + 0] / Lp[sym_idx(2, 2)];
                const double b3 = (S.O[j][3 * a + 1] - b2 * Lp[sym_idx(3, 2)]) / Lp[sym_idx(3, 3)];
                const double b4 = (S.O[j][3 * a + 2] - b2 * Lp[sym_idx(4, 2)] - b3 * Lp[sym_idx(4, 3)]) / Lp[sym_idx(4, 4)];
                S.Bf[j][3 * a + 0] = b2; S.Bf[j][3 * a + 1] = b3; S.Bf[j][3 * a + 2] = b4;
            }
#pragma unroll
            for (int a = 0; a < 5; ++a) {
#pragma unroll
                for (int c = 0; c <= a; ++c)
                    Sj[sym_idx(a, c)] -= S.Bf[j][3 * a] * S.Bf[j][3 * c] + S.Bf[j][3 * a + 1] * S.Bf[j][3 * c + 1] +
                                         S.Bf[j][3 * a + 2] * S.Bf[j][3 * c + 2];
                rhs[a] -= S.Bf[j][3 * a] * y[j - 1][2] + S.Bf[j][3 * a + 1] * y[j - 1][3] + S.Bf[j][3 * a + 2] * y[j - 1][4];
            }
        }
        double *L = S.Lf[j];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            double dd = Sj[sym_idx(c, c)];
#pragma unroll
            for (int m = 0; m < c; ++m) dd -= L[sym_idx(c, m)] * L[sym_idx(c, m)];
            if (!(dd > 0.0)) return false;
            const double ld = sqrt(dd), ild = 1.0 / ld;
            L[sym_idx(c, c)] = ld;
#pragma unroll
            for (int a = c + 1; a < 5; ++a) {
                double v = Sj[sym_idx(a, c)];
#pragma unroll
                for (int m = 0; m < c; ++m) v -= L[sym_idx(a, m)] * L[sym_idx(c, m)];
                L[sym_idx(a, c)] = v * ild;
            }
        }
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            double v = rhs[a];
#pragma unroll
            for (int m = 0; m < a; ++m) v -= L[sym_idx(a, m)] * y[j][m];
            y[j][a] = v / L[sym_idx(a, a)];
        }
    }
DRE_UNROLL_K
    for (int jj = 0; jj < KC; ++jj) {
        const int j = C.K - 1 - jj;
        if (j < 0) break;
        double rhs[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) rhs[a] = y[j][a];
        if (j + 1 < C.K) {
            // rhs -= B_{j+1}^T d_{j+1}; only rows 2..4 of z_j couple
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 5; ++m) v += S.Bf[j + 1][3 * m + c] * d[j + 1][m];
                rhs[2 + c] -= v;
            }
        }
        const double *L = S.Lf[j];
#pragma unroll
        for (int a = 4; a >= 0; --a) {
            double v = rhs[a];
#pragma unroll
            for (int m = a + 1; m < 5; ++m) v -= L[sym_idx(m, a)] * d[j][m];
            d[j][a] = v / L[sym_idx(a, a)];
        }
    }
    return true;
}

// b^T p - 1/2 p^T A p with the undamped A (lev_marq...py:167-171)
template <int KC>
DRE_HD double mpc_predict(const MpcConst &C, const MpcSystem<KC> &S, const double d[KC][5])
{
    double g = 0.0, q = 0.0;
DRE_UNROLL_K
    for (int j = 0; j < KC; ++j) {
        if (j >= C.K) break;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            g += S.b[j][a] * d[j][a];
            double r = 0.0;
#pragma unroll
            for (int c = 0; c < 5; ++c) r += S.D[j][sym_idx(a, c)] * d[j][c];
            if (j > 0) {
                double o = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) o += S.O[j][3 * a + c] * d[j - 1][2 + c];
                r += 2.0 * o;
            }
            q += d[j][a] * r;
        }
    }
    return g - 0.5 * q;
}

struct MpcResult { int iterations, lm_solves, cost_evals, term; double final_cost; };

// One optimize() at fixed pose_error_scaling s. term > 0: terminated normally; 0: unsuccessful step
// (pysteam raises -> mpc.py:188-190 retries with s*=5).
template <int KC>
DRE_HD int mpc_optimize(const MpcConst &C, const dre_mpc_params &prm, MpcProblem<KC> &P, double s, MpcSystem<KC> &S,
                        MpcResult &R)
{
    double lambda = 1e-7;            // lev_marq...py:28, reset for every solver object
    double curr = mpc_cost(C, P, s), prev;
    int iter = 0;
    for (;;) {
        ++iter;
        prev = curr;
        mpc_build(C, P, s, S);
        double gn = 0.0;
DRE_UNROLL_K
        for (int j = 0; j < KC; ++j) {
            if (j >= C.K) break;
#pragma unroll
            for (int a = 0; a < 5; ++a) gn += S.b[j][a] * S.b[j][a];
        }
        bool success = false;
        double new_cost = prev;
        for (int nb = 0; nb < 50 && !success; ++nb) {
            double d[KC][5];
            R.lm_solves++;
            if (mpc_solve(C, S, lambda, d)) {
                // whole-step feasibility scale (lev_marq...py:52-95): walk u_0..u_{K-1}, v then w
                double scale = 1.0;
DRE_UNROLL_K
                for (int k = 0; k < KC; ++k) {
                    if (k >= C.K) break;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double dp = d[k][c];
                        if (dp != 0.0) {
                            const double hi = c == 0 ? C.vmax : C.wmax, lo = c == 0 ? 0.0 : -C.wmax;
                            const double maxp = dp > 0.0 ? (hi - 1e-7) - P.u[k][c] : (lo + 1e-7) - P.u[k][c];
                            const double allowed = maxp / dp;
                            scale = fmin(scale, allowed);
                            if (allowed < 1.0) scale = fmax(scale - 1e-7, 0.0);
                        }
                    }
                }
DRE_UNROLL_K
                for (int k = 0; k < KC; ++k) {
                    if (k >= C.K) break;
#pragma unroll
                    for (int a = 0; a < 5; ++a) d[k][a] *= scale;
                }
                for (int bt = 0; bt < 2; ++bt) {
                    if (bt > 0) {
DRE_UNROLL_K
                        for (int k = 0; k < KC; ++k) {
                            if (k >= C.K) break;
#pragma unroll
                            for (int a = 0; a < 5; ++a) d[k][a] *= 0.8;
                        }
                    }
                    MpcProblem<KC> trial;
                    trial.T[0] = P.T[0];
DRE_UNROLL_K
                    for (int k = 0; k < KC; ++k) {
                        if (k >= C.K) break;
                        trial.M[k] = P.M[k];
                        trial.u[k][0] = P.u[k][0] + d[k][0];
                        trial.u[k][1] = P.u[k][1] + d[k][1];
                        trial.T[k + 1] = se2_mul(se2_exp(d[k][2], d[k][3], d[k][4]), P.T[k + 1]);
                    }
                    const double proposed = mpc_cost(C, trial, s);
                    R.cost_evals++;
                    const double actual = prev - proposed;
                    const double pred = mpc_predict(C, S, d);
                    const double ratio = (pred == 0.0) ? 0.0 : actual / pred;
                    if (ratio > 0.25) {
DRE_UNROLL_K
                        for (int k = 0; k < KC; ++k) {
                            if (k >= C.K) break;
                            P.u[k][0] = trial.u[k][0]; P.u[k][1] = trial.u[k][1]; P.T[k + 1] = trial.T[k + 1];
                        }
                        lambda = fmax(lambda * 0.1, 1e-7);
                        new_cost = proposed;
                        success = true;
                        break;
                    }
                }
                if (!success) {
                    // at the damping cap a failed attempt repeats itself verbatim: fast-forward (see mpc_coop.cuh)
                    if (lambda >= 1e7) { const int rem = 50 - (nb + 1); R.lm_solves += rem; R.cost_evals += 2 * rem; nb = 49; }
                    lambda = fmin(lambda * 10.0, 1e7);
                }
            } else {
                if (lambda >= 1e7) { R.lm_solves += 50 - (nb + 1); nb = 49; }
                lambda = fmin(lambda * 10.0, 1e7);
            }
        }
        curr = new_cost;
        R.iterations = iter;
        R.final_cost = curr;
        if (!success && sqrt(gn) < 1e-6) return 1;
        if (!success) return 0;
        if (iter >= prm.max_iterations) return 3;
        if (curr <= 0.0) return 4;
        if (fabs(prev - curr) <= prm.abs_change_tol) return 5;
        if (fabs(prev - curr) / prev <= prm.rel_change_tol) return 6;
    }
}

// mpc.py:208-237: reference k at node index interp + (k+1)*inc; component-wise lerp (theta included, no wrap
// handling), clamped to the last node. path rows: x, y, theta with row stride `stride`.
DRE_HD void mpc_ref_pose(const double *path, int stride, int L, double target, double out[3])
{
    if (target >= (double)(L - 1)) {
        out[0] = path[L - 1]; out[1] = path[stride + L - 1]; out[2] = path[2 * stride + L - 1];
    } else {
        const double fl = floor(target);
        const int lo = (int)fl, hi = (int)ceil(target);
        const double f = target - fl;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double a = path[r * stride + lo], b = path[r * stride + hi];
            out[r] = a + f * (b - a);
        }
    }
}

// MPC.act for one problem. warm_T/warm_u are addressed with element stride `ws` (horizon-major SoA on device:
// element (k, n) at [k*ws + n]); it0 is this problem's iteration0 flag.
template <int KC>
DRE_HD void mpc_act_one(const dre_mpc_params &prm, const double *path, int stride, int L, double interp_idx,
                        double px, double py, double pth, double4 *warm_T, double2 *warm_u, size_t ws,
                        uint8_t *it0, double *controls, int ncontrols, dre_mpc_status *status_out, MpcSystem<KC> &S)
{
    MpcConst C;
    C.K = prm.horizon; C.dt = prm.time_step; C.vmax = prm.v_max; C.wmax = prm.w_max;
    C.inv_vmax = 1.0 / prm.v_max; C.inv_2wmax = 1.0 / (2.0 * prm.w_max); C.pose_jac_exact = prm.pose_jac_exact;
    const int K = C.K;
    const double inc = (prm.v_max * 0.99) * prm.time_step / prm.node_separation;   // mpc.py:29-30
    MpcProblem<KC> init, retry_init;
    init.T[0] = se2_from_pose_inv(px, py, pth);
    retry_init.T[0] = init.T[0];
    double target = interp_idx + inc;
    const bool first = (*it0 != 0);
    // Reference quirk (mpc.py:72-75,188-190): iteration0 is cleared before the solve, so the pose_error_scaling retry of a
    // failed FIRST solve starts from the stored warm start of the last successful solve (from before the reset()). Nothing
    // stored yet (warm_T still all zeros; the reference raises there): the iteration0 guess is used again.
    const bool has_stored = warm_T[0].x * warm_T[0].x + warm_T[0].y * warm_T[0].y > 0.5;
DRE_UNROLL_K
    for (int k = 0; k < KC; ++k) {
        if (k >= K) break;
        double r[3];
        mpc_ref_pose(path, stride, L, target, r);
        target += inc;
        init.M[k] = se2_from_pose_inv(r[0], r[1], r[2]);
        retry_init.M[k] = init.M[k];
        const double4 w = warm_T[k * ws];
        const double2 uu = warm_u[k * ws];
        retry_init.T[k + 1].c = w.x; retry_init.T[k + 1].s = w.y; retry_init.T[k + 1].x = w.z; retry_init.T[k + 1].y = w.w;
        retry_init.u[k][0] = uu.x; retry_init.u[k][1] = uu.y;
        if (first) {
            init.T[k + 1] = init.M[k]; init.u[k][0] = prm.v_max / 2.0; init.u[k][1] = 0.0;   // mpc.py:68-71
            if (!has_stored) { retry_init.T[k + 1] = init.T[k + 1]; retry_init.u[k][0] = init.u[k][0]; retry_init.u[k][1] = init.u[k][1]; }
        } else {
            init.T[k + 1] = retry_init.T[k + 1];
            init.u[k][0] = uu.x; init.u[k][1] = uu.y;
        }
    }
    *it0 = 0;
    MpcProblem<KC> P;
    MpcResult R;
    R.iterations = 0; R.lm_solves = 0; R.cost_evals = 0; R.term = 0; R.final_cost = 0.0;
    double s = 1.0;
    int retries = 0, term = 0;
    for (;;) {
        P = retries == 0 ? init : retry_init;
        term = mpc_optimize(C, prm, P, s, S, R);
        if (term) break;
        if (retries >= prm.max_retries) { term = -1; break; }
        s *= 5.0; ++retries;
    }
    for (int k = 0; k < K; ++k) {
        if (2 * k < ncontrols) controls[2 * k] = P.u[k][0];
        if (2 * k + 1 < ncontrols) controls[2 * k + 1] = P.u[k][1];
    }
    // warm start (mpc.py:192-205): poses [T_1..T_{K-1}, T_{K-1}] (not shifted), controls [u_1..u_{K-1}, (0.5, 0)]
    for (int i = 1; i < K; ++i) {
        warm_T[(i - 1) * ws] = make_double4(P.T[i].c, P.T[i].s, P.T[i].x, P.T[i].y);
        warm_u[(i - 1) * ws] = make_double2(P.u[i][0], P.u[i][1]);
    }
    {
        const Se2 TL = P.T[K - 1 >= 1 ? K - 1 : 0];
        warm_T[(K - 1) * ws] = make_double4(TL.c, TL.s, TL.x, TL.y);
        warm_u[(K - 1) * ws] = make_double2(0.5, 0.0);
    }
    if (status_out) {
        dre_mpc_status st;
        st.iterations = R.iterations; st.retries = retries; st.lm_solves = R.lm_solves; st.cost_evals = R.cost_evals;
        st.term = term; st.zero_steps = 0; st.final_cost = R.final_cost;
        *status_out = st;
    }
}
