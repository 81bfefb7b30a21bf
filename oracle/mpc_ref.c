/*
 * oracle/mpc_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat planar (SE(2)) double-precision CPU restatement of DR-MPC's path-tracking MPC:
 *     environment/path_tracking/utils/mpc.py:58-237           (problem set-up, warm start, retry, references)
 *     pysteam_augmented/lev_marq_gauss_newton_custom_solver.py:30-171 (LM inner loop, feasibility scaling,
 *                                                              2-trial backtracking, damping schedule)
 *     pysteam_augmented/extra_se3_evaluators.py:20-44          (clamped log barrier)
 *     pysteam_augmented/extra_loss_func.py:5-9                 (pose loss: cost 1/2 r^2, GN weight 3)
 * plus the outer loop / termination order of pysteam's Solver and the evaluator Jacobians of
 * pysteam/pylgmath (third-party, un-vendored, un-pinned; reference README.md:39-44, environment.yml:41),
 * restated from their published algorithm (STEAM, Barfoot ch. 7) -- see SURVEY.md Appendix B.
 *
 * The reference solves an 8K-dim SE(3) problem with a dense solve; for planar poses the in-plane and
 * out-of-plane coordinates decouple exactly, so this file solves the identical 5K-dim SE(2) problem with
 * a block-tridiagonal Cholesky over blocks z_j = (u_{j-1}, delta_j). It is pinned against the reference's own
 * unmodified Python run on the shims in oracle/shims (tests/golden/mpc_*.npz, tests/test_oracle_mpc.py).
 * PARITY UNPINNED w.r.t. the real pysteam/pylgmath (absent from this container).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MPC_MAXK 64

typedef struct {
    int32_t K;
    double dt, vmax, wmax;
    double node_inc;          /* mpc.py:29-30: 0.99*vmax*dt/node_separation */
    int32_t max_iterations;   /* mpc.py:144: 1500 */
    double abs_change_tol;    /* pysteam Solver default 1e-4 */
    double rel_change_tol;    /* mpc.py:144: 1e-4 */
    int32_t pose_jac_exact;   /* 0: -J_l^-1(e) (STEAM), 1: -J_r^-1(e) */
    int32_t max_retries;      /* cap on the pose_error_scaling*=5 loop (reference: unbounded) */
} mpc_params;

typedef struct {
    int32_t iterations;       /* outer LM iterations of the successful solve                    */
    int32_t retries;          /* number of pose_error_scaling *= 5 retries (mpc.py:188-190)      */
    int32_t lm_solves;        /* total damped linear solves (all attempts)                        */
    int32_t cost_evals;       /* total trial cost evaluations                                     */
    int32_t term;             /* 1 zero-grad, 3 max-iter, 4 abs cost, 5 abs change, 6 rel change, -1 gave up */
    int32_t pad;
    double final_cost;
} mpc_stats;

/* ---------- SE(2) helpers; element = (c, s, x, y) meaning [[c,-s,x],[s,c,y],[0,0,1]] ---------- */
typedef struct { double c, s, x, y; } se2;

static inline se2 se2_mul(se2 a, se2 b)
{
    se2 r;
    r.c = a.c * b.c - a.s * b.s;
    r.s = a.s * b.c + a.c * b.s;
    r.x = a.c * b.x - a.s * b.y + a.x;
    r.y = a.s * b.x + a.c * b.y + a.y;
    return r;
}
static inline se2 se2_inv(se2 a)
{
    se2 r;
    r.c = a.c; r.s = -a.s;
    r.x = -(a.c * a.x + a.s * a.y);
    r.y = -(-a.s * a.x + a.c * a.y);
    return r;
}
/* sin(p)/p, (1-cos p)/p, (1-cos p)/p^2, (p - sin p)/p^2 with series for small |p| */
static inline void phi_coeffs(double p, double *A, double *B, double *beta, double *alpha)
{
    if (fabs(p) < 1e-4) {
        const double p2 = p * p;
        *A = 1.0 - p2 / 6.0 + p2 * p2 / 120.0;
        *beta = 0.5 - p2 / 24.0 + p2 * p2 / 720.0;
        *B = p * *beta;
        *alpha = p * (1.0 / 6.0 - p2 / 120.0 + p2 * p2 / 5040.0);
    } else {
        const double sn = sin(p), cs = cos(p);
        *A = sn / p;
        *B = (1.0 - cs) / p;
        *beta = *B / p;
        *alpha = (p - sn) / (p * p);
    }
}
static inline se2 se2_exp(const double d[3])
{
    double A, B, beta, alpha;
    phi_coeffs(d[2], &A, &B, &beta, &alpha);
    se2 r;
    r.c = cos(d[2]); r.s = sin(d[2]);
    r.x = A * d[0] - B * d[1];
    r.y = B * d[0] + A * d[1];
    return r;
}
/* (phi/2) cot(phi/2) */
static inline double half_cot(double p)
{
    if (fabs(p) < 1e-4) { const double p2 = p * p; return 1.0 - p2 / 12.0 - p2 * p2 / 720.0; }
    const double h = 0.5 * p;
    return h / tan(h);
}
static inline void se2_log(se2 T, double e[3])
{
    const double p = atan2(T.s, T.c);
    const double ct = half_cot(p), h = 0.5 * p;
    e[0] = ct * T.x + h * T.y;
    e[1] = -h * T.x + ct * T.y;
    e[2] = p;
}
/* planar block of the SE(3) adjoint: [[c,-s,y],[s,c,-x],[0,0,1]] */
static inline void se2_Ad(se2 T, double M[9])
{
    M[0] = T.c; M[1] = -T.s; M[2] = T.y;
    M[3] = T.s; M[4] = T.c;  M[5] = -T.x;
    M[6] = 0.0; M[7] = 0.0;  M[8] = 1.0;
}
/* planar block of the SE(3) left Jacobian J_l(xi), xi = (rx, ry, p) */
static inline void se2_Jl(const double xi[3], double M[9])
{
    double A, B, beta, alpha;
    phi_coeffs(xi[2], &A, &B, &beta, &alpha);
    M[0] = A; M[1] = -B; M[2] = alpha * xi[0] + beta * xi[1];
    M[3] = B; M[4] = A;  M[5] = alpha * xi[1] - beta * xi[0];
    M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
}
/* planar block of J_l(xi)^-1 = [[Vinv, -Vinv q],[0,1]] */
static inline void se2_Jlinv(const double xi[3], double M[9])
{
    double A, B, beta, alpha;
    phi_coeffs(xi[2], &A, &B, &beta, &alpha);
    const double ct = half_cot(xi[2]), h = 0.5 * xi[2];
    const double qx = alpha * xi[0] + beta * xi[1], qy = alpha * xi[1] - beta * xi[0];
    M[0] = ct; M[1] = h;  M[2] = -(ct * qx + h * qy);
    M[3] = -h; M[4] = ct; M[5] = -(-h * qx + ct * qy);
    M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
}
static inline void m3_mul(const double a[9], const double b[9], double c[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

/* clamped log barrier, extra_se3_evaluators.py:20-44 */
static inline double barrier_val(double x) { return x <= 0.0 ? -500.0 : (x > 1.0 ? 500.0 : log(x)); }
static inline double barrier_grad(double x) { return (x <= 0.0 || x > 1.0) ? 10.0 : 1.0 / x; }

/* ---------- problem instance ---------- */
typedef struct {
    int K;
    double dt, vmax, wmax, s;      /* s = pose_error_scaling */
    int pose_jac_exact;
    se2 T[MPC_MAXK + 1];           /* T[0] locked */
    double u[MPC_MAXK][2];
    se2 M[MPC_MAXK];               /* inverse reference poses */
} mpc_prob;

/* cost = sum of 1/2 |whitened e|^2 in the reference's term order (mpc.py:99-142) */
static double mpc_cost(const mpc_prob *p)
{
    double cost = 0.0;
    for (int k = 0; k < p->K; ++k) {
        double e[3];
        se2_log(se2_mul(p->M[k], se2_inv(p->T[k + 1])), e);
        const double wi = 1.0 / sqrt(p->s);
        const double n0 = sqrt((wi * e[0]) * (wi * e[0]) + (wi * e[1]) * (wi * e[1]) + (wi * e[2]) * (wi * e[2]));
        cost += 0.5 * n0 * n0;
        const double xi[3] = {-p->dt * p->u[k][0], 0.0, -p->dt * p->u[k][1]};
        se2 E = se2_mul(se2_mul(p->T[k + 1], se2_inv(p->T[k])), se2_inv(se2_exp(xi)));
        se2_log(E, e);
        const double wk = 1.0 / sqrt(0.001);
        const double n1 = sqrt((wk * e[0]) * (wk * e[0]) + (wk * e[1]) * (wk * e[1]) + (wk * e[2]) * (wk * e[2]));
        cost += 0.5 * n1 * n1;
        const double v = p->u[k][0], w = p->u[k][1];
        const double xs[4] = {(p->vmax - v) * (1.0 / p->vmax), (v - 0.0) * (1.0 / p->vmax),
                              (p->wmax - w) * (1.0 / (2.0 * p->wmax)), (w + p->wmax) * (1.0 / (2.0 * p->wmax))};
        const double wb = 1.0 / sqrt(10.0);
        for (int t = 0; t < 4; ++t) {
            const double eb = fabs(wb * (barrier_val(xs[t]) * (1.0 / 500.0)));
            cost += 0.5 * eb * eb;
        }
    }
    return cost;
}

/* Block-tridiagonal normal equations over z_j=(u_{j-1}(2), delta_j(3)), j=1..K.
 * D[j] 5x5 (full, symmetric), O[j] 5x5 coupling rows z_j with cols z_{j-1} (only cols 2..4 non-zero). */
typedef struct {
    double D[MPC_MAXK][25];
    double O[MPC_MAXK][25];
    double b[MPC_MAXK][5];
} mpc_sys;

static void mpc_build(const mpc_prob *p, mpc_sys *S)
{
    const int K = p->K;
    memset(S->D, 0, sizeof(double) * 25 * (size_t)K);
    memset(S->O, 0, sizeof(double) * 25 * (size_t)K);
    memset(S->b, 0, sizeof(double) * 5 * (size_t)K);
    for (int k = 0; k < K; ++k) {
        const int j = k; /* block index of z_{k+1} */
        /* ---- pose term on delta_{k+1} ---- */
        {
            double e[3], G[9], Jp[9];
            const se2 E = se2_mul(p->M[k], se2_inv(p->T[k + 1]));
            se2_log(E, e);
            se2_Jlinv(e, G);
            if (p->pose_jac_exact) { double Ad[9], t[9]; se2_Ad(E, Ad); m3_mul(G, Ad, t); memcpy(G, t, sizeof t); }
            for (int i = 0; i < 9; ++i) Jp[i] = -G[i];
            const double wgt = 3.0 / p->s;
            for (int a = 0; a < 3; ++a) {
                double g = 0.0;
                for (int r = 0; r < 3; ++r) g += Jp[3 * r + a] * e[r];
                S->b[j][2 + a] += -wgt * g;
                for (int c = 0; c < 3; ++c) {
                    double h = 0.0;
                    for (int r = 0; r < 3; ++r) h += Jp[3 * r + a] * Jp[3 * r + c];
                    S->D[j][5 * (2 + a) + (2 + c)] += wgt * h;
                }
            }
        }
        /* ---- kinematic term on delta_{k+1}, delta_k, u_k ---- */
        {
            double e[3], G[9], Ad[9], Jl[9], t1[9], t2[9];
            const double xi[3] = {-p->dt * p->u[k][0], 0.0, -p->dt * p->u[k][1]};
            const se2 Lh = se2_mul(p->T[k + 1], se2_inv(p->T[k]));
            const se2 E = se2_mul(Lh, se2_inv(se2_exp(xi)));
            se2_log(E, e);
            se2_Jlinv(e, G);
            /* J5 = [J_u (3x2) | J_{k+1} (3x3)] */
            double J5[15], Jk[9];
            se2_Ad(E, Ad); se2_Jl(xi, Jl);
            m3_mul(G, Ad, t1); m3_mul(t1, Jl, t2);
            for (int r = 0; r < 3; ++r) {
                J5[5 * r + 0] = p->dt * t2[3 * r + 0];
                J5[5 * r + 1] = p->dt * t2[3 * r + 2];
                J5[5 * r + 2] = G[3 * r + 0]; J5[5 * r + 3] = G[3 * r + 1]; J5[5 * r + 4] = G[3 * r + 2];
            }
            const double wgt = 1.0 / 0.001;
            for (int a = 0; a < 5; ++a) {
                double g = 0.0;
                for (int r = 0; r < 3; ++r) g += J5[5 * r + a] * e[r];
                S->b[j][a] += -wgt * g;
                for (int c = 0; c < 5; ++c) {
                    double h = 0.0;
                    for (int r = 0; r < 3; ++r) h += J5[5 * r + a] * J5[5 * r + c];
                    S->D[j][5 * a + c] += wgt * h;
                }
            }
            if (k >= 1) {
                se2_Ad(Lh, Ad); m3_mul(G, Ad, t1);
                for (int i = 0; i < 9; ++i) Jk[i] = -t1[i];
                for (int a = 0; a < 3; ++a) {
                    double g = 0.0;
                    for (int r = 0; r < 3; ++r) g += Jk[3 * r + a] * e[r];
                    S->b[j - 1][2 + a] += -wgt * g;
                    for (int c = 0; c < 3; ++c) {
                        double h = 0.0;
                        for (int r = 0; r < 3; ++r) h += Jk[3 * r + a] * Jk[3 * r + c];
                        S->D[j - 1][5 * (2 + a) + (2 + c)] += wgt * h;
                    }
                }
                for (int a = 0; a < 5; ++a)
                    for (int c = 0; c < 3; ++c) {
                        double h = 0.0;
                        for (int r = 0; r < 3; ++r) h += J5[5 * r + a] * Jk[3 * r + c];
                        S->O[j][5 * a + (2 + c)] += wgt * h;
                    }
            }
        }
        /* ---- four log barriers on u_k ---- */
        {
            const double v = p->u[k][0], w = p->u[k][1];
            const double iv = 1.0 / p->vmax, iw = 1.0 / (2.0 * p->wmax);
            const double xs[4] = {(p->vmax - v) * iv, (v - 0.0) * iv, (p->wmax - w) * iw, (w + p->wmax) * iw};
            const double dx[4] = {-iv, iv, -iw, iw};
            const double wgt = 1.0 / 10.0;
            for (int t = 0; t < 4; ++t) {
                const double e = barrier_val(xs[t]) * (1.0 / 500.0);
                const double J = (1.0 / 500.0) * barrier_grad(xs[t]) * dx[t];
                const int a = t / 2;
                S->b[j][a] += -wgt * J * e;
                S->D[j][5 * a + a] += wgt * J * J;
            }
        }
    }
}

/* Damped block-tridiagonal Cholesky solve: (A with diag*(1+lambda)) x = b. Returns 0 on a non-positive pivot. */
static int mpc_solve_lm(const mpc_sys *S, int K, double lambda, double x[][5])
{
    static _Thread_local double Lc[MPC_MAXK][25], Bc[MPC_MAXK][25];
    double y[MPC_MAXK][5];
    for (int j = 0; j < K; ++j) {
        double Sj[25];
        memcpy(Sj, S->D[j], sizeof Sj);
        for (int a = 0; a < 5; ++a) Sj[5 * a + a] = S->D[j][5 * a + a] * (1.0 + lambda);
        double rhs[5];
        memcpy(rhs, S->b[j], sizeof rhs);
        if (j > 0) {
            /* B = O_j L_{j-1}^{-T}: solve B L^T = O row by row (forward substitution on columns) */
            const double *Lp = Lc[j - 1];
            for (int a = 0; a < 5; ++a) {
                for (int c = 0; c < 5; ++c) {
                    double v = S->O[j][5 * a + c];
                    for (int m = 0; m < c; ++m) v -= Bc[j][5 * a + m] * Lp[5 * c + m];
                    Bc[j][5 * a + c] = v / Lp[5 * c + c];
                }
            }
            for (int a = 0; a < 5; ++a) {
                for (int c = 0; c < 5; ++c) {
                    double v = 0.0;
                    for (int m = 0; m < 5; ++m) v += Bc[j][5 * a + m] * Bc[j][5 * c + m];
                    Sj[5 * a + c] -= v;
                }
                double v = 0.0;
                for (int m = 0; m < 5; ++m) v += Bc[j][5 * a + m] * y[j - 1][m];
                rhs[a] -= v;
            }
        }
        /* dense 5x5 Cholesky, lower */
        double *L = Lc[j];
        memset(L, 0, sizeof(double) * 25);
        for (int c = 0; c < 5; ++c) {
            double d = Sj[5 * c + c];
            for (int m = 0; m < c; ++m) d -= L[5 * c + m] * L[5 * c + m];
            if (!(d > 0.0)) return 0;
            const double ld = sqrt(d);
            L[5 * c + c] = ld;
            for (int a = c + 1; a < 5; ++a) {
                double v = Sj[5 * a + c];
                for (int m = 0; m < c; ++m) v -= L[5 * a + m] * L[5 * c + m];
                L[5 * a + c] = v / ld;
            }
        }
        /* y_j = L^-1 rhs */
        for (int a = 0; a < 5; ++a) {
            double v = rhs[a];
            for (int m = 0; m < a; ++m) v -= L[5 * a + m] * y[j][m];
            y[j][a] = v / L[5 * a + a];
        }
    }
    /* back substitution: L_j^T x_j = y_j - B_{j+1}^T x_{j+1} */
    for (int j = K - 1; j >= 0; --j) {
        double rhs[5];
        memcpy(rhs, y[j], sizeof rhs);
        if (j + 1 < K)
            for (int a = 0; a < 5; ++a) {
                double v = 0.0;
                for (int m = 0; m < 5; ++m) v += Bc[j + 1][5 * m + a] * x[j + 1][m];
                rhs[a] -= v;
            }
        const double *L = Lc[j];
        for (int a = 4; a >= 0; --a) {
            double v = rhs[a];
            for (int m = a + 1; m < 5; ++m) v -= L[5 * m + a] * x[j][m];
            x[j][a] = v / L[5 * a + a];
        }
    }
    return 1;
}

/* b^T p - 1/2 p^T A p with the undamped A (lev_marq...py:167-171) */
static double mpc_predict(const mpc_sys *S, int K, double p[][5])
{
    double g = 0.0, q = 0.0;
    for (int j = 0; j < K; ++j) {
        for (int a = 0; a < 5; ++a) {
            g += S->b[j][a] * p[j][a];
            double r = 0.0;
            for (int c = 0; c < 5; ++c) r += S->D[j][5 * a + c] * p[j][c];
            q += p[j][a] * r;
            if (j > 0) {
                double o = 0.0;
                for (int c = 0; c < 5; ++c) o += S->O[j][5 * a + c] * p[j - 1][c];
                q += 2.0 * p[j][a] * o;
            }
        }
    }
    return g - 0.5 * q;
}

static void mpc_apply(mpc_prob *p, double d[][5])
{
    for (int k = 0; k < p->K; ++k) {
        p->u[k][0] += d[k][0]; p->u[k][1] += d[k][1];
        p->T[k + 1] = se2_mul(se2_exp(&d[k][2]), p->T[k + 1]);
    }
}

/* One optimize() call at fixed pose_error_scaling. Returns term code (>0) or 0 when the step failed. */
static int mpc_optimize(mpc_prob *p, const mpc_params *P, mpc_stats *st)
{
    static _Thread_local mpc_sys S;
    const int K = p->K;
    double lambda = 1e-7;                      /* lev_marq...py:28 */
    double curr = mpc_cost(p), prev;
    int iter = 0;
    for (;;) {
        ++iter;
        prev = curr;
        mpc_build(p, &S);
        double gn = 0.0;
        for (int j = 0; j < K; ++j) for (int a = 0; a < 5; ++a) gn += S.b[j][a] * S.b[j][a];
        gn = sqrt(gn);
        int success = 0;
        double new_cost = prev;
        for (int nb = 0; nb < 50 && !success; ++nb) {      /* max_shrink_steps */
            double d[MPC_MAXK][5];
            st->lm_solves++;
            if (mpc_solve_lm(&S, K, lambda, d)) {
                /* feasibility scaling, lev_marq...py:52-91 (walk u_0..u_{K-1}, v then w) */
                double scale = 1.0;
                for (int k = 0; k < K; ++k) {
                    const double lim_hi[2] = {P->vmax, P->wmax}, lim_lo[2] = {0.0, -P->wmax};
                    for (int c = 0; c < 2; ++c) {
                        const double dp = d[k][c];
                        if (dp != 0.0) {
                            const double maxp = dp > 0.0 ? (lim_hi[c] - 1e-7) - p->u[k][c] : (lim_lo[c] + 1e-7) - p->u[k][c];
                            const double allowed = maxp / dp;
                            scale = fmin(scale, allowed);
                            if (allowed < 1.0) scale = fmax(scale - 1e-7, 0.0);
                        }
                    }
                }
                for (int j = 0; j < K; ++j) for (int a = 0; a < 5; ++a) d[j][a] *= scale;
                for (int bt = 0; bt < 2; ++bt) {            /* max_backtrack_steps */
                    if (bt > 0) for (int j = 0; j < K; ++j) for (int a = 0; a < 5; ++a) d[j][a] *= 0.8;
                    mpc_prob trial = *p;
                    mpc_apply(&trial, d);
                    const double proposed = mpc_cost(&trial);
                    st->cost_evals++;
                    const double actual = prev - proposed;
                    const double pred = mpc_predict(&S, K, d);
                    const double ratio = (pred == 0.0) ? 0.0 : actual / pred;
#ifdef MPC_TRACE
                    fprintf(stderr, "iter %d nb %d bt %d lambda %.1e scale %.3e prev %.9g proposed %.9g actual %.3e pred %.3e ratio %.4f\n",
                            iter, nb, bt, lambda, scale, prev, proposed, actual, pred, ratio);
#endif
                    if (ratio > 0.25) {
                        *p = trial;
                        lambda = fmax(lambda * 0.1, 1e-7);
                        new_cost = proposed;
                        success = 1;
                        break;
                    }
                }
                if (!success) lambda = fmin(lambda * 10.0, 1e7);
            } else {
                lambda = fmin(lambda * 10.0, 1e7);
            }
        }
        curr = new_cost;
        st->iterations = iter;
        st->final_cost = curr;
        if (!success && fabs(gn) < 1e-6) return 1;
        if (!success) return 0;
        if (iter >= P->max_iterations) return 3;
        if (curr <= 0.0) return 4;
        if (fabs(prev - curr) <= P->abs_change_tol) return 5;
        if (fabs(prev - curr) / prev <= P->rel_change_tol) return 6;
    }
}

static se2 pose_to_inv_se2(double x, double y, double th)
{
    se2 T = {cos(th), sin(th), x, y};
    return se2_inv(T);
}

/* mpc.py:208-237: K references at interp_idx + (k+1)*inc, lerp incl. theta, clamped to the last node.
 * path is [3][Lstride] row-major (x row, y row, theta row), L valid nodes. */
void mpc_ref_local_path(const double *path, int Lstride, int L, double interp_idx, double inc, int K, double *ref /*[K][3]*/)
{
    double target = interp_idx + inc;
    for (int i = 0; i < K; ++i) {
        if (target >= (double)(L - 1)) {
            for (int r = 0; r < 3; ++r) ref[3 * i + r] = path[(size_t)r * Lstride + (L - 1)];
        } else {
            const int lo = (int)floor(target), hi = (int)ceil(target);
            const double f = target - (double)lo;
            for (int r = 0; r < 3; ++r) {
                const double a = path[(size_t)r * Lstride + lo], b = path[(size_t)r * Lstride + hi];
                ref[3 * i + r] = a + f * (b - a);
            }
        }
        target += inc;
    }
}

/*
 * MPC.act for one problem. Persistent state (mpc.py:51-56,192-205): warm_T [K][4] (c,s,x,y of the
 * inverse poses T_1..T_K guess), warm_u [K][2], *iteration0.
 */
void mpc_ref_act(const mpc_params *P, const double *path, int Lstride, int L, double interp_idx, const double pose[3],
                 double *warm_T, double *warm_u, int32_t *iteration0, double *out_u, mpc_stats *st_out)
{
    const int K = P->K;
    mpc_stats st; memset(&st, 0, sizeof st);
    double ref[MPC_MAXK * 3];
    mpc_ref_local_path(path, Lstride, L, interp_idx, P->node_inc, K, ref);
    mpc_prob init; memset(&init, 0, sizeof init);
    init.K = K; init.dt = P->dt; init.vmax = P->vmax; init.wmax = P->wmax; init.pose_jac_exact = P->pose_jac_exact;
    for (int k = 0; k < K; ++k) init.M[k] = pose_to_inv_se2(ref[3 * k], ref[3 * k + 1], ref[3 * k + 2]);
    init.T[0] = pose_to_inv_se2(pose[0], pose[1], pose[2]);
    /* what `deepcopy(self.next_*_state_vars)` (mpc.py:73-75) would yield: the warm start stored by the last successful
     * solve. next_* are the ints 0 until the first solve has finished (mpc.py:52-53): warm_T is all zeros then. */
    mpc_prob stored = init;
    const int has_stored = warm_T[0] * warm_T[0] + warm_T[1] * warm_T[1] > 0.5;
    for (int k = 0; k < K; ++k) {
        se2 T = {warm_T[4 * k], warm_T[4 * k + 1], warm_T[4 * k + 2], warm_T[4 * k + 3]};
        stored.T[k + 1] = T; stored.u[k][0] = warm_u[2 * k]; stored.u[k][1] = warm_u[2 * k + 1];
    }
    mpc_prob retry_init;
    if (*iteration0) {
        for (int k = 0; k < K; ++k) { init.T[k + 1] = init.M[k]; init.u[k][0] = P->vmax / 2.0; init.u[k][1] = 0.0; }
        *iteration0 = 0;
        /* Reference quirk (mpc.py:72): iteration0 is cleared BEFORE the solve, so the pose_error_scaling retry of a failed
         * first solve (mpc.py:188-190 -> :73-75) starts from the STORED warm start -- the one from before the reset().
         * If nothing was ever stored the reference dies there (`deepcopy(0).insert` raises outside the try); the batch
         * restarts from the iteration0 guess instead. */
        retry_init = has_stored ? stored : init;
    } else {
        init = stored;
        retry_init = stored;
    }
    mpc_prob p;
    double s = 1.0;
    int term = 0;
    for (;;) {
        p = st.retries == 0 ? init : retry_init; p.s = s;
        term = mpc_optimize(&p, P, &st);
        if (term) break;
        if (st.retries >= P->max_retries) { term = -1; break; }
        s *= 5.0; st.retries++;
    }
    st.term = term;
    for (int k = 0; k < K; ++k) { out_u[2 * k] = p.u[k][0]; out_u[2 * k + 1] = p.u[k][1]; }
    /* warm start for the next call, mpc.py:192-205: poses [T_1..T_{K-1}, T_{K-1}], controls [u_1..u_{K-1},(0.5,0)] */
    for (int i = 1; i < K; ++i) {
        warm_T[4 * (i - 1)] = p.T[i].c; warm_T[4 * (i - 1) + 1] = p.T[i].s;
        warm_T[4 * (i - 1) + 2] = p.T[i].x; warm_T[4 * (i - 1) + 3] = p.T[i].y;
        warm_u[2 * (i - 1)] = p.u[i][0]; warm_u[2 * (i - 1) + 1] = p.u[i][1];
    }
    {
        const int last = (K - 1 >= 1) ? K - 1 : 0; /* K==1: se3_state_vars[0] is the locked current pose */
        const se2 TL = p.T[last];
        warm_T[4 * (K - 1)] = TL.c; warm_T[4 * (K - 1) + 1] = TL.s; warm_T[4 * (K - 1) + 2] = TL.x; warm_T[4 * (K - 1) + 3] = TL.y;
        warm_u[2 * (K - 1)] = 0.5; warm_u[2 * (K - 1) + 1] = 0.0;
    }
    if (st_out) *st_out = st;
}

/* Batched act over N independent problems; paths [P][3][Lstride], lens[P], path_id[N]. */
void mpc_ref_act_batch(const mpc_params *P, int N, const double *paths, int Lstride, const int32_t *lens,
                       const int32_t *path_id, const double *interp_idx, const double *pose /*[N][3]*/,
                       double *warm_T /*[N][K][4]*/, double *warm_u /*[N][K][2]*/, int32_t *iteration0 /*[N]*/,
                       double *out_u /*[N][2K]*/, mpc_stats *stats /*[N] or NULL*/, int nthreads)
{
    const int K = P->K;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int n = 0; n < N; ++n) {
        const int pid = path_id ? path_id[n] : 0;
        mpc_ref_act(P, paths + (size_t)pid * 3 * Lstride, Lstride, lens[pid], interp_idx[n], pose + 3 * (size_t)n,
                    warm_T + (size_t)n * K * 4, warm_u + (size_t)n * K * 2, iteration0 + n,
                    out_u + (size_t)n * 2 * K, stats ? stats + n : NULL);
    }
}

/* exposed for tests: cost / normal equations of an explicit state */
double mpc_ref_cost_of(int K, double dt, double vmax, double wmax, double s, const double *T /*[K+1][4]*/,
                       const double *u /*[K][2]*/, const double *M /*[K][4]*/)
{
    mpc_prob p; memset(&p, 0, sizeof p);
    p.K = K; p.dt = dt; p.vmax = vmax; p.wmax = wmax; p.s = s;
    for (int k = 0; k <= K; ++k) { se2 t = {T[4 * k], T[4 * k + 1], T[4 * k + 2], T[4 * k + 3]}; p.T[k] = t; }
    for (int k = 0; k < K; ++k) {
        se2 m = {M[4 * k], M[4 * k + 1], M[4 * k + 2], M[4 * k + 3]}; p.M[k] = m;
        p.u[k][0] = u[2 * k]; p.u[k][1] = u[2 * k + 1];
    }
    return mpc_cost(&p);
}
