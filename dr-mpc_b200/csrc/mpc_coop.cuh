// mpc_coop.cuh -- lane-per-horizon-step cooperative LM solve (the fast path of the MPC kernel).
//
// A group of G lanes (G = 4, 8, 16 or 32 >= K) owns one problem; lane k owns horizon step k: pose T_{k+1}, control
// u_k, reference M_k and block row k of the block-tridiagonal normal equations (diagonal block D_k packed 5x5,
// coupling block O_k 5x3, rhs b_k), all in REGISTERS -- nothing lives in local memory, which is what bounds the
// one-thread-per-problem formulation (its 3.5 KB/thread working set falls out of L1). Residuals, Jacobians and
// trial costs are embarrassingly parallel over lanes (the neighbour pose comes by shuffle); the block Cholesky
// runs as K pipeline stages in which lane j consumes lane j-1's factor rows (9 doubles by shuffle); sums over
// the horizon are taken in step order so they round like the sequential reference.
//
// Control flow is a per-group state machine with the phases in a fixed order (initial cost -> build -> damped solve ->
// trial 1 -> trial 2 -> bookkeeping), so the groups of a warp, which follow different LM paths, re-converge at every
// loop iteration instead of serialising whole solves.
#pragma once
#include "mpc_device.cuh"

#if defined(__CUDACC__)

template <int G>
struct LaneGroup {
    unsigned mask;
    int lane;   // rank within the group
    __device__ LaneGroup()
    {
        const unsigned wl = threadIdx.x & 31u;
        lane = (int)(wl & (G - 1));
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(unsigned)(G - 1)));
    }
    __device__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ double up(double v) const { return __shfl_up_sync(mask, v, 1, G); }
    __device__ double down(double v) const { return __shfl_down_sync(mask, v, 1, G); }
    __device__ bool any(bool p) const { return (__ballot_sync(mask, p) & mask) != 0u; }
    // sum over lanes 0..K-1 in lane order (every lane gets the same bits)
    __device__ double sum_ordered(double v, int K) const
    {
        if (G <= 8) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < G; ++k) {
                const double t = __shfl_sync(mask, v, k, G);
                if (k < K) s += t;
            }
            return s;
        } else {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
            return v;
        }
    }
    __device__ Se2 up(const Se2 &T) const
    {
        Se2 r;
        r.c = up(T.c); r.s = up(T.s); r.x = up(T.x); r.y = up(T.y);
        return r;
    }
};

// per-lane cost of step k: pose + kinematic + 4 barriers (mpc.py:99-142)
__device__ __forceinline__ double coop_lane_cost(const MpcConst &C, const Se2 &Tprev, const Se2 &Tk1, const double u[2],
                                                 const Se2 &Mk, double inv_s)
{
    double e[3], ct;
    se2_log(se2_mul_inv(Mk, Tk1), e, ct);
    double cost = 0.5 * inv_s * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double p = C.dt * u[1];
    double sn, cs;
    sincos(p, &sn, &cs);
    const PhiCoef q = phi_coef(p, sn, cs);
    const Se2 E = se2_mul(se2_mul_inv(Tk1, Tprev), se2_exp_x(C.dt * u[0], p, sn, cs, q));
    se2_log(E, e, ct);
    cost += 0.5 * 1000.0 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double b0 = barrier_val((C.vmax - u[0]) * C.inv_vmax), b1 = barrier_val(u[0] * C.inv_vmax);
    const double b2 = barrier_val((C.wmax - u[1]) * C.inv_2wmax), b3 = barrier_val((u[1] + C.wmax) * C.inv_2wmax);
    cost += 0.5 * 0.1 * (1.0 / 250000.0) * (b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
    return cost;
}

struct CoopRow {
    double D[15];   // packed lower 5x5 diagonal block of z_k = (u_k, delta_{k+1})
    double O[15];   // 5x3 coupling with delta_k (previous lane)
    double b[5];
    double L[15];   // Cholesky factor of the Schur complement
    double B[15];   // O L_{prev}^{-T}
};

// Build block row k. `first` = lane 0 (T_k is the locked current pose: no coupling / no contribution to a previous block).
// pc6 / pcb: this step's kinematic contribution to the PREVIOUS block's delta-delta part (sent down by the caller).
__device__ __forceinline__ void coop_lane_build(const MpcConst &C, const Se2 &Tprev, const Se2 &Tk1, const double u[2],
                                                const Se2 &Mk, double s, bool first, CoopRow &R, double pc6[6], double pcb[3])
{
#pragma unroll
    for (int i = 0; i < 15; ++i) { R.D[i] = 0.0; R.O[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 5; ++i) R.b[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) pc6[i] = 0.0;
    pcb[0] = pcb[1] = pcb[2] = 0.0;
    {   // pose term
        double e[3], ct;
        const Se2 E = se2_mul_inv(Mk, Tk1);
        se2_log(E, e, ct);
        Aff Gm = se2_Jlinv(e, E.s, E.c, ct);
        if (C.pose_jac_exact) Gm = aff_mul(Gm, se2_Ad(E));
        const double r0[3] = {-Gm.m[0], -Gm.m[1], -Gm.m[2]}, r1[3] = {-Gm.m[3], -Gm.m[4], -Gm.m[5]}, r2[3] = {0.0, 0.0, -1.0};
        acc_block<3>(R.D, R.b, 2, r0, r1, r2, e, 3.0 / s);
    }
    {   // kinematic term
        const double p = C.dt * u[1];
        double sn, cs;
        sincos(p, &sn, &cs);
        const PhiCoef q = phi_coef(p, sn, cs);
        const double rx = C.dt * u[0];
        const Se2 Lh = se2_mul_inv(Tk1, Tprev);
        const Se2 E = se2_mul(Lh, se2_exp_x(rx, p, sn, cs, q));
        double e[3], ct;
        se2_log(E, e, ct);
        const Aff Gm = se2_Jlinv(e, E.s, E.c, ct);
        Aff Jl;
        Jl.m[0] = q.A; Jl.m[1] = q.B;  Jl.m[2] = q.alpha * rx;
        Jl.m[3] = -q.B; Jl.m[4] = q.A; Jl.m[5] = q.beta * rx;
        const Aff Ju = aff_mul(aff_mul(Gm, se2_Ad(E)), Jl);
        const double r0[5] = {C.dt * Ju.m[0], C.dt * Ju.m[2], Gm.m[0], Gm.m[1], Gm.m[2]};
        const double r1[5] = {C.dt * Ju.m[3], C.dt * Ju.m[5], Gm.m[3], Gm.m[4], Gm.m[5]};
        const double r2[5] = {0.0, C.dt, 0.0, 0.0, 1.0};
        acc_block<5>(R.D, R.b, 0, r0, r1, r2, e, 1000.0);
        if (!first) {
            const Aff GL = aff_mul(Gm, se2_Ad(Lh));
            const double q0[3] = {-GL.m[0], -GL.m[1], -GL.m[2]}, q1[3] = {-GL.m[3], -GL.m[4], -GL.m[5]}, q2[3] = {0.0, 0.0, -1.0};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pcb[a] = -1000.0 * (q0[a] * e[0] + q1[a] * e[1] + q2[a] * e[2]);
#pragma unroll
                for (int c = 0; c <= a; ++c) pc6[a * (a + 1) / 2 + c] = 1000.0 * (q0[a] * q0[c] + q1[a] * q1[c] + q2[a] * q2[c]);
            }
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) R.O[3 * a + c] = 1000.0 * (r0[a] * q0[c] + r1[a] * q1[c] + r2[a] * q2[c]);
        }
    }
    {   // barriers
        const double v = u[0], w = u[1];
        const double x0 = (C.vmax - v) * C.inv_vmax, x1 = v * C.inv_vmax;
        const double x2 = (C.wmax - w) * C.inv_2wmax, x3 = (w + C.wmax) * C.inv_2wmax;
        const double k500 = 1.0 / 500.0;
        const double e0 = barrier_val(x0) * k500, e1 = barrier_val(x1) * k500;
        const double e2 = barrier_val(x2) * k500, e3 = barrier_val(x3) * k500;
        const double j0 = -k500 * barrier_grad(x0) * C.inv_vmax, j1 = k500 * barrier_grad(x1) * C.inv_vmax;
        const double j2 = -k500 * barrier_grad(x2) * C.inv_2wmax, j3 = k500 * barrier_grad(x3) * C.inv_2wmax;
        R.b[0] -= 0.1 * (j0 * e0 + j1 * e1);
        R.b[1] -= 0.1 * (j2 * e2 + j3 * e3);
        R.D[0] += 0.1 * (j0 * j0 + j1 * j1);
        R.D[2] += 0.1 * (j2 * j2 + j3 * j3);
    }
}

// Damped block-tridiagonal Cholesky as a K-stage lane pipeline. Returns false (group-uniform) on a non-positive pivot.
// The diagonal slots of R.L hold the INVERSE pivots 1/l_cc (computed with one rsqrt per pivot): every later use of a
// pivot is a division, so the dependent chain of a stage is multiplications only.
template <int G>
__device__ __forceinline__ bool coop_solve(const LaneGroup<G> &g, int K, CoopRow &R, double lambda, double d[5])
{
    const double damp = 1.0 + lambda;
    double y[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 15; ++i) { R.L[i] = 0.0; R.B[i] = 0.0; }
    R.L[0] = R.L[2] = R.L[5] = R.L[9] = R.L[14] = 1.0;
    for (int j = 0; j < K; ++j) {
        // factor rows 2..4 and the forward-substituted rhs of the previous block
        const double i22 = g.up(R.L[5]), l32 = g.up(R.L[8]), i33 = g.up(R.L[9]);
        const double l42 = g.up(R.L[12]), l43 = g.up(R.L[13]), i44 = g.up(R.L[14]);
        const double y2 = g.up(y[2]), y3 = g.up(y[3]), y4 = g.up(y[4]);
        if (g.lane == j) {
            double S[15], rhs[5];
#pragma unroll
            for (int i = 0; i < 15; ++i) S[i] = R.D[i];
            S[0] *= damp; S[2] *= damp; S[5] *= damp; S[9] *= damp; S[14] *= damp;
#pragma unroll
            for (int a = 0; a < 5; ++a) rhs[a] = R.b[a];
            if (j > 0) {
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    const double b2 = R.O[3 * a] * i22;
                    const double b3 = (R.O[3 * a + 1] - b2 * l32) * i33;
                    const double b4 = (R.O[3 * a + 2] - b2 * l42 - b3 * l43) * i44;
                    R.B[3 * a] = b2; R.B[3 * a + 1] = b3; R.B[3 * a + 2] = b4;
                }
#pragma unroll
                for (int a = 0; a < 5; ++a) {
#pragma unroll
                    for (int c = 0; c <= a; ++c)
                        S[a * (a + 1) / 2 + c] -= R.B[3 * a] * R.B[3 * c] + R.B[3 * a + 1] * R.B[3 * c + 1] + R.B[3 * a + 2] * R.B[3 * c + 2];
                    rhs[a] -= R.B[3 * a] * y2 + R.B[3 * a + 1] * y3 + R.B[3 * a + 2] * y4;
                }
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                double dd = S[c * (c + 1) / 2 + c];
#pragma unroll
                for (int m = 0; m < c; ++m) dd -= R.L[c * (c + 1) / 2 + m] * R.L[c * (c + 1) / 2 + m];
                if (!(dd > 0.0)) { ok = false; dd = 1.0; }
                const double ild = rsqrt(dd);
                R.L[c * (c + 1) / 2 + c] = ild;
#pragma unroll
                for (int a = c + 1; a < 5; ++a) {
                    double v = S[a * (a + 1) / 2 + c];
#pragma unroll
                    for (int m = 0; m < c; ++m) v -= R.L[a * (a + 1) / 2 + m] * R.L[c * (c + 1) / 2 + m];
                    R.L[a * (a + 1) / 2 + c] = v * ild;
                }
            }
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                double v = rhs[a];
#pragma unroll
                for (int m = 0; m < a; ++m) v -= R.L[a * (a + 1) / 2 + m] * y[m];
                y[a] = v * R.L[a * (a + 1) / 2 + a];
            }
        }
    }
    ok = !g.any(!ok);
    // back substitution: L_j^T d_j = y_j - B_{j+1}^T d_{j+1}
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int j = K - 1; j >= 0; --j) {
        const double r0 = g.down(t0), r1 = g.down(t1), r2 = g.down(t2);
        if (g.lane == j) {
            double rhs[5] = {y[0], y[1], y[2], y[3], y[4]};
            if (j + 1 < K) { rhs[2] -= r0; rhs[3] -= r1; rhs[4] -= r2; }
#pragma unroll
            for (int a = 4; a >= 0; --a) {
                double v = rhs[a];
#pragma unroll
                for (int m = a + 1; m < 5; ++m) v -= R.L[m * (m + 1) / 2 + a] * d[m];
                d[a] = v * R.L[a * (a + 1) / 2 + a];
            }
            t0 = t1 = t2 = 0.0;
#pragma unroll
            for (int m = 0; m < 5; ++m) { t0 += R.B[3 * m] * d[m]; t1 += R.B[3 * m + 1] * d[m]; t2 += R.B[3 * m + 2] * d[m]; }
        }
    }
    return ok;
}

// Arguments of one batched act() call (device pointers).
struct CoopBatch {
    int N;
    const int32_t *path_id;
    const double *interp_index, *pose;
    int pose_stride;
    double *controls;
    int controls_stride, controls_count;
    dre_mpc_status *status;
    const uint8_t *active_mask;
    const double *paths;
    const int32_t *lens;
    int num_paths, path_stride;
    double4 *warm_T;      // [N][K]
    double2 *warm_u;      // [N][K]
    uint8_t *iteration0;  // [N]
    int *queue;           // work counter (zeroed before the launch)
    unsigned long long *trace;   // TRACE builds only: [0..5] cycles in fetch/cost0/build/solve/trials/bookkeeping,
                                 // [6] loop iterations, [7] sum of the warp's live groups over iterations, [8] max iterations of a group,
                                 // [9] groups, [10] cycles of the longest-running group (all summed over group leaders)
};

// PERSISTENT lane-group loop: a group of G lanes pulls problems from a global counter and runs one LM attempt
// (damped solve + up to two trials) per loop iteration. LM paths differ per problem (3 to 60 attempts), and the 32/G
// groups of a warp advance in lock-step per iteration; letting a finished group fetch the next problem immediately
// keeps every group busy instead of idling until the slowest problem of its warp is done.
// Semantics per problem: identical to mpc_act_one (mpc_device.cuh) = reference MPC.act (mpc.py:58-206).
template <int G, bool TRACE = false>
__device__ void mpc_coop_persistent(const LaneGroup<G> &g, const dre_mpc_params &prm, const CoopBatch &B)
{
    unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_t = 0, tr_start = 0;
    if (TRACE) { tr_t = clock64(); tr_start = tr_t; }
#define DRE_TR(i) do { if (TRACE) { const long long _n = clock64(); tr[i] += (unsigned long long)(_n - tr_t); tr_t = _n; } } while (0)
    MpcConst C;
    C.K = prm.horizon; C.dt = prm.time_step; C.vmax = prm.v_max; C.wmax = prm.w_max;
    C.inv_vmax = 1.0 / prm.v_max; C.inv_2wmax = 1.0 / (2.0 * prm.w_max); C.pose_jac_exact = prm.pose_jac_exact;
    const int K = C.K, k = g.lane;
    const bool act = k < K;
    const double inc = (prm.v_max * 0.99) * prm.time_step / prm.node_separation;

    // ---- per-problem state (registers) ----
    int n = -1;
    bool have = false, exhausted = false;
    Se2 T0, Mk, Tinit, Tk1;
    double uinit[2] = {0.0, 0.0}, u[2] = {0.0, 0.0};
    CoopRow R;
    double s = 1.0, lambda = 1e-7, curr = 0.0, prev = 0.0, gn = 0.0;
    int iter = 0, nb = 0, retries = 0, term = 0, lm_solves = 0, cost_evals = 0, zero_steps = 0;
    bool need_cost0 = true, need_build = false;
    T0.c = 1.0; T0.s = 0.0; T0.x = 0.0; T0.y = 0.0;
    Mk = T0; Tinit = T0; Tk1 = T0;

    for (;;) {
        // ---- fetch + initialise the next problem ----
        if (!have && !exhausted) {
            for (;;) {
                int t = 0;
                if (g.lane == 0) t = atomicAdd(B.queue, 1);
                t = __shfl_sync(g.mask, t, 0, G);
                if (t >= B.N) { exhausted = true; break; }
                if (B.active_mask == nullptr || B.active_mask[t]) { n = t; have = true; break; }
            }
            if (have) {
                int pid = B.path_id ? B.path_id[n] : 0;
                if (pid < 0 || pid >= B.num_paths) pid = 0;
                const double *path = B.paths + (size_t)pid * 3 * B.path_stride;
                const int L = B.lens[pid];
                const double *ps = B.pose + (size_t)n * B.pose_stride;
                T0 = se2_from_pose_inv(ps[0], ps[1], ps[2]);
                double target = B.interp_index[n] + inc;
                for (int i = 0; i < k && i < K; ++i) target += inc;      // accumulate like the reference loop (mpc.py:214,224)
                double r[3];
                mpc_ref_pose(path, B.path_stride, L, target, r);
                Mk = se2_from_pose_inv(r[0], r[1], r[2]);
                const bool first_call = (B.iteration0[n] != 0);
                if (first_call || !act) { Tinit = Mk; uinit[0] = prm.v_max / 2.0; uinit[1] = 0.0; }   // mpc.py:68-71
                else {
                    const double4 w = B.warm_T[(size_t)n * K + k];
                    const double2 uu = B.warm_u[(size_t)n * K + k];
                    Tinit.c = w.x; Tinit.s = w.y; Tinit.x = w.z; Tinit.y = w.w;
                    uinit[0] = uu.x; uinit[1] = uu.y;
                }
                Tk1 = Tinit; u[0] = uinit[0]; u[1] = uinit[1];
                s = 1.0; lambda = 1e-7; curr = 0.0; prev = 0.0; gn = 0.0;
                iter = 0; nb = 0; retries = 0; term = 0; lm_solves = 0; cost_evals = 0; zero_steps = 0;
                need_cost0 = true; need_build = false;
            }
        }
        if (!have) break;
        if (TRACE) { tr[6] += 1; tr[7] += __popc(__activemask()) / G; }
        DRE_TR(0);

        // ---- LM state machine: one damped solve + up to two trials per loop iteration (lev_marq...py:43-134) ----
        if (need_cost0) {
            const Se2 Tp = g.up(Tk1);
            const double c = act ? coop_lane_cost(C, k == 0 ? T0 : Tp, Tk1, u, Mk, 1.0 / s) : 0.0;
            curr = g.sum_ordered(c, K);
            need_cost0 = false; need_build = true;
        }
        DRE_TR(1);
        if (need_build) {
            ++iter; prev = curr; nb = 0; need_build = false;
            const Se2 Tp = g.up(Tk1);
            double pc6[6], pcb[3];
            coop_lane_build(C, k == 0 ? T0 : Tp, Tk1, u, Mk, s, k == 0, R, pc6, pcb);
            // this step's kinematic term also lands on the previous block's delta-delta part: receive from lane k+1
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double rb = g.down(pcb[a]);
#pragma unroll
                for (int c = 0; c <= a; ++c) {
                    const double rv = g.down(pc6[a * (a + 1) / 2 + c]);
                    if (k + 1 < K) R.D[(2 + a) * (3 + a) / 2 + 2 + c] += rv;
                }
                if (k + 1 < K) R.b[2 + a] += rb;
            }
            if (!act) {   // padding lanes: identity block, zero rhs
#pragma unroll
                for (int i = 0; i < 15; ++i) { R.D[i] = 0.0; R.O[i] = 0.0; }
                R.D[0] = R.D[2] = R.D[5] = R.D[9] = R.D[14] = 1.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) R.b[i] = 0.0;
            }
            double bb = 0.0;
#pragma unroll
            for (int a = 0; a < 5; ++a) bb += R.b[a] * R.b[a];
            gn = sqrt(g.sum_ordered(act ? bb : 0.0, K));
        }
        DRE_TR(2);
        double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        ++lm_solves;
        bool success = false;
        double proposed = 0.0;
        Se2 Tt = Tk1;
        double ut[2] = {u[0], u[1]};
        const bool solved = coop_solve(g, K, R, lambda, d);
        DRE_TR(3);
        if (solved) {
            // whole-step feasibility scale, walking u_0..u_{K-1}, v then w (lev_marq...py:52-91)
            // each lane prices its own two components; the walk itself only needs the quotients
            double allowed[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double dp = d[c], uc = u[c];
                const double hi = c == 0 ? C.vmax : C.wmax, lo = c == 0 ? 0.0 : -C.wmax;
                const double maxp = dp > 0.0 ? (hi - 1e-7) - uc : (lo + 1e-7) - uc;
                allowed[c] = (dp != 0.0 && act) ? maxp / dp : INFINITY;
            }
            double scale = 1.0;
            for (int kk = 0; kk < K; ++kk) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double al = g.bcast(allowed[c], kk);
                    scale = fmin(scale, al);
                    if (al < 1.0) scale = fmax(scale - 1e-7, 0.0);
                }
            }
#pragma unroll
            for (int a = 0; a < 5; ++a) d[a] *= scale;
            // scale == 0 (a control sits on its bound and the step pushes outward): the proposed state IS the current
            // state, both trials re-evaluate `curr` bit for bit, the predicted reduction is 0 and the ratio is defined as 0
            // (lev_marq...py:110-113) -> both are rejected. Skip the two cost evaluations, keep the counters.
            if (scale == 0.0) { cost_evals += 2; ++zero_steps; }
            for (int bt = 0; bt < 2 && scale != 0.0; ++bt) {
                if (bt > 0) {
#pragma unroll
                    for (int a = 0; a < 5; ++a) d[a] *= 0.8;
                }
                ut[0] = u[0] + d[0]; ut[1] = u[1] + d[1];
                Tt = se2_mul(se2_exp(d[2], d[3], d[4]), Tk1);
                const Se2 Tp = g.up(Tt);
                const double c = act ? coop_lane_cost(C, k == 0 ? T0 : Tp, Tt, ut, Mk, 1.0 / s) : 0.0;
                proposed = g.sum_ordered(c, K);
                ++cost_evals;
                // predicted reduction b^T p - 1/2 p^T A p with the undamped A (lev_marq...py:167-171)
                const double p2 = g.up(d[2]), p3 = g.up(d[3]), p4 = g.up(d[4]);
                double gg = 0.0, qq = 0.0;
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    gg += R.b[a] * d[a];
                    double r = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 < 5; ++c2) r += R.D[a >= c2 ? a * (a + 1) / 2 + c2 : c2 * (c2 + 1) / 2 + a] * d[c2];
                    if (k > 0) r += 2.0 * (R.O[3 * a] * p2 + R.O[3 * a + 1] * p3 + R.O[3 * a + 2] * p4);
                    qq += d[a] * r;
                }
                const double pred = g.sum_ordered(act ? gg : 0.0, K) - 0.5 * g.sum_ordered(act ? qq : 0.0, K);
                const double ratio = (pred == 0.0) ? 0.0 : (prev - proposed) / pred;
                if (ratio > 0.25) { success = true; break; }
            }
        }
        DRE_TR(4);
        if (success) {
            Tk1 = Tt; u[0] = ut[0]; u[1] = ut[1];
            lambda = fmax(lambda * 0.1, 1e-7);
            curr = proposed;
        } else {
            // Once lambda sits at its cap (lev_marq...py:126,131: min(10*lambda, 1e7)) a failed attempt repeats itself
            // verbatim -- same system, same step, same scale, same two rejected trials -- until max_shrink_steps runs out.
            // Fast-forward those identical attempts (the counters advance as if they had been executed).
            if (lambda >= 1e7) {
                const int rem = 50 - (nb + 1);
                lm_solves += rem;
                if (solved) cost_evals += 2 * rem;
                nb = 49;
            }
            lambda = fmin(lambda * 10.0, 1e7);
            ++nb;
        }
        if (success || nb >= 50) {
            // end of an outer iteration: STEAM's termination order
            if (!success && gn < 1e-6) term = 1;
            else if (!success) {
                if (retries >= prm.max_retries) term = -1;
                else {   // mpc.py:188-190: pose_error_scaling *= 5 and start over from the stored initial guess
                    s *= 5.0; ++retries;
                    Tk1 = Tinit; u[0] = uinit[0]; u[1] = uinit[1];
                    lambda = 1e-7; iter = 0; need_cost0 = true;
                }
            }
            else if (iter >= prm.max_iterations) term = 3;
            else if (curr <= 0.0) term = 4;
            else if (fabs(prev - curr) <= prm.abs_change_tol) term = 5;
            else if (fabs(prev - curr) / prev <= prm.rel_change_tol) term = 6;
            else need_build = true;
        }

        // ---- problem finished: outputs + warm start (mpc.py:147,192-205) ----
        if (term != 0) {
            const Se2 Tprev_final = g.up(Tk1);
            if (act) {
                double *ctl = B.controls + (size_t)n * B.controls_stride;
                if (2 * k < B.controls_count) ctl[2 * k] = u[0];
                if (2 * k + 1 < B.controls_count) ctl[2 * k + 1] = u[1];
                double4 *wT = B.warm_T + (size_t)n * K;
                double2 *wu = B.warm_u + (size_t)n * K;
                if (k + 1 < K) wT[k] = make_double4(Tk1.c, Tk1.s, Tk1.x, Tk1.y);                 // T_{k+1} -> slot k
                else {
                    const Se2 TL = (K >= 2) ? Tprev_final : T0;                                   // duplicate T_{K-1}
                    wT[k] = make_double4(TL.c, TL.s, TL.x, TL.y);
                }
                if (k >= 1) wu[k - 1] = make_double2(u[0], u[1]);
                if (k == K - 1) wu[K - 1] = make_double2(0.5, 0.0);
                if (k == 0) {
                    B.iteration0[n] = 0;
                    if (B.status) {
                        dre_mpc_status st;
                        st.iterations = iter; st.retries = retries; st.lm_solves = lm_solves; st.cost_evals = cost_evals;
                        st.term = term; st.zero_steps = zero_steps; st.final_cost = curr;
                        B.status[n] = st;
                    }
                }
            }
            have = false;
        }
        DRE_TR(5);
    }
    if (TRACE && B.trace != nullptr && g.lane == 0) {
        for (int i = 0; i < 8; ++i) atomicAdd(B.trace + i, tr[i]);
        atomicMax(B.trace + 8, tr[6]);
        atomicAdd(B.trace + 9, 1ull);
        atomicMax(B.trace + 10, (unsigned long long)(clock64() - tr_start));
    }
#undef DRE_TR
}
#endif  // __CUDACC__
