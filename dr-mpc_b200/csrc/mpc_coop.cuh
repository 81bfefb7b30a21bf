// mpc_coop.cuh -- lane-per-horizon-step cooperative LM solve (the fast path of the MPC kernel).
//
// A group of G lanes (G = 4, 8, 16, 32, or 64 = a pair of warps, G >= K) owns one problem; lane k owns horizon step k:
// pose T_{k+1}, control u_k, reference M_k and block row k of the block-tridiagonal normal equations (diagonal block D_k
// packed 5x5, coupling block O_k 5x3, rhs b_k). Residuals, Jacobians and trial costs are embarrassingly parallel over
// lanes (the neighbour pose comes by shuffle). The damped solve first eliminates every step's two controls in parallel,
// then runs a block Cholesky over the remaining 3x3 pose blocks as K pipeline stages in which lane j consumes lane j-1's
// factor rows (9 doubles by shuffle); sums over the horizon are taken in step order so they round like the sequential
// reference.
//
// What bounds the kernel (profiles/README.md): in steady tracking the dependent fp64 chains of its longest problem; when
// every warp is busy (the steps after a path switch) instruction issue and, before the code-size diet, instruction FETCH
// (154 KB of SASS, 59 % i-cache hit rate). Hence:
//   * the linearisation (D, O, b: 35 doubles per lane, read-only between rebuilds), the retry guess and the barrier values
//     live in a conflict-free shared-memory column per thread instead of registers -> 168 registers, 3 CTAs per SM;
//   * every pose carries its heading, updated additively by the LM step, so the two log maps per cost / build take their
//     angle from a subtraction + wrap instead of atan2;
//   * sin/cos of the small angles (dt*w and the LM step) use a Taylor kernel, ln x on (0, 1] a 45-instruction routine, with
//     the polynomial coefficients in constant memory; the full-range routines sit out of line;
//   * the factors store inverse pivots (one rsqrt per pivot, no divisions on the chain).
//
// Control flow is a per-group state machine with the phases in a fixed order (initial cost -> build -> damped solve ->
// trial 1 -> trial 2 -> bookkeeping), so the groups of a warp, which follow different LM paths, re-converge at every
// loop iteration instead of serialising whole solves.
#pragma once
#include "mpc_device.cuh"

#if defined(__CUDACC__)

#ifndef DRE_MPC_CTA
#define DRE_MPC_CTA 128   // threads per CTA of the cooperative kernel (shared-memory columns are strided by this)
#endif

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
    __device__ void attach(double *) {}
    __device__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ int bcast_i(int v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ double up(double v) const { return __shfl_up_sync(mask, v, 1, G); }
    __device__ double down(double v) const { return __shfl_down_sync(mask, v, 1, G); }
    // N values from lane-1 / lane+1 (`boundary`: only the two-warp group cares, see LaneGroup<64>)
    template <int N> __device__ void up_n(const double *in, double *out, bool boundary = true) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = up(in[i]);
    }
    template <int N> __device__ void down_n(const double *in, double *out, bool boundary = true) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = down(in[i]);
    }
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
};

// K > 32: one problem per PAIR of warps (lane = threadIdx & 63). Inside a warp everything is a shuffle as above; what crosses
// the warp boundary (lane 31 <-> lane 32, broadcasts, votes, sums) goes through a small shared-memory exchange area and a
// named barrier of the pair's 64 threads. The area is double-buffered by the parity of the operation count, so one barrier
// per operation is enough (a warp cannot get two operations ahead of its partner). The pipeline stages of the solve hand
// data across the boundary only at stage 32 (forward) / 31 (backward): the other stages pass `boundary = false`.
template <>
struct LaneGroup<64> {
    unsigned mask;
    int lane;
    double *xch;          // [2][16] doubles of this pair
    mutable int par;
    int bar_id;
    __device__ LaneGroup() : mask(0xffffffffu), lane((int)(threadIdx.x & 63u)), xch(nullptr), par(0), bar_id(1 + (int)(threadIdx.x >> 6)) {}
    __device__ void attach(double *area) { xch = area + (threadIdx.x >> 6) * 32; }
    __device__ void pair_sync() const { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); }
    __device__ double *buf() const { double *b = xch + par * 16; par ^= 1; return b; }
    __device__ double bcast(double v, int src) const
    {
        double *b = buf();
        if (lane == src) b[0] = v;
        pair_sync();
        return b[0];
    }
    __device__ int bcast_i(int v, int src) const
    {
        double *b = buf();
        if (lane == src) reinterpret_cast<int *>(b)[0] = v;
        pair_sync();
        return reinterpret_cast<int *>(b)[0];
    }
    template <int N> __device__ void up_n(const double *in, double *out, bool boundary = true) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = __shfl_up_sync(0xffffffffu, in[i], 1);
        if (boundary) {
            double *b = buf();
            if (lane == 31) {
#pragma unroll
                for (int i = 0; i < N; ++i) b[i] = in[i];
            }
            pair_sync();
            if (lane == 32) {
#pragma unroll
                for (int i = 0; i < N; ++i) out[i] = b[i];
            }
        }
    }
    template <int N> __device__ void down_n(const double *in, double *out, bool boundary = true) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = __shfl_down_sync(0xffffffffu, in[i], 1);
        if (boundary) {
            double *b = buf();
            if (lane == 32) {
#pragma unroll
                for (int i = 0; i < N; ++i) b[i] = in[i];
            }
            pair_sync();
            if (lane == 31) {
#pragma unroll
                for (int i = 0; i < N; ++i) out[i] = b[i];
            }
        }
    }
    __device__ double up(double v) const { double r; up_n<1>(&v, &r); return r; }
    __device__ double down(double v) const { double r; down_n<1>(&v, &r); return r; }
    __device__ bool any(bool p) const
    {
        const unsigned w = __ballot_sync(0xffffffffu, p);
        unsigned *b = reinterpret_cast<unsigned *>(buf());
        if ((lane & 31) == 0) b[lane >> 5] = w;
        pair_sync();
        return (b[0] | b[1]) != 0u;
    }
    __device__ double sum_ordered(double v, int) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        double *b = buf();
        if ((lane & 31) == 0) b[lane >> 5] = v;
        pair_sync();
        return b[0] + b[1];
    }
};

// pose with its heading: th == atan2(s, c) modulo 2 pi, tracked additively
struct Se2h { Se2 T; double th; };

template <int G>
__device__ __forceinline__ Se2h lane_up(const LaneGroup<G> &g, const Se2h &a)
{
    const double in[5] = {a.T.c, a.T.s, a.T.x, a.T.y, a.th};
    double out[5];
    g.template up_n<5>(in, out);
    Se2h r;
    r.T.c = out[0]; r.T.s = out[1]; r.T.x = out[2]; r.T.y = out[3]; r.th = out[4];
    return r;
}

// angle of a rotation known modulo 2 pi -> [-pi, pi] like atan2
__device__ __forceinline__ double wrap_angle(double a)
{
    return a - (2.0 * DRE_PI) * rint(a * (1.0 / (2.0 * DRE_PI)));
}

// Polynomial coefficients live in constant memory: a 64-bit immediate costs two extra MOVs per DFMA, a constant-bank
// operand costs none (the kernel is instruction-fetch bound, see sincos_full).
static __constant__ double DRE_PS[7] = {-1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,
                                        -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0};                       // sin: -1/15! .. -1/3!
static __constant__ double DRE_PC[7] = {1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0,
                                        1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0};                      // cos: 1/16! .. 1/4!
static __constant__ double DRE_PL[12] = {2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0,
                                         2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0, 1.9082149292705877e-10, 0.693147180369123816490};   // 2 atanh, ln2 lo / hi

// One out-of-line copy of the full-range sincos (rare: |x| > 0.5). The kernel is instruction-fetch bound when every warp
// is busy (154 KB of SASS before the diet: 59 % i-cache hit rate in the post-path-switch steps), so rare paths stay out
// of the hot loop's footprint.
__device__ __noinline__ void sincos_full(double x, double *sn, double *cs) { sincos(x, sn, cs); }

__device__ __noinline__ double atan2_full(double y, double x) { return atan2(y, x); }
// se2_from_pose_inv of mpc_device.cuh through the shared sincos copy (once per problem, not in the LM loop)
__device__ __forceinline__ Se2 se2_from_pose_inv_ool(double x, double y, double th)
{
    double sn, cs;
    sincos_full(th, &sn, &cs);
    Se2 r;
    r.c = cs; r.s = -sn;
    r.x = -(cs * x + sn * y);
    r.y = sn * x - cs * y;
    return r;
}

// sin and cos for |x| <= 0.5 by Taylor series (truncation < 5e-17 relative); larger arguments fall back
__device__ __forceinline__ void sincos_small(double x, double *sn, double *cs)
{
    if (fabs(x) > 0.5) { sincos_full(x, sn, cs); return; }
    const double z = x * x;
    double ps = DRE_PS[0], pc = DRE_PC[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) { ps = fma(ps, z, DRE_PS[i]); pc = fma(pc, z, DRE_PC[i]); }
    *sn = fma(x * z, ps, x);
    *cs = fma(z * z, pc, fma(z, -0.5, 1.0));
}

// 1/x from the hardware seed (rcp.approx: ~2^-23) and two Newton steps: within 1-2 ulp, 5 instructions instead of the
// ~25 of an IEEE division. Every quotient of the LM attempt goes through it (each fp64 instruction costs the scheduler's
// fp64 pipe two cycles); none of them needs correct rounding: they feed cost terms, Jacobians and threshold tests.
__device__ __forceinline__ double rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
// ln x for 0 < x <= 1 in ~35 instructions (CUDA's log() is ~90 and there are four per cost evaluation): x = 2^e m with
// m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh(s), s = (m - 1)/(m + 1), |s| < 0.1716 -> 10 odd terms (truncation < 3e-17).
// Measured against log() over (1e-9, 1]: <= 2 ulp (tests/test_gpu_mpc.py). Subnormal arguments take the library routine.
__device__ __noinline__ double log_full(double x) { return log(x); }
__device__ __forceinline__ double log_unit(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    if (hi < 0x00100000) return log_full(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = f * rcp_fast(2.0 + f);
    const double z = s * s;
    double q = DRE_PL[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) q = fma(q, z, DRE_PL[i]);
    const double ed = (double)e;
    // ln 2 split: hi part exact in 32 bits of mantissa so ed * hi is exact
    const double r = fma(s * z, q, fma(ed, DRE_PL[10], 2.0 * s));
    return fma(ed, DRE_PL[11], r);
}
// barrier_val of mpc_device.cuh with the compact logarithm
__device__ __forceinline__ double barrier_val_fast(double x) { return x <= 0.0 ? -500.0 : (x > 1.0 ? 500.0 : log_unit(x)); }

// half_cot / phi_coef / barrier_grad of mpc_device.cuh with reciprocal-multiplies
__device__ __forceinline__ double half_cot_fast(double p, double sn, double cs)
{
    if (fabs(p) < 1e-3) { const double p2 = p * p; return 1.0 - p2 * (1.0 / 12.0) - p2 * p2 * (1.0 / 720.0); }
    return cs >= 0.0 ? 0.5 * p * (1.0 + cs) * rcp_fast(sn) : 0.5 * p * sn * rcp_fast(1.0 - cs);
}
__device__ __forceinline__ PhiCoef phi_coef_fast(double p, double sn, double cs)
{
    PhiCoef q;
    if (fabs(p) < 1e-3) {
        const double p2 = p * p;
        q.A = 1.0 - p2 * (1.0 / 6.0) + p2 * p2 * (1.0 / 120.0);
        q.beta = 0.5 - p2 * (1.0 / 24.0) + p2 * p2 * (1.0 / 720.0);
        q.B = p * q.beta;
        q.alpha = p * (1.0 / 6.0 - p2 * (1.0 / 120.0) + p2 * p2 * (1.0 / 5040.0));
    } else {
        const double ip = rcp_fast(p);
        q.A = sn * ip;
        q.B = (1.0 - cs) * ip;
        q.beta = q.B * ip;
        q.alpha = (p - sn) * ip * ip;
    }
    return q;
}
__device__ __forceinline__ Aff se2_Jlinv_fast(const double e[3], double sn, double cs, double ct)
{
    const PhiCoef q = phi_coef_fast(e[2], sn, cs);
    const double h = 0.5 * e[2];
    const double qx = q.alpha * e[0] + q.beta * e[1], qy = q.alpha * e[1] - q.beta * e[0];
    Aff a;
    a.m[0] = ct; a.m[1] = h;  a.m[2] = -(ct * qx + h * qy);
    a.m[3] = -h; a.m[4] = ct; a.m[5] = h * qx - ct * qy;
    return a;
}
__device__ __forceinline__ double barrier_grad_fast(double x) { return (x <= 0.0 || x > 1.0) ? 10.0 : rcp_fast(x); }

// log map of T whose rotation angle p (wrapped) is already known
__device__ __forceinline__ void se2_log_known(const Se2 &T, double p, double e[3], double &ct)
{
    ct = half_cot_fast(p, T.s, T.c);
    const double h = 0.5 * p;
    e[0] = ct * T.x + h * T.y;
    e[1] = ct * T.y - h * T.x;
    e[2] = p;
}

// exp(rx, ry, p) * T with the heading carried along (the LM update T <- exp(delta^) T)
__device__ __forceinline__ Se2h se2h_step(double rx, double ry, double p, const Se2h &T)
{
    double sn, cs;
    sincos_small(p, &sn, &cs);
    const PhiCoef q = phi_coef_fast(p, sn, cs);
    Se2 E;
    E.c = cs; E.s = sn;
    E.x = q.A * rx - q.B * ry;
    E.y = q.B * rx + q.A * ry;
    Se2h r;
    r.T = se2_mul(E, T.T);
    r.th = T.th + p;
    return r;
}

// per-lane cost of step k: pose + kinematic + 4 barriers (mpc.py:99-142)
// bv[0..3]: the four barrier values at u, reused by the linearisation if this state becomes the current one
__device__ __forceinline__ double coop_lane_cost(const MpcConst &C, const Se2h &Tprev, const Se2h &Tk1, const double u[2],
                                                 const Se2h &Mk, double inv_s, double bv[4])
{
    double e[3], ct;
    se2_log_known(se2_mul_inv(Mk.T, Tk1.T), wrap_angle(Mk.th - Tk1.th), e, ct);
    double cost = 0.5 * inv_s * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double p = C.dt * u[1];
    double sn, cs;
    sincos_small(p, &sn, &cs);
    const PhiCoef q = phi_coef_fast(p, sn, cs);
    const Se2 E = se2_mul(se2_mul_inv(Tk1.T, Tprev.T), se2_exp_x(C.dt * u[0], p, sn, cs, q));
    se2_log_known(E, wrap_angle(Tk1.th - Tprev.th + p), e, ct);
    cost += 0.5 * 1000.0 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double b0 = barrier_val_fast((C.vmax - u[0]) * C.inv_vmax), b1 = barrier_val_fast(u[0] * C.inv_vmax);
    const double b2 = barrier_val_fast((C.wmax - u[1]) * C.inv_2wmax), b3 = barrier_val_fast((u[1] + C.wmax) * C.inv_2wmax);
    cost += 0.5 * 0.1 * (1.0 / 250000.0) * (b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
    bv[0] = b0; bv[1] = b1; bv[2] = b2; bv[3] = b3;
    return cost;
}

// this thread's shared-memory column: slot i lives at col[i * DRE_MPC_CTA] (consecutive threads -> consecutive banks)
struct CoopCol {
    double *col;
    __device__ double &operator[](int i) const { return col[i * DRE_MPC_CTA]; }
};
enum { CS_D = 0, CS_O = 15, CS_B = 30, CS_TI = 35, CS_UI = 40, CS_BV = 42, CS_SLOTS = 50 };   // D 15, O 15, b 5, Tinit (c,s,x,y,th), uinit 2,
                                                                                              // barrier values of the current / the trial state 2 x 4

struct CoopFactor {
    double L[6];    // Cholesky factor of the reduced 3x3 pose pivot block (packed lower; diagonal slots hold INVERSE pivots)
    double B[9];    // reduced coupling block, then (in place) coupling * L_{prev}^{-T}
};

// Build block row k into the shared-memory column (lane 0's T_k is the locked current pose: no coupling and no
// contribution to a previous block). Returns this lane's |b_k|^2.
template <int G>
__device__ __forceinline__ double coop_lane_build(const LaneGroup<G> &g, const MpcConst &C, const Se2h &Tprev, const Se2h &Tk1,
                                                  const double u[2], const Se2h &Mk, double s, bool act, const CoopCol &M, int bv_slot)
{
    const int k = g.lane;
    const bool first = k == 0;
    double D[15], O[15], b[5], pc6[6], pcb[3];
#pragma unroll
    for (int i = 0; i < 15; ++i) { D[i] = 0.0; O[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 5; ++i) b[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) pc6[i] = 0.0;
    pcb[0] = pcb[1] = pcb[2] = 0.0;
    {   // pose term
        double e[3], ct;
        const Se2 E = se2_mul_inv(Mk.T, Tk1.T);
        se2_log_known(E, wrap_angle(Mk.th - Tk1.th), e, ct);
        Aff Gm = se2_Jlinv_fast(e, E.s, E.c, ct);
        if (C.pose_jac_exact) Gm = aff_mul(Gm, se2_Ad(E));
        const double r0[3] = {-Gm.m[0], -Gm.m[1], -Gm.m[2]}, r1[3] = {-Gm.m[3], -Gm.m[4], -Gm.m[5]}, r2[3] = {0.0, 0.0, -1.0};
        acc_block<3>(D, b, 2, r0, r1, r2, e, 3.0 / s);
    }
    {   // kinematic term
        const double p = C.dt * u[1];
        double sn, cs;
        sincos_small(p, &sn, &cs);
        const PhiCoef q = phi_coef_fast(p, sn, cs);
        const double rx = C.dt * u[0];
        const Se2 Lh = se2_mul_inv(Tk1.T, Tprev.T);
        const Se2 E = se2_mul(Lh, se2_exp_x(rx, p, sn, cs, q));
        double e[3], ct;
        se2_log_known(E, wrap_angle(Tk1.th - Tprev.th + p), e, ct);
        const Aff Gm = se2_Jlinv_fast(e, E.s, E.c, ct);
        Aff Jl;
        Jl.m[0] = q.A; Jl.m[1] = q.B;  Jl.m[2] = q.alpha * rx;
        Jl.m[3] = -q.B; Jl.m[4] = q.A; Jl.m[5] = q.beta * rx;
        const Aff Ju = aff_mul(aff_mul(Gm, se2_Ad(E)), Jl);
        const double r0[5] = {C.dt * Ju.m[0], C.dt * Ju.m[2], Gm.m[0], Gm.m[1], Gm.m[2]};
        const double r1[5] = {C.dt * Ju.m[3], C.dt * Ju.m[5], Gm.m[3], Gm.m[4], Gm.m[5]};
        const double r2[5] = {0.0, C.dt, 0.0, 0.0, 1.0};
        acc_block<5>(D, b, 0, r0, r1, r2, e, 1000.0);
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
                for (int c = 0; c < 3; ++c) O[3 * a + c] = 1000.0 * (r0[a] * q0[c] + r1[a] * q1[c] + r2[a] * q2[c]);
        }
    }
    {   // barriers
        const double v = u[0], w = u[1];
        const double x0 = (C.vmax - v) * C.inv_vmax, x1 = v * C.inv_vmax;
        const double x2 = (C.wmax - w) * C.inv_2wmax, x3 = (w + C.wmax) * C.inv_2wmax;
        const double k500 = 1.0 / 500.0;
        // the barrier values of this state were computed by the cost evaluation that made it the current state
        const double e0 = M[bv_slot] * k500, e1 = M[bv_slot + 1] * k500, e2 = M[bv_slot + 2] * k500, e3 = M[bv_slot + 3] * k500;
        const double j0 = -k500 * barrier_grad_fast(x0) * C.inv_vmax, j1 = k500 * barrier_grad_fast(x1) * C.inv_vmax;
        const double j2 = -k500 * barrier_grad_fast(x2) * C.inv_2wmax, j3 = k500 * barrier_grad_fast(x3) * C.inv_2wmax;
        b[0] -= 0.1 * (j0 * e0 + j1 * e1);
        b[1] -= 0.1 * (j2 * e2 + j3 * e3);
        D[0] += 0.1 * (j0 * j0 + j1 * j1);
        D[2] += 0.1 * (j2 * j2 + j3 * j3);
    }
    // this step's kinematic term also lands on the previous block's delta-delta part: receive from lane k+1
    {
        const double snd[9] = {pc6[0], pc6[1], pc6[2], pc6[3], pc6[4], pc6[5], pcb[0], pcb[1], pcb[2]};
        double rcv[9];
        g.template down_n<9>(snd, rcv);
        if (k + 1 < C.K) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int c = 0; c <= a; ++c) D[(2 + a) * (3 + a) / 2 + 2 + c] += rcv[a * (a + 1) / 2 + c];
                b[2 + a] += rcv[6 + a];
            }
        }
    }
    if (!act) {   // padding lanes: identity block, zero rhs
#pragma unroll
        for (int i = 0; i < 15; ++i) { D[i] = 0.0; O[i] = 0.0; }
        D[0] = D[2] = D[5] = D[9] = D[14] = 1.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) b[i] = 0.0;
    }
    double bb = 0.0;
#pragma unroll
    for (int i = 0; i < 15; ++i) { M[CS_D + i] = D[i]; M[CS_O + i] = O[i]; }
#pragma unroll
    for (int a = 0; a < 5; ++a) { M[CS_B + a] = b[a]; bb += b[a] * b[a]; }
    return bb;
}

// Lu^{-1} applied to the control rows of block row k: Wv = Lu^{-1} V^T, Wo = Lu^{-1} O_u, wb = Lu^{-1} b_u (Lu = [1/i0, 0; l10, 1/i1])
__device__ __forceinline__ void coop_control_rows(const CoopCol &M, double i0, double l10, double i1, double wv0[3], double wv1[3],
                                                  double wo0[3], double wo1[3], double &wb0, double &wb1)
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        wv0[a] = M[CS_D + (2 + a) * (3 + a) / 2] * i0;
        wv1[a] = (M[CS_D + (2 + a) * (3 + a) / 2 + 1] - l10 * wv0[a]) * i1;
        wo0[a] = M[CS_O + a] * i0;
        wo1[a] = (M[CS_O + 3 + a] - l10 * wo0[a]) * i1;
    }
    wb0 = M[CS_B] * i0;
    wb1 = (M[CS_B + 1] - l10 * wb0) * i1;
}

// Damped block-tridiagonal solve (A with diag * (1 + lambda)) d = b. Returns false (group-uniform) on a non-positive pivot.
//
// Block row k couples [u_k (2); p_k (3)] with p_{k-1} only, so the controls are eliminated first, by every lane at once
// (2x2 Cholesky, Schur complement onto p_k and, by one shuffle, onto p_{k-1}); what is left for the K sequential pipeline
// stages is a block-tridiagonal system of 3x3 blocks in the poses -- less than half the dependent fp64 chain of 5x5 stages.
// Diagonal slots of the factors hold the INVERSE pivots (one rsqrt per pivot; every later use is a multiplication).
template <int G>
__device__ __forceinline__ bool coop_solve(const LaneGroup<G> &g, int K, const CoopCol &M, CoopFactor &R, double lambda, double d[5])
{
    const double damp = 1.0 + lambda;
    bool ok = true;
    const bool has_next = g.lane + 1 < K;
    // ---- (1) controls: U = Lu Lu^T, Wv = Lu^{-1} V^T, Wo = Lu^{-1} O_u, wb = Lu^{-1} b_u
    double u00 = M[CS_D + 0] * damp, t11;
    if (!(u00 > 0.0)) { ok = false; u00 = 1.0; }
    const double i0 = rsqrt(u00), l10 = M[CS_D + 1] * i0;
    t11 = M[CS_D + 2] * damp - l10 * l10;
    if (!(t11 > 0.0)) { ok = false; t11 = 1.0; }
    const double i1 = rsqrt(t11);
    double wv0[3], wv1[3], wo0[3], wo1[3], wb0, wb1;
    coop_control_rows(M, i0, l10, i1, wv0, wv1, wo0, wo1, wb0, wb1);
    // reduced pose system: P_k = P - Wv^T Wv - (Wo^T Wo of lane k+1), C_k = O_p - Wv^T Wo, r_k = b_p - Wv^T wb - (Wo^T wb of lane k+1)
    double P[6], r[3];
    {
        double snd[9], rcv[9];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c <= a; ++c) snd[a * (a + 1) / 2 + c] = wo0[a] * wo0[c] + wo1[a] * wo1[c];
            snd[6 + a] = wo0[a] * wb0 + wo1[a] * wb1;
        }
        g.template down_n<9>(snd, rcv);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c <= a; ++c) {
                double v = M[CS_D + (2 + a) * (3 + a) / 2 + 2 + c];
                if (a == c) v *= damp;
                v -= wv0[a] * wv0[c] + wv1[a] * wv1[c];
                if (has_next) v -= rcv[a * (a + 1) / 2 + c];
                P[a * (a + 1) / 2 + c] = v;
            }
            double v = M[CS_B + 2 + a] - (wv0[a] * wb0 + wv1[a] * wb1);
            if (has_next) v -= rcv[6 + a];
            r[a] = v;
#pragma unroll
            for (int c = 0; c < 3; ++c) R.B[3 * a + c] = M[CS_O + 3 * (2 + a) + c] - (wv0[a] * wo0[c] + wv1[a] * wo1[c]);
        }
    }
    // ---- (2) forward pipeline over the 3x3 pose blocks
    double y[3] = {0.0, 0.0, 0.0};
    R.L[0] = R.L[2] = R.L[5] = 1.0; R.L[1] = R.L[3] = R.L[4] = 0.0;
    for (int j = 0; j < K; ++j) {
        const double snd[9] = {R.L[0], R.L[1], R.L[2], R.L[3], R.L[4], R.L[5], y[0], y[1], y[2]};
        double rcv[9];
        g.template up_n<9>(snd, rcv, j == 32);
        const double q00 = rcv[0], q10 = rcv[1], q11 = rcv[2], q20 = rcv[3], q21 = rcv[4], q22 = rcv[5];
        const double y0 = rcv[6], y1 = rcv[7], y2 = rcv[8];
        if (g.lane == j) {
            if (j > 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double b0 = R.B[3 * a] * q00;
                    const double b1 = (R.B[3 * a + 1] - b0 * q10) * q11;
                    const double b2 = (R.B[3 * a + 2] - b0 * q20 - b1 * q21) * q22;
                    R.B[3 * a] = b0; R.B[3 * a + 1] = b1; R.B[3 * a + 2] = b2;
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
#pragma unroll
                    for (int c = 0; c <= a; ++c)
                        P[a * (a + 1) / 2 + c] -= R.B[3 * a] * R.B[3 * c] + R.B[3 * a + 1] * R.B[3 * c + 1] + R.B[3 * a + 2] * R.B[3 * c + 2];
                    r[a] -= R.B[3 * a] * y0 + R.B[3 * a + 1] * y1 + R.B[3 * a + 2] * y2;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double dd = P[c * (c + 1) / 2 + c];
#pragma unroll
                for (int m = 0; m < c; ++m) dd -= R.L[c * (c + 1) / 2 + m] * R.L[c * (c + 1) / 2 + m];
                if (!(dd > 0.0)) { ok = false; dd = 1.0; }
                const double ild = rsqrt(dd);
                R.L[c * (c + 1) / 2 + c] = ild;
#pragma unroll
                for (int a = c + 1; a < 3; ++a) {
                    double v = P[a * (a + 1) / 2 + c];
#pragma unroll
                    for (int m = 0; m < c; ++m) v -= R.L[a * (a + 1) / 2 + m] * R.L[c * (c + 1) / 2 + m];
                    R.L[a * (a + 1) / 2 + c] = v * ild;
                }
            }
            y[0] = r[0] * R.L[0];
            y[1] = (r[1] - R.L[1] * y[0]) * R.L[2];
            y[2] = (r[2] - R.L[3] * y[0] - R.L[4] * y[1]) * R.L[5];
        }
    }
    ok = !g.any(!ok);
    // ---- (3) back substitution over the poses: L_j^T p_j = y_j - B_{j+1}^T p_{j+1}
    double p[3] = {0.0, 0.0, 0.0}, t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int j = K - 1; j >= 0; --j) {
        const double tsnd[3] = {t0, t1, t2};
        double trcv[3];
        g.template down_n<3>(tsnd, trcv, j == 31);
        const double r0 = trcv[0], r1 = trcv[1], r2 = trcv[2];
        if (g.lane == j) {
            double h0 = y[0], h1 = y[1], h2 = y[2];
            if (j + 1 < K) { h0 -= r0; h1 -= r1; h2 -= r2; }
            p[2] = h2 * R.L[5];
            p[1] = (h1 - R.L[4] * p[2]) * R.L[2];
            p[0] = (h0 - R.L[1] * p[1] - R.L[3] * p[2]) * R.L[0];
            t0 = R.B[0] * p[0] + R.B[3] * p[1] + R.B[6] * p[2];
            t1 = R.B[1] * p[0] + R.B[4] * p[1] + R.B[7] * p[2];
            t2 = R.B[2] * p[0] + R.B[5] * p[1] + R.B[8] * p[2];
        }
    }
    // ---- (4) controls back: Lu^T u = wb - Wv p_k - Wo p_{k-1}   (lane 0 has Wo = 0)
    // (Wv, Wo, wb are recomputed from the shared-memory column: 21 multiply-adds instead of 14 registers held across the pipeline)
    double ppv[3];
    g.template up_n<3>(p, ppv);
    const double pp0 = ppv[0], pp1 = ppv[1], pp2 = ppv[2];
    double xv0[3], xv1[3], xo0[3], xo1[3], xb0, xb1;
    coop_control_rows(M, i0, l10, i1, xv0, xv1, xo0, xo1, xb0, xb1);
    const double w0 = xb0 - (xv0[0] * p[0] + xv0[1] * p[1] + xv0[2] * p[2]) - (xo0[0] * pp0 + xo0[1] * pp1 + xo0[2] * pp2);
    const double w1 = xb1 - (xv1[0] * p[0] + xv1[1] * p[1] + xv1[2] * p[2]) - (xo1[0] * pp0 + xo1[1] * pp1 + xo1[2] * pp2);
    d[1] = w1 * i1;
    d[0] = (w0 - l10 * d[1]) * i0;
    d[2] = p[0]; d[3] = p[1]; d[4] = p[2];
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
// (damped solve + up to two trials) per loop iteration. LM paths differ per problem (3 to >100 attempts), and the 32/G
// groups of a warp advance in lock-step per iteration; letting a finished group fetch the next problem immediately
// keeps every group busy instead of idling until the slowest problem of its warp is done.
// Semantics per problem: identical to mpc_act_one (mpc_device.cuh) = reference MPC.act (mpc.py:58-206).
// smem: CS_SLOTS * DRE_MPC_CTA doubles.
template <int G, bool TRACE = false>
__device__ void mpc_coop_persistent(const LaneGroup<G> &g, const dre_mpc_params &prm, const CoopBatch &B, double *smem)
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
    CoopCol M;
    M.col = smem + threadIdx.x;

    // ---- per-problem state (registers) ----
    int n = -1;
    bool have = false, exhausted = false;
    const int warps_per_cta = G > 32 ? DRE_MPC_CTA / G : DRE_MPC_CTA / 32;   // G = 64: a pair of warps counts as one
    const int total_groups = (int)gridDim.x * (DRE_MPC_CTA / G), total_warps = (int)gridDim.x * warps_per_cta;
    const int my_warp = (int)blockIdx.x * warps_per_cta + (G > 32 ? (int)threadIdx.x / G : (int)threadIdx.x / 32);
    const int my_slot = G >= 32 ? 0 : ((int)threadIdx.x & 31) / G;
    int first_rank = my_slot * total_warps + my_warp;   // < total_groups
    Se2h T0, Mk, Tk1;
    double u[2] = {0.0, 0.0};
    CoopFactor R;
    double s = 1.0, lambda = 1e-7, curr = 0.0, prev = 0.0, gn = 0.0;
    int iter = 0, nb = 0, retries = 0, term = 0, lm_solves = 0, cost_evals = 0, zero_steps = 0;
    bool need_cost0 = true, need_build = false;
    int bv_cur = CS_BV;   // which of the two barrier-value slot sets belongs to the current state
    T0.T.c = 1.0; T0.T.s = 0.0; T0.T.x = 0.0; T0.T.y = 0.0; T0.th = 0.0;
    Mk = T0; Tk1 = T0;

    for (;;) {
        // ---- fetch + initialise the next problem ----
        if (!have && !exhausted) {
            for (;;) {
                // A group's first problem is assigned statically, slot-major: slot 0 of every warp of the grid first, then slot
                // 1, ... so a batch smaller than the grid's group count is spread one problem per warp (a warp pays for every
                // phase any of its groups needs, and a small batch lasts as long as its slowest warp). Later problems come from
                // the shared counter, which therefore starts behind the static range.
                int t = first_rank;
                if (t < 0) {
                    if (g.lane == 0) t = atomicAdd(B.queue, 1) + total_groups;
                    t = g.bcast_i(t, 0);
                }
                first_rank = -1;
                if (t >= B.N) { exhausted = true; break; }
                if (B.active_mask == nullptr || B.active_mask[t]) { n = t; have = true; break; }
            }
            if (have) {
                int pid = B.path_id ? B.path_id[n] : 0;
                if (pid < 0 || pid >= B.num_paths) pid = 0;
                const double *path = B.paths + (size_t)pid * 3 * B.path_stride;
                const int L = B.lens[pid];
                const double *ps = B.pose + (size_t)n * B.pose_stride;
                T0.T = se2_from_pose_inv_ool(ps[0], ps[1], ps[2]);
                T0.th = -ps[2];
                double target = B.interp_index[n] + inc;
                for (int i = 0; i < k && i < K; ++i) target += inc;      // accumulate like the reference loop (mpc.py:214,224)
                double r[3];
                mpc_ref_pose(path, B.path_stride, L, target, r);
                Mk.T = se2_from_pose_inv_ool(r[0], r[1], r[2]);
                Mk.th = -r[2];
                const bool first_call = (B.iteration0[n] != 0);
                // the stored warm start of the last successful solve (all zeros until one has finished)
                double4 w = make_double4(0.0, 0.0, 0.0, 0.0);
                double2 uu = make_double2(0.0, 0.0);
                if (act) { w = B.warm_T[(size_t)n * K + k]; uu = B.warm_u[(size_t)n * K + k]; }
                const bool has_stored = act && (w.x * w.x + w.y * w.y > 0.5);
                if (first_call || !act) { Tk1 = Mk; u[0] = prm.v_max / 2.0; u[1] = 0.0; }   // mpc.py:68-71
                else {
                    Tk1.T.c = w.x; Tk1.T.s = w.y; Tk1.T.x = w.z; Tk1.T.y = w.w;
                    Tk1.th = atan2_full(w.y, w.x);
                    u[0] = uu.x; u[1] = uu.y;
                }
                // The guess the pose_error_scaling retry starts from (mpc.py:188-190 -> :73-75). Reference quirk: iteration0
                // is cleared BEFORE the solve (mpc.py:72), so the retry of a failed FIRST solve deep-copies the stored warm
                // start -- the one from before the reset(). With nothing stored the reference raises outside its try; the
                // batch uses the iteration0 guess again.
                if (first_call && has_stored) {
                    M[CS_TI] = w.x; M[CS_TI + 1] = w.y; M[CS_TI + 2] = w.z; M[CS_TI + 3] = w.w; M[CS_TI + 4] = atan2_full(w.y, w.x);
                    M[CS_UI] = uu.x; M[CS_UI + 1] = uu.y;
                } else {
                    M[CS_TI] = Tk1.T.c; M[CS_TI + 1] = Tk1.T.s; M[CS_TI + 2] = Tk1.T.x; M[CS_TI + 3] = Tk1.T.y; M[CS_TI + 4] = Tk1.th;
                    M[CS_UI] = u[0]; M[CS_UI + 1] = u[1];
                }
                s = 1.0; lambda = 1e-7; curr = 0.0; prev = 0.0; gn = 0.0;
                iter = 0; nb = 0; retries = 0; term = 0; lm_solves = 0; cost_evals = 0; zero_steps = 0;
                need_cost0 = true; need_build = false;
            }
        }
#ifndef DRE_MPC_NO_CTA_SYNC
        // keep the CTA's warps on the same stretch of code (shared i-cache lines): one barrier per LM attempt
        if (!__syncthreads_or(have ? 1 : 0)) break;
        if (!have) continue;
#else
        if (!have) break;
#endif
        if (TRACE) { tr[6] += 1; tr[7] += __popc(__activemask()) / G; }
        DRE_TR(0);

        // ---- LM state machine: one damped solve + up to two trials per loop iteration (lev_marq...py:43-134) ----
        if (need_cost0) {
            const Se2h Tp = lane_up(g, Tk1);
            double bv[4] = {0.0, 0.0, 0.0, 0.0};
            const double c = act ? coop_lane_cost(C, k == 0 ? T0 : Tp, Tk1, u, Mk, 1.0 / s, bv) : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) M[bv_cur + i] = bv[i];
            curr = g.sum_ordered(c, K);
            need_cost0 = false; need_build = true;
        }
        DRE_TR(1);
        if (need_build) {
            ++iter; prev = curr; nb = 0; need_build = false;
            const Se2h Tp = lane_up(g, Tk1);
            const double bb = coop_lane_build(g, C, k == 0 ? T0 : Tp, Tk1, u, Mk, s, act, M, bv_cur);
            gn = sqrt(g.sum_ordered(act ? bb : 0.0, K));
        }
        DRE_TR(2);
        double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        ++lm_solves;
        bool success = false;
        double proposed = 0.0;
        Se2h Tt = Tk1;
        double ut[2] = {u[0], u[1]};
        const bool solved = coop_solve(g, K, M, R, lambda, d);
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
                allowed[c] = (dp != 0.0 && act) ? maxp / dp : INFINITY;   // IEEE: the 1e-7 margins of the box are part of the semantics
            }
            // a quotient >= 1 leaves the running scale untouched, so the walk only happens when some control would leave its box
            double scale = 1.0;
            if (g.any(allowed[0] < 1.0 || allowed[1] < 1.0)) {
#pragma unroll 1
                for (int kk = 0; kk < 2 * K; ++kk) {
                    const double al = g.bcast((kk & 1) ? allowed[1] : allowed[0], kk >> 1);
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
            double lin = 0.0, quad = 0.0;
#pragma unroll 1
            for (int bt = 0; bt < 2 && scale != 0.0; ++bt) {
                if (bt > 0) {
#pragma unroll
                    for (int a = 0; a < 5; ++a) d[a] *= 0.8;
                }
                ut[0] = u[0] + d[0]; ut[1] = u[1] + d[1];
                Tt = se2h_step(d[2], d[3], d[4], Tk1);
                const Se2h Tp = lane_up(g, Tt);
                double bv[4] = {0.0, 0.0, 0.0, 0.0};
                const double c = act ? coop_lane_cost(C, k == 0 ? T0 : Tp, Tt, ut, Mk, 1.0 / s, bv) : 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) M[(bv_cur ^ (CS_BV ^ (CS_BV + 4))) + i] = bv[i];
                proposed = g.sum_ordered(c, K);
                ++cost_evals;
                // predicted reduction b^T p - 1/2 p^T A p with the undamped A (lev_marq...py:167-171). The second trial's
                // step is 0.8 x the first one's, so its linear and quadratic forms are 0.8 x and 0.64 x the first trial's
                // (equal up to the rounding of the products; the forms only feed the ratio > 0.25 test)
                if (bt == 0) {
                    double dpv[3];
                    g.template up_n<3>(d + 2, dpv);
                    const double p2 = dpv[0], p3 = dpv[1], p4 = dpv[2];
                    double gg = 0.0, qq = 0.0;
#pragma unroll
                    for (int a = 0; a < 5; ++a) {
                        gg += M[CS_B + a] * d[a];
                        double r = 0.0;
#pragma unroll
                        for (int c2 = 0; c2 < 5; ++c2) r += M[CS_D + (a >= c2 ? a * (a + 1) / 2 + c2 : c2 * (c2 + 1) / 2 + a)] * d[c2];
                        if (k > 0) r += 2.0 * (M[CS_O + 3 * a] * p2 + M[CS_O + 3 * a + 1] * p3 + M[CS_O + 3 * a + 2] * p4);
                        qq += d[a] * r;
                    }
                    lin = g.sum_ordered(act ? gg : 0.0, K);
                    quad = g.sum_ordered(act ? qq : 0.0, K);
                } else { lin *= 0.8; quad *= 0.64; }
                const double pred = lin - 0.5 * quad;
                const double ratio = (pred == 0.0) ? 0.0 : (prev - proposed) * rcp_fast(pred);
                if (ratio > 0.25) { success = true; break; }
            }
        }
        DRE_TR(4);
        if (success) {
            Tk1 = Tt; u[0] = ut[0]; u[1] = ut[1];
            bv_cur ^= CS_BV ^ (CS_BV + 4);
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
                    Tk1.T.c = M[CS_TI]; Tk1.T.s = M[CS_TI + 1]; Tk1.T.x = M[CS_TI + 2]; Tk1.T.y = M[CS_TI + 3]; Tk1.th = M[CS_TI + 4];
                    u[0] = M[CS_UI]; u[1] = M[CS_UI + 1];
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
            const Se2h Tprev_final = lane_up(g, Tk1);
            if (act) {
                double *ctl = B.controls + (size_t)n * B.controls_stride;
                if (2 * k < B.controls_count) ctl[2 * k] = u[0];
                if (2 * k + 1 < B.controls_count) ctl[2 * k + 1] = u[1];
                double4 *wT = B.warm_T + (size_t)n * K;
                double2 *wu = B.warm_u + (size_t)n * K;
                if (k + 1 < K) wT[k] = make_double4(Tk1.T.c, Tk1.T.s, Tk1.T.x, Tk1.T.y);           // T_{k+1} -> slot k
                else {
                    const Se2 TL = (K >= 2) ? Tprev_final.T : T0.T;                                 // duplicate T_{K-1}
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
