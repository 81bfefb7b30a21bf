// orca_device.cuh -- group-cooperative ORCA solve (fp32), one lane-group of G lanes per agent.
//
// Computes what RVO2's Agent::computeNeighbors + computeNewVelocity compute for agent 0 of the
// private simulator that reference scripts/policy/orca.py:64-117 builds per human (neighbour list
// sorted by squared distance, one half-plane per neighbour, LP2 with LP3 fallback), in the same
// fp32 operation order, but with the G lanes of a group working on the neighbours / lines in
// parallel: distances, ranking and half-plane construction are per-lane; LP1's scan over earlier
// lines is a min/max reduction (REDUX, or done redundantly per lane when short); LP2 and LP3 share one
// loop body and find the next violated line with a per-lane scan + one REDUX.MIN.
// This translation unit must be compiled with -fmad=false so every operation rounds once like the
// SSE build of RVO2 (bit-exact parity with the oracle).
#pragma once
#include "dre_common.cuh"

#if defined(__CUDACC__)
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int G>
struct TileGroup {
    cg::thread_block_tile<G> t;
    __device__ explicit TileGroup(cg::thread_block_tile<G> tt) : t(tt) {}
    __device__ int rank() const { return (int)t.thread_rank(); }
    static constexpr int size = G;
    __device__ unsigned ballot(bool p) const { return t.ballot(p); }
    __device__ bool any(bool p) const { return t.any(p); }
    __device__ float shfl_xor(float v, int m) const { return t.shfl_xor(v, m); }
    __device__ int shfl_xor(int v, int m) const { return t.shfl_xor(v, m); }
    __device__ unsigned shfl_xor(unsigned v, int m) const { return t.shfl_xor(v, m); }
    __device__ int bcast(int v, int src) const { return t.shfl(v, src); }
    __device__ void sync() const { t.sync(); }
    __device__ unsigned lane_mask() const
    {
        const unsigned wl = threadIdx.x & 31u;
        return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(unsigned)(G - 1)));
    }
    // group-wide reductions in one REDUX instruction (sm_100a has the fp32 min/max forms; min/max are exact, so the
    // result equals the sequential scan bit for bit)
    // Group-wide min / max / sum: log2(G) butterfly shuffles (min and max are exact, so linearProgram1's reductions equal the
    // sequential scan bit for bit). REDUX with a sub-warp member mask takes a slow path when the warp's groups are on different
    // iterations -- 9 % of ha_step's stall samples sat on REDUX.MIN, half of them branch_resolving; measured on B200: shuffles
    // take the dense-window step from 0.888 to 0.854 ms for the integer ones, the fp32 ones (redux.sync.{min,max}.f32 on
    // sm_100a) another 0.4 %.
    __device__ int min_int(int v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) { const int w = t.shfl_xor(v, o); v = w < v ? w : v; }
        return v;
    }
    __device__ int sum_int(int v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += t.shfl_xor(v, o);
        return v;
    }
    __device__ float min_f32(float v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, t.shfl_xor(v, o));
        return v;
    }
    __device__ float max_f32(float v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, t.shfl_xor(v, o));
        return v;
    }
};
#endif

// single-lane group: lets the same code run on the host (tests) or one thread per agent
struct SerialGroup {
    DRE_HD int rank() const { return 0; }
    static constexpr int size = 1;
    DRE_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
    DRE_HD bool any(bool p) const { return p; }
    DRE_HD float shfl_xor(float v, int) const { return v; }
    DRE_HD unsigned shfl_xor(unsigned v, int) const { return v; }
    DRE_HD int bcast(int v, int) const { return v; }
    DRE_HD void sync() const {}
    DRE_HD int min_int(int v) const { return v; }
    DRE_HD int sum_int(int v) const { return v; }
    DRE_HD float min_f32(float v) const { return v; }
    DRE_HD float max_f32(float v) const { return v; }
};

// 1: "first violated line" by one vote per chunk of G lines with an early exit, instead of every lane testing all its lines
// + a butterfly min. Measured on B200 (16384 x 20): the dense post-reset window gets 2.0 % faster (0.817 -> 0.800 ms), the
// sparse crowd 1.7 % slower (ha_step 0.482 -> 0.490 ms; the vote is loop-carried, a full scan without a hit is the common
// case there); votes only inside linearProgram3: -0.7 % / +1.1 %. Off: the steady state counts more than the first 25 steps.
#ifndef DRE_SCAN_BALLOT
#define DRE_SCAN_BALLOT 0
#endif
#ifndef DRE_LP1_SERIAL
#define DRE_LP1_SERIAL 1   // LP1 scans over at most this many earlier lines run redundantly on every lane (no group reduction).
                           // Measured on B200 at 16384 x 20: 4 -> 1 takes ha_step from 0.607 to 0.592 ms (sparse crowd) and the step
                           // of the dense first 25 steps after a reset from 1.091 to 1.057 ms; 0 and 2 are in between, 8 is slower
#endif
struct DRE_ALIGN16 OLine { float px, py, dx, dy; };   // point, direction: 16 B, 16 B-aligned so that a line moves as ONE 128-bit
                                                      // shared-memory access (a quarter-warp = one lane-group of 8 then touches 128 contiguous bytes: no bank conflicts;
                                                      // as two 64-bit halves the groups of a warp collided: 44 M excess wavefronts per launch, ncu r2)

// Scratch accessors. The lane-group kernels give every group contiguous arrays (plain pointers); the thread-per-agent
// kernel interleaves its threads' arrays (element i of thread t at base[i * stride + t]: conflict-free 16 B accesses).
DRE_HD float4 dre_ld4(const float *p, int k) { return *reinterpret_cast<const float4 *>(p + k); }
struct StridedLines {
    OLine *p; int stride;
    DRE_HD OLine &operator[](int i) const { return p[(size_t)i * stride]; }
    DRE_HD StridedLines operator+(int n) const { StridedLines r = {p + (size_t)n * stride, stride}; return r; }
};
// lines [0, split) in one strided array (shared memory), [split, ...) in another (global scratch)
struct TwoSpaceLines {
    OLine *a; int sa; OLine *b; int sb; int split; int off;
    DRE_HD OLine &operator[](int i) const { const int k = i + off; return k < split ? a[(size_t)k * sa] : b[(size_t)(k - split) * sb]; }
    DRE_HD TwoSpaceLines operator+(int n) const { TwoSpaceLines r = *this; r.off += n; return r; }
};
struct StridedDist {   // float4 chunks interleaved across threads
    float4 *p; int stride;
    DRE_HD float &operator[](int j) const { return reinterpret_cast<float *>(p + (size_t)(j >> 2) * stride)[j & 3]; }
};
DRE_HD float4 dre_ld4(const StridedDist &d, int k) { return d.p[(size_t)(k >> 2) * d.stride]; }

DRE_HD int dre_ffs(unsigned b)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)b);
#else
    return __builtin_ffs((int)b);
#endif
}
DRE_HD int dre_popc(unsigned b)
{
#if defined(__CUDA_ARCH__)
    return __popc(b);
#else
    return __builtin_popcount(b);
#endif
}

DRE_HD unsigned dre_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; return c.u;
#endif
}

// first line index i in [cursor, m) with det(dir_i, point_i - res) > thresh, else m.
// Every lane tests all of its lines (independent loads, no loop-carried vote), then one group-wide min.
template <class Grp, class LinesT>
DRE_HD int orca_first_violated(const Grp &g, const LinesT &L, int m, int cursor, float rx, float ry, float thresh)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
#if DRE_SCAN_BALLOT
    // one vote per chunk of G lines, leaving at the first chunk that holds a violated line
#pragma unroll 1
    for (int base = cursor; base < m; base += G) {
        const int i = base + lane;
        bool v = false;
        if (i < m) {
            const OLine l = L[i];
            v = (l.dx * (l.py - ry) - l.dy * (l.px - rx)) > thresh;
        }
        const unsigned b = g.ballot(v);
        if (b) return base + dre_ffs(b) - 1;
    }
    return m;
#endif
    int best = m;
#pragma unroll 1
    for (int i = cursor + lane; i < m; i += G) {       // from the cursor, not from its 8-aligned base (no wasted tests, -1.7 %);
                                                       // a lane's indices grow, so its first hit is its smallest
        const OLine l = L[i];
        const bool v = (l.dx * (l.py - ry) - l.dy * (l.px - rx)) > thresh;
        if (v && best == m) best = i;
    }
    return g.min_int(best);
}

// RVO2 linearProgram1 on pivot line i; lanes share the scan over lines j < i
template <class Grp, class LinesT>
DRE_HD bool orca_lp1(const Grp &g, const LinesT &L, int i, float radius, float ox, float oy, bool dirOpt, float &rx, float &ry)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    const OLine li = L[i];
    const float dotProduct = li.px * li.dx + li.py * li.dy;
    const float disc = dotProduct * dotProduct + radius * radius - (li.px * li.px + li.py * li.py);
    if (disc < 0.0f) return false;
    const float sq = sqrtf(disc);
    float tL = -dotProduct - sq;
    float tR = -dotProduct + sq;
    // The scan over the earlier lines is a min/max reduction (tL only grows, tR only shrinks, so RVO2's early exits
    // change nothing); an infeasible parallel line is folded in as tL = +inf. Short scans -- most LP1 calls sit on one
    // of the first lines -- are done redundantly by every lane: a group-wide reduction inside a diverged warp costs
    // more than four line tests.
    const bool serial = G == 1 || i <= DRE_LP1_SERIAL;
#pragma unroll 1
    for (int j = serial ? 0 : lane; j < i; j += serial ? 1 : G) {
        const OLine lj = L[j];
        const float den = li.dx * lj.dy - li.dy * lj.dx;
        const float num = lj.dx * (li.py - lj.py) - lj.dy * (li.px - lj.px);
        if (fabsf(den) <= DRE_RVO_EPSILON) {
            if (num < 0.0f) tL = INFINITY;
        } else {
            const float t = num / den;
            if (den >= 0.0f) tR = fminf(tR, t);
            else tL = fmaxf(tL, t);
        }
    }
    if (!serial) {
        tR = g.min_f32(tR);
        tL = g.max_f32(tL);
    }
    if (tL > tR) return false;
    if (dirOpt) {
        if (ox * li.dx + oy * li.dy > 0.0f) { rx = li.px + tR * li.dx; ry = li.py + tR * li.dy; }
        else { rx = li.px + tL * li.dx; ry = li.py + tL * li.dy; }
    } else {
        const float t = li.dx * (ox - li.px) + li.dy * (oy - li.py);
        if (t < tL) { rx = li.px + tL * li.dx; ry = li.py + tL * li.dy; }
        else if (t > tR) { rx = li.px + tR * li.dx; ry = li.py + tR * li.dy; }
        else { rx = li.px + t * li.dx; ry = li.py + t * li.dy; }
    }
    return true;
}

// One ORCA half-plane (RVO2 Agent::computeNewVelocity, agent-agent branch)
DRE_HD OLine orca_make_line(float px, float py, float vx, float vy, float rs, float opx, float opy, float ovx, float ovy,
                            float orad, float invTH, float invDt, unsigned &mask)
{
    const float rpx = opx - px, rpy = opy - py;
    const float rvx = vx - ovx, rvy = vy - ovy;
    const float distSq = rpx * rpx + rpy * rpy;
    const float cr = rs + orad;
    const float crSq = cr * cr;
    OLine line;
    float ux, uy;
    if (distSq > crSq) {
        const float wx = rvx - invTH * rpx, wy = rvy - invTH * rpy;
        const float wLenSq = wx * wx + wy * wy;
        const float dp1 = wx * rpx + wy * rpy;
        if (dp1 < 0.0f && dp1 * dp1 > crSq * wLenSq) {
            const float wLen = sqrtf(wLenSq);
            const float inv = 1.0f / wLen;
            const float uwx = wx * inv, uwy = wy * inv;
            line.dx = uwy; line.dy = -uwx;
            const float s = cr * invTH - wLen;
            ux = s * uwx; uy = s * uwy;
            mask |= 1u;
        } else {
            const float leg = sqrtf(distSq - crSq);
            const float inv = 1.0f / distSq;
            if (rpx * wy - rpy * wx > 0.0f) {
                line.dx = (rpx * leg - rpy * cr) * inv; line.dy = (rpx * cr + rpy * leg) * inv;
                mask |= 2u;
            } else {
                line.dx = -((rpx * leg + rpy * cr) * inv); line.dy = -((-rpx * cr + rpy * leg) * inv);
                mask |= 4u;
            }
            const float dp2 = rvx * line.dx + rvy * line.dy;
            ux = dp2 * line.dx - rvx; uy = dp2 * line.dy - rvy;
        }
    } else {
        const float wx = rvx - invDt * rpx, wy = rvy - invDt * rpy;
        const float wLen = sqrtf(wx * wx + wy * wy);
        const float inv = 1.0f / wLen;
        const float uwx = wx * inv, uwy = wy * inv;
        line.dx = uwy; line.dy = -uwx;
        const float s = cr * invDt - wLen;
        ux = s * uwx; uy = s * uwy;
        mask |= 8u;
    }
    line.px = vx + 0.5f * ux; line.py = vy + 0.5f * uy;
    return line;
}

// Neighbour ranking + ORCA lines of agent `self` into L[0..m) (sorted like RVO2's neighbour list). Returns m.
// Variant for ranges that cover (nearly) the whole crowd -- the reference's neighbor_dist = 10: every candidate is ranked
// against all keys, out-of-range ones carry +inf. See orca_build_lines_cutoff for short ranges.
template <class Grp, class DistT, class LinesT>
DRE_HD int orca_build_lines_all(const Grp &g, int nA, const float *ax, const float *ay, const float *avx, const float *avy,
                            const float *ar, int self, float neighborDist, int maxNeighbors, float invTH, float invDt,
                            DistT sdist, LinesT L, unsigned &mask)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    const float px = ax[self], py = ay[self], vx = avx[self], vy = avy[self], rs = ar[self];
    const float rangeSq = neighborDist * neighborDist;
    const int nA4 = (nA + 3) & ~3;
    int cnt = 0;
    for (int j = lane; j < nA4; j += G) {
        float d = INFINITY;
        if (j < nA) {
            const float dx = px - ax[j], dy = py - ay[j];
            const float dd = dx * dx + dy * dy;
            if (j != self && dd < rangeSq) { d = dd; ++cnt; }
        }
        sdist[j] = d;
    }
    cnt = g.sum_int(cnt);   // also orders the sdist stores before the ranking reads
    g.sync();
    // rank of neighbour j = #{k : (d_k, k) < (d_j, j)} (RVO2 insertNeighbor: ascending distSq, ties by visit order).
    // Squared distances are >= +0, so their bit patterns order like the values: integer compares, 4 keys per load.
    for (int j = lane; j < nA; j += G) {
        const bool in = sdist[j] < INFINITY;
        if (in) {
            // (d_k, k) < (d_j, j)  <=>  k < j ? bits_k <= bits_j : bits_k < bits_j. One compare per key: chunks of four
            // keys below j's own chunk count `<= bits_j` (= `< bits_j + 1`), the others `< bits_j`; the (rare) exact ties
            // with a smaller index inside j's own chunk are added afterwards from that one chunk.
            const unsigned uj = dre_f2u(sdist[j]);
            const int jc = j & ~3;
            int r = 0;
#pragma unroll 2
            for (int k = 0; k < nA4; k += 4) {
                const float4 q = dre_ld4(sdist, k);
                const unsigned thr = uj + (k < jc ? 1u : 0u);
                r += (int)(dre_f2u(q.x) < thr) + (int)(dre_f2u(q.y) < thr) + (int)(dre_f2u(q.z) < thr) + (int)(dre_f2u(q.w) < thr);
            }
            {
                const float4 q = dre_ld4(sdist, jc);
                const int jr = j - jc;   // 0..3: position of j in its chunk
                r += (int)(jr > 0 && dre_f2u(q.x) == uj) + (int)(jr > 1 && dre_f2u(q.y) == uj) + (int)(jr > 2 && dre_f2u(q.z) == uj);
            }
            if (r < maxNeighbors) L[r] = orca_make_line(px, py, vx, vy, rs, ax[j], ay[j], avx[j], avy[j], ar[j], invTH, invDt, mask);
        }
    }
    const int m = cnt < maxNeighbors ? cnt : maxNeighbors;
    g.sync();

    return m;
}

// Neighbour search + ORCA lines of agent `self` into L[0..m) (sorted like RVO2's neighbour list). Returns m.
//
// RVO2 finds the neighbours with a kd-tree range query (Agent::computeNeighbors); an env holds at most 65 agents, all of
// them in one shared-memory tile, so the query here is a CUT-OFF + COMPACTION instead of a tree or a uniform grid: every
// lane tests its candidates against neighborDist^2 (strict <, like RVO2), a ballot + prefix count packs the in-range
// ones -- squared distance into sdist[0..cnt), agent index into sidx[0..cnt) -- and only those are ranked. With the
// reference's neighbor_dist = 10 everything is in range and the compaction only costs (measured on B200: ha_step 0.608 ->
// 0.664 ms at 16384 x 20, the ORCA-only step of 8192 x 50 1.53 -> 1.75 ms), so the kernels are instantiated for both and
// the host picks: cut-off when neighbor_dist is smaller than the crowd's radius (BASELINE config 4's neighbor_dist = 3:
// 6 of 50 in range, 0.875 -> 0.790 ms), the full ranking otherwise.
template <class Grp, class DistT, class LinesT>
DRE_HD int orca_build_lines_cutoff(const Grp &g, int nA, const float *ax, const float *ay, const float *avx, const float *avy,
                            const float *ar, int self, float neighborDist, int maxNeighbors, float invTH, float invDt,
                            DistT sdist, DistT sidx, LinesT L, unsigned &mask)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    const float px = ax[self], py = ay[self], vx = avx[self], vy = avy[self], rs = ar[self];
    const float rangeSq = neighborDist * neighborDist;
    int cnt = 0;
    for (int base = 0; base < nA; base += G) {
        const int j = base + lane;
        float d = 0.0f;
        bool in = false;
        if (j < nA) {
            const float dx = px - ax[j], dy = py - ay[j];
            d = dx * dx + dy * dy;
            in = j != self && d < rangeSq;
        }
        const unsigned b = g.ballot(in);
        if (in) {
            const int pos = cnt + dre_popc(b & ((1u << lane) - 1u));
            sdist[pos] = d;
            sidx[pos] = (float)j;            // small integers are exact in fp32
        }
        cnt += dre_popc(b);
    }
    const int cnt4 = (cnt + 3) & ~3;
    for (int k = cnt + lane; k < cnt4; k += G) sdist[k] = INFINITY;   // pad the last chunk of four keys
    g.sync();
    // rank of neighbour c = #{k : (d_k, k) < (d_c, c)} (RVO2 insertNeighbor: ascending distSq, ties by visit order; the
    // compaction keeps the agents in index order, so the compact position orders ties like the agent index does).
    // Squared distances are >= +0, so their bit patterns order like the values: integer compares, 4 keys per load.
    for (int c = lane; c < cnt; c += G) {
        // (d_k, k) < (d_c, c)  <=>  k < c ? bits_k <= bits_c : bits_k < bits_c. One compare per key: chunks of four keys
        // below c's own chunk count `<= bits_c` (= `< bits_c + 1`), the others `< bits_c`; the (rare) exact ties with a
        // smaller position inside c's own chunk are added afterwards from that one chunk.
        const unsigned uc = dre_f2u(sdist[c]);
        const int cc = c & ~3;
        int r = 0;
#pragma unroll 2
        for (int k = 0; k < cnt4; k += 4) {
            const float4 q = dre_ld4(sdist, k);
            const unsigned thr = uc + (k < cc ? 1u : 0u);
            r += (int)(dre_f2u(q.x) < thr) + (int)(dre_f2u(q.y) < thr) + (int)(dre_f2u(q.z) < thr) + (int)(dre_f2u(q.w) < thr);
        }
        {
            const float4 q = dre_ld4(sdist, cc);
            const int cr = c - cc;   // 0..3: position of c in its chunk
            r += (int)(cr > 0 && dre_f2u(q.x) == uc) + (int)(cr > 1 && dre_f2u(q.y) == uc) + (int)(cr > 2 && dre_f2u(q.z) == uc);
        }
        if (r < maxNeighbors) {
            const int j = (int)sidx[c];
            L[r] = orca_make_line(px, py, vx, vy, rs, ax[j], ay[j], avx[j], avy[j], ar[j], invTH, invDt, mask);
        }
    }
    const int m = cnt < maxNeighbors ? cnt : maxNeighbors;
    g.sync();

    return m;
}

// RVO2 linearProgram3's projection of the lines before line i onto line i (li3 = L[i]) into P[0..pc). Returns pc.
template <class Grp, class LinesT>
DRE_HD int orca_project_lines(const Grp &g, const LinesT &L, int i, const OLine &li3, const LinesT &P)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    int pc = 0;
    for (int base = 0; base < i; base += G) {
        const int j = base + lane;
        bool keep = false;
        OLine ln;
        ln.px = ln.py = ln.dx = ln.dy = 0.0f;
        if (j < i) {
            const OLine lj = L[j];
            const float determinant = li3.dx * lj.dy - li3.dy * lj.dx;
            keep = true;
            if (fabsf(determinant) <= DRE_RVO_EPSILON) {
                if (li3.dx * lj.dx + li3.dy * lj.dy > 0.0f) keep = false;
                else { ln.px = 0.5f * (li3.px + lj.px); ln.py = 0.5f * (li3.py + lj.py); }
            } else {
                const float t = (lj.dx * (li3.py - lj.py) - lj.dy * (li3.px - lj.px)) / determinant;
                ln.px = li3.px + t * li3.dx; ln.py = li3.py + t * li3.dy;
            }
            if (keep) {
                const float ex = lj.dx - li3.dx, ey = lj.dy - li3.dy;
                const float inv = 1.0f / sqrtf(ex * ex + ey * ey);
                ln.dx = ex * inv; ln.dy = ey * inv;
            }
        }
        const unsigned b = g.ballot(keep);
        if (keep) P[pc + dre_popc(b & ((1u << lane) - 1u))] = ln;
        pc += dre_popc(b);
    }
    return pc;
}

// Full solve for agent `self` of a tile of nA agents held in (shared) arrays ax..ar. The other agents are the
// tile's agents in index order skipping `self` (crowd_sim.py:274-280: other humans in list order, robot last).
//   sdist : scratch float[round_up(nA,4)], 16 B aligned, private to the group (in-range squared distances, compacted)
//   sidx  : scratch float[round_up(nA,4)] private to the group (their agent indices)
//   L     : scratch OLine[2*nA] private to the group (ORCA lines, then LP3's projected lines)
//
// RVO2's linearProgram2 / linearProgram3 are driven from ONE loop with a single LP2 body: stage 0 runs LP2 on the ORCA
// lines with the preferred velocity; if it fails at line `fail`, every later violated line i re-enters the same body
// with the lines projected onto line i and the direction objective (linearProgram3). One copy of LP2/LP1 keeps the
// solve inside the instruction cache (the two-copy version stalled 23 % of its cycles on instruction fetch).
template <bool CUTOFF = false, class Grp, class DistT, class LinesT>
DRE_HD int orca_solve_group(const Grp &g, int nA, const float *ax, const float *ay, const float *avx, const float *avy,
                             const float *ar, int self, float prefx, float prefy, float maxSpeed, float neighborDist,
                             int maxNeighbors, float invTH, float invDt, DistT sdist, DistT sidx, LinesT L, float &outx, float &outy,
                             dre_orca_stats *st, bool debug_skip_lp3 = false)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    unsigned mask = 0;
    const int m = CUTOFF ? orca_build_lines_cutoff(g, nA, ax, ay, avx, avy, ar, self, neighborDist, maxNeighbors, invTH, invDt, sdist, sidx, L, mask)
                         : orca_build_lines_all(g, nA, ax, ay, avx, avy, ar, self, neighborDist, maxNeighbors, invTH, invDt, sdist, L, mask);

    LinesT P = L + nA;                       // LP3's projected lines
    float rx = 0.0f, ry = 0.0f;
    int n_lp1 = 0, lp1_work = 0, nproj = 0, fail = m;
    // arguments of the current LP2 run
    LinesT cl = L;
    int cm = m;
    float cox = prefx, coy = prefy;
    bool dirOpt = false;
    // LP3 state
    bool stage0 = true;
    int i3 = 0;
    float distance = 0.0f, tx = 0.0f, ty = 0.0f;
    OLine li3;
    li3.px = li3.py = li3.dx = li3.dy = 0.0f;
    for (;;) {
        // ---- linearProgram2 on (cl, cm) ----
        if (dirOpt) { rx = cox * maxSpeed; ry = coy * maxSpeed; }
        else if (cox * cox + coy * coy > maxSpeed * maxSpeed) {
            const float inv = 1.0f / sqrtf(cox * cox + coy * coy);
            rx = (cox * inv) * maxSpeed; ry = (coy * inv) * maxSpeed;
        } else { rx = cox; ry = coy; }
        int failed_at = cm;
        for (int cursor = 0;;) {
            const int i = orca_first_violated(g, cl, cm, cursor, rx, ry, 0.0f);
            if (i >= cm) break;
            const float sx = rx, sy = ry;
            if (stage0) { n_lp1 += 1; lp1_work += i; }
            if (!orca_lp1(g, cl, i, maxSpeed, cox, coy, dirOpt, rx, ry)) { rx = sx; ry = sy; failed_at = i; break; }
            cursor = i + 1;
        }
        // ---- linearProgram3 driver ----
        if (stage0) {
            stage0 = false;
            fail = failed_at;
            if (fail >= m || debug_skip_lp3) break;
            i3 = fail;
        } else {
            if (failed_at < cm) { rx = tx; ry = ty; }
            distance = li3.dx * (li3.py - ry) - li3.dy * (li3.px - rx);
            g.sync();   // P is rewritten below
            i3 = i3 + 1;
        }
        const int i = orca_first_violated(g, L, m, i3, rx, ry, distance);
        if (i >= m) break;
        i3 = i;
        li3 = L[i];
        const int pc = orca_project_lines(g, L, i, li3, P);
        g.sync();
        tx = rx; ty = ry;
        ++nproj;
        cl = P; cm = pc; cox = -li3.dy; coy = li3.dx; dirOpt = true;
    }
    outx = rx; outy = ry;
    // cost class of this solve (0..7) for the queue order of the next pass: 0..3 = linearProgram2 sufficed, by the number of
    // linearProgram1 calls (0-1, 2-3, 4-5, 6+); 4..7 = linearProgram3 ran, by its number of projections (1, 2, 3, 4+).
    // (Measured on B200, dense post-reset window of 16384 x 20: 2 classes 0.945 ms, 4 classes 0.926, these 8 0.903; a
    // 16-class running work estimate updated inside the LP loops 0.927: the bookkeeping costs more than it orders.)
    const int cost_class = fail < m ? 4 + (nproj > 4 ? 3 : nproj - 1) : (n_lp1 > 7 ? 3 : n_lp1 >> 1);
    if (st) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mask |= g.shfl_xor(mask, o);
        if (lane == 0) {
            st->n_lines = m; st->lp2_fail = fail; st->n_lp1 = n_lp1; st->lp1_work = lp1_work;
            st->branch_mask = mask; st->n_lp3_proj = nproj;
        }
    }
    return cost_class;
}
