// orca_device.cuh -- group-cooperative ORCA solve (fp32), one lane-group of G lanes per agent.
//
// Computes what RVO2's Agent::computeNeighbors + computeNewVelocity compute for agent 0 of the
// private simulator that reference scripts/policy/orca.py:64-117 builds per human (neighbour list
// sorted by squared distance, one half-plane per neighbour, LP2 with LP3 fallback), in the same
// fp32 operation order, but with the G lanes of a group working on the neighbours / lines in
// parallel: distances, ranking and half-plane construction are per-lane; LP1's scan over earlier
// lines is a min/max reduction with shuffles; LP2/LP3 find the next violated line with a ballot.
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
    __device__ unsigned shfl_xor(unsigned v, int m) const { return t.shfl_xor(v, m); }
    __device__ int bcast(int v, int src) const { return t.shfl(v, src); }
    __device__ void sync() const { t.sync(); }
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
};

struct OLine { float px, py, dx, dy; };   // point, direction (16 B, float4-compatible)

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

// first line index i in [cursor, m) with det(dir_i, point_i - res) > thresh, else m
template <class Grp>
DRE_HD int orca_first_violated(const Grp &g, const OLine *L, int m, int cursor, float rx, float ry, float thresh)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    for (int base = (cursor / G) * G; base < m; base += G) {
        const int i = base + lane;
        bool v = false;
        if (i >= cursor && i < m) {
            const OLine l = L[i];
            v = (l.dx * (l.py - ry) - l.dy * (l.px - rx)) > thresh;
        }
        const unsigned b = g.ballot(v);
        if (b) return base + dre_ffs(b) - 1;
    }
    return m;
}

// RVO2 linearProgram1 on pivot line i; lanes share the scan over lines j < i
template <class Grp>
DRE_HD bool orca_lp1(const Grp &g, const OLine *L, int i, float radius, float ox, float oy, bool dirOpt, float &rx, float &ry)
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
    bool fail = false;
    for (int j = lane; j < i; j += G) {
        const OLine lj = L[j];
        const float den = li.dx * lj.dy - li.dy * lj.dx;
        const float num = lj.dx * (li.py - lj.py) - lj.dy * (li.px - lj.px);
        if (fabsf(den) <= DRE_RVO_EPSILON) {
            if (num < 0.0f) fail = true;
        } else {
            const float t = num / den;
            if (den >= 0.0f) tR = fminf(tR, t);
            else tL = fmaxf(tL, t);
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        tR = fminf(tR, g.shfl_xor(tR, o));
        tL = fmaxf(tL, g.shfl_xor(tL, o));
    }
    fail = g.any(fail);
    if (fail || tL > tR) return false;
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

// RVO2 linearProgram2; returns the failing line index or m. n_lp1 / lp1_work are optional counters.
template <class Grp>
DRE_HD int orca_lp2(const Grp &g, const OLine *L, int m, float radius, float ox, float oy, bool dirOpt,
                    float &rx, float &ry, int *n_lp1, int *lp1_work)
{
    if (dirOpt) { rx = ox * radius; ry = oy * radius; }
    else if (ox * ox + oy * oy > radius * radius) {
        const float inv = 1.0f / sqrtf(ox * ox + oy * oy);
        rx = (ox * inv) * radius; ry = (oy * inv) * radius;
    } else { rx = ox; ry = oy; }
    int cursor = 0;
    for (;;) {
        const int i = orca_first_violated(g, L, m, cursor, rx, ry, 0.0f);
        if (i >= m) return m;
        const float tx = rx, ty = ry;
        if (n_lp1) { *n_lp1 += 1; *lp1_work += i; }
        if (!orca_lp1(g, L, i, radius, ox, oy, dirOpt, rx, ry)) { rx = tx; ry = ty; return i; }
        cursor = i + 1;
    }
}

// RVO2 linearProgram3 (no obstacle lines). P is scratch for the projected lines (capacity m).
template <class Grp>
DRE_HD int orca_lp3(const Grp &g, const OLine *L, int m, int beginLine, float radius, float &rx, float &ry, OLine *P)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    float distance = 0.0f;
    int cursor = beginLine, nproj = 0;
    for (;;) {
        const int i = orca_first_violated(g, L, m, cursor, rx, ry, distance);
        if (i >= m) break;
        const OLine li = L[i];
        int cnt = 0;
        for (int base = 0; base < i; base += G) {
            const int j = base + lane;
            bool keep = false;
            OLine ln;
            ln.px = ln.py = ln.dx = ln.dy = 0.0f;
            if (j < i) {
                const OLine lj = L[j];
                const float determinant = li.dx * lj.dy - li.dy * lj.dx;
                keep = true;
                if (fabsf(determinant) <= DRE_RVO_EPSILON) {
                    if (li.dx * lj.dx + li.dy * lj.dy > 0.0f) keep = false;
                    else { ln.px = 0.5f * (li.px + lj.px); ln.py = 0.5f * (li.py + lj.py); }
                } else {
                    const float t = (lj.dx * (li.py - lj.py) - lj.dy * (li.px - lj.px)) / determinant;
                    ln.px = li.px + t * li.dx; ln.py = li.py + t * li.dy;
                }
                if (keep) {
                    const float ex = lj.dx - li.dx, ey = lj.dy - li.dy;
                    const float inv = 1.0f / sqrtf(ex * ex + ey * ey);
                    ln.dx = ex * inv; ln.dy = ey * inv;
                }
            }
            const unsigned b = g.ballot(keep);
            if (keep) P[cnt + dre_popc(b & ((1u << lane) - 1u))] = ln;
            cnt += dre_popc(b);
        }
        g.sync();
        const float tx = rx, ty = ry;
        ++nproj;
        if (orca_lp2(g, P, cnt, radius, -li.dy, li.dx, true, rx, ry, (int *)nullptr, (int *)nullptr) < cnt) { rx = tx; ry = ty; }
        distance = li.dx * (li.py - ry) - li.dy * (li.px - rx);
        g.sync();   // P is rewritten by the next outer iteration
        cursor = i + 1;
    }
    return nproj;
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

// Full solve for agent `self` of a tile of nA agents held in (shared) arrays ax..ar. The other agents are the
// tile's agents in index order skipping `self` (crowd_sim.py:274-280: other humans in list order, robot last).
//   sdist : scratch float[nA] private to the group      L : scratch OLine[2*nA] private to the group
template <class Grp>
DRE_HD void orca_solve_group(const Grp &g, int nA, const float *ax, const float *ay, const float *avx, const float *avy,
                             const float *ar, int self, float prefx, float prefy, float maxSpeed, float neighborDist,
                             int maxNeighbors, float invTH, float invDt, float *sdist, OLine *L, float &outx, float &outy,
                             dre_orca_stats *st, bool debug_skip_lp3 = false)
{
    constexpr int G = Grp::size;
    const int lane = g.rank();
    const float px = ax[self], py = ay[self], vx = avx[self], vy = avy[self], rs = ar[self];
    const float rangeSq = neighborDist * neighborDist;
    for (int j = lane; j < nA; j += G) {
        const float dx = px - ax[j], dy = py - ay[j];
        const float d = dx * dx + dy * dy;
        sdist[j] = (j != self && d < rangeSq) ? d : INFINITY;
    }
    g.sync();
    int cnt = 0;
    unsigned mask = 0;
    for (int base = 0; base < nA; base += G) {
        const int j = base + lane;
        const bool in = (j < nA) && (sdist[j] < INFINITY);
        cnt += dre_popc(g.ballot(in));
        if (in) {
            const float dj = sdist[j];
            int r = 0;
            for (int k = 0; k < nA; ++k) {
                const float dk = sdist[k];
                r += (dk < dj || (dk == dj && k < j)) ? 1 : 0;
            }
            if (r < maxNeighbors) L[r] = orca_make_line(px, py, vx, vy, rs, ax[j], ay[j], avx[j], avy[j], ar[j], invTH, invDt, mask);
        }
    }
    const int m = cnt < maxNeighbors ? cnt : maxNeighbors;
    g.sync();
    float rx, ry;
    int n_lp1 = 0, lp1_work = 0, nproj = 0;
    const int fail = orca_lp2(g, L, m, maxSpeed, prefx, prefy, false, rx, ry, &n_lp1, &lp1_work);
    if (fail < m && !debug_skip_lp3) nproj = orca_lp3(g, L, m, fail, maxSpeed, rx, ry, L + nA);
    outx = rx; outy = ry;
    if (st) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mask |= g.shfl_xor(mask, o);
        if (lane == 0) {
            st->n_lines = m; st->lp2_fail = fail; st->n_lp1 = n_lp1; st->lp1_work = lp1_work;
            st->branch_mask = mask; st->n_lp3_proj = nproj;
        }
    }
}
