// tests/host_harness/orca_host.cpp -- TEST INFRASTRUCTURE. Compiles the PRODUCT's ORCA solve (dr-mpc_b200/csrc/orca_device.cuh:
// orca_solve_group and everything below it -- neighbour ranking, half-plane construction, the single LP2 / LP3 loop, LP1) for the
// host with the one-lane SerialGroup, so the CPU-only suite can run the kernel's own arithmetic on the reference's golden crowds.
// The tile set-up mirrors orca_thread_kernel / ha_step_kernel (fp64 -> fp32 at the Cython boundary, robot last with zero
// velocity, preferred velocity in fp64: scripts/policy/orca.py:84-111). Never shipped. Build with -ffp-contract=off.
#include <cuda_runtime.h>
#include <vector>
#include "../../dr-mpc_b200/csrc/orca_device.cuh"

template <bool CUTOFF>
static void crowd_pass(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal, const double *rpos,
                       const uint8_t *is_static, float *vnew, dre_orca_stats *stats)
{
    const int nA = Nh + (p.robot_visible ? 1 : 0), nA4 = (nA + 3) & ~3;
    std::vector<float> tile((size_t)5 * nA);
    std::vector<OLine> lines((size_t)2 * nA + 1);
    float *sd = nullptr, *si = nullptr;
    // 16-byte aligned scratch for the float4 loads of the ranking
    std::vector<float4> dist4((size_t)2 * (nA4 / 4) + 1);
    sd = reinterpret_cast<float *>(dist4.data());
    si = sd + nA4;
    const float invTH = 1.0f / p.time_horizon, invDt = 1.0f / p.time_step;
    const int maxNb = p.max_neighbors > 0 ? p.max_neighbors : nA - 1;
    SerialGroup g;
    for (int e = 0; e < E; ++e) {
        float *A = tile.data();
        for (int a = 0; a < nA; ++a) {
            float x, y, vx = 0.f, vy = 0.f;
            if (a < Nh) {
                x = (float)hpos[((size_t)e * Nh + a) * 2]; y = (float)hpos[((size_t)e * Nh + a) * 2 + 1];
                vx = hvel[((size_t)e * Nh + a) * 2]; vy = hvel[((size_t)e * Nh + a) * 2 + 1];
            } else { x = (float)rpos[(size_t)e * 2]; y = (float)rpos[(size_t)e * 2 + 1]; }
            A[a] = x; A[nA + a] = y; A[2 * nA + a] = vx; A[3 * nA + a] = vy; A[4 * nA + a] = p.radius;
        }
        for (int i = 0; i < Nh; ++i) {
            const size_t gidx = (size_t)e * Nh + i;
            float ox = 0.f, oy = 0.f;
            if (is_static && is_static[gidx]) {
                if (stats) { dre_orca_stats z = {0, 0, 0, 0, 0u, 0}; stats[gidx] = z; }
            } else {
                const double dx = goal[gidx * 2] - hpos[gidx * 2], dy = goal[gidx * 2 + 1] - hpos[gidx * 2 + 1];
                const double speed = sqrt(dx * dx + dy * dy);
                double pvx = dx, pvy = dy;
                if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
                orca_solve_group<CUTOFF>(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, (float)pvx, (float)pvy, p.max_speed_self,
                                         p.neighbor_dist, maxNb, invTH, invDt, sd, si, lines.data(), ox, oy, stats ? &stats[gidx] : nullptr, false);
            }
            vnew[gidx * 2] = ox; vnew[gidx * 2 + 1] = oy;
        }
    }
}

extern "C" int host_orca_crowd_pass(const dre_orca_params *p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                                    const double *rpos, const uint8_t *is_static, int cutoff, float *vnew, dre_orca_stats *stats)
{
    if (cutoff) crowd_pass<true>(*p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats);
    else crowd_pass<false>(*p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The LANE-PARALLEL form on the host: G OS threads play the G lanes of one lane-group. Every collective of the group
// interface (shuffle, ballot, reductions, sync) is an exchange through a shared slot array between two barriers, i.e. the
// lanes run the kernel's cooperative code path -- per-lane distance tests and rankings, lines stored at their sorted rank,
// lane-parallel violated-line scans and linearProgram1 reductions -- exactly as the 4- / 8- / 16-lane CUDA groups do.
// ---------------------------------------------------------------------------------------------------------------
#include <atomic>
#include <thread>
#include <sched.h>

struct LaneShared {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    unsigned slot[64];
};

template <int G>
struct HostLaneGroup {
    int r;
    LaneShared *S;
    mutable int my_phase = 0;
    static constexpr int size = G;
    int rank() const { return r; }
    void barrier() const
    {
        const int next = my_phase ^ 1;
        if (S->arrived.fetch_add(1, std::memory_order_acq_rel) == G - 1) {
            S->arrived.store(0, std::memory_order_relaxed);
            S->phase.store(next, std::memory_order_release);
        } else {
            for (unsigned spin = 0; S->phase.load(std::memory_order_acquire) != next; ++spin)
                if ((spin & 63u) == 63u) sched_yield();
        }
        my_phase = next;
    }
    unsigned exchange(unsigned v, int src) const
    {
        S->slot[r] = v;
        barrier();
        const unsigned o = S->slot[src];
        barrier();
        return o;
    }
    static unsigned f2u(float f) { union { float f; unsigned u; } c; c.f = f; return c.u; }
    static float u2f(unsigned u) { union { float f; unsigned u; } c; c.u = u; return c.f; }
    unsigned ballot(bool p) const
    {
        S->slot[r] = p ? 1u : 0u;
        barrier();
        unsigned b = 0;
        for (int k = 0; k < G; ++k) b |= S->slot[k] << k;
        barrier();
        return b;
    }
    bool any(bool p) const { return ballot(p) != 0u; }
    float shfl_xor(float v, int m) const { return u2f(exchange(f2u(v), r ^ m)); }
    int shfl_xor(int v, int m) const { return (int)exchange((unsigned)v, r ^ m); }
    unsigned shfl_xor(unsigned v, int m) const { return exchange(v, r ^ m); }
    int bcast(int v, int src) const { return (int)exchange((unsigned)v, src); }
    void sync() const { barrier(); }
    // the same butterflies as TileGroup (orca_device.cuh)
    int min_int(int v) const { for (int o = G / 2; o > 0; o >>= 1) { const int w = shfl_xor(v, o); v = w < v ? w : v; } return v; }
    int sum_int(int v) const { for (int o = G / 2; o > 0; o >>= 1) v += shfl_xor(v, o); return v; }
    float min_f32(float v) const { for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, shfl_xor(v, o)); return v; }
    float max_f32(float v) const { for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, shfl_xor(v, o)); return v; }
};

template <int G, bool CUTOFF>
static void crowd_pass_lanes(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                             const double *rpos, const uint8_t *is_static, float *vnew, dre_orca_stats *stats)
{
    const int nA = Nh + (p.robot_visible ? 1 : 0), nA4 = (nA + 3) & ~3;
    std::vector<float> tile((size_t)5 * nA);
    std::vector<OLine> lines((size_t)2 * nA + 1);
    std::vector<float4> dist4((size_t)2 * (nA4 / 4) + 1);
    float *sd = reinterpret_cast<float *>(dist4.data()), *si = sd + nA4;
    const float invTH = 1.0f / p.time_horizon, invDt = 1.0f / p.time_step;
    const int maxNb = p.max_neighbors > 0 ? p.max_neighbors : nA - 1;
    LaneShared S;
    auto lane_main = [&](int r) {
        HostLaneGroup<G> g;
        g.r = r; g.S = &S;
        float *A = tile.data();
        for (int e = 0; e < E; ++e) {
            g.sync();                                   // every lane is done with the previous env's tile
            for (int a = r; a < nA; a += G) {           // the lanes stage the tile together, like the kernels
                float x, y, vx = 0.f, vy = 0.f;
                if (a < Nh) {
                    x = (float)hpos[((size_t)e * Nh + a) * 2]; y = (float)hpos[((size_t)e * Nh + a) * 2 + 1];
                    vx = hvel[((size_t)e * Nh + a) * 2]; vy = hvel[((size_t)e * Nh + a) * 2 + 1];
                } else { x = (float)rpos[(size_t)e * 2]; y = (float)rpos[(size_t)e * 2 + 1]; }
                A[a] = x; A[nA + a] = y; A[2 * nA + a] = vx; A[3 * nA + a] = vy; A[4 * nA + a] = p.radius;
            }
            g.sync();
            for (int i = 0; i < Nh; ++i) {
                const size_t gidx = (size_t)e * Nh + i;
                float ox = 0.f, oy = 0.f;
                if (is_static && is_static[gidx]) {
                    if (stats && r == 0) { dre_orca_stats z = {0, 0, 0, 0, 0u, 0}; stats[gidx] = z; }
                } else {
                    const double dx = goal[gidx * 2] - hpos[gidx * 2], dy = goal[gidx * 2 + 1] - hpos[gidx * 2 + 1];
                    const double speed = sqrt(dx * dx + dy * dy);
                    double pvx = dx, pvy = dy;
                    if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
                    g.sync();                           // the scratch of the previous solve is free
                    orca_solve_group<CUTOFF>(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, (float)pvx, (float)pvy, p.max_speed_self,
                                             p.neighbor_dist, maxNb, invTH, invDt, sd, si, lines.data(), ox, oy, stats ? &stats[gidx] : nullptr, false);
                }
                if (r == 0) { vnew[gidx * 2] = ox; vnew[gidx * 2 + 1] = oy; }
            }
        }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(lane_main, r);
    lane_main(0);
    for (auto &t : th) t.join();
}

// lanes = 1: the one-lane SerialGroup; 2 / 4 / 8 / 16: that many cooperating lanes (OS threads)
extern "C" int host_orca_crowd_pass_lanes(const dre_orca_params *p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                                          const double *rpos, const uint8_t *is_static, int cutoff, int lanes, float *vnew, dre_orca_stats *stats)
{
#define DRE_LANES_CASE(G) case G: if (cutoff) crowd_pass_lanes<G, true>(*p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats); \
                                  else crowd_pass_lanes<G, false>(*p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats); return 0
    switch (lanes) {
    case 1: return host_orca_crowd_pass(p, E, Nh, hpos, hvel, goal, rpos, is_static, cutoff, vnew, stats);
    DRE_LANES_CASE(2);
    DRE_LANES_CASE(4);
    DRE_LANES_CASE(8);
    DRE_LANES_CASE(16);
    default: return -1;
    }
#undef DRE_LANES_CASE
}
