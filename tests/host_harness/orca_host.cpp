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
