// env_kernels.cu -- the batched env step around the ORCA pass and the MPC solve.
// Compile with -fmad=false (bit-faithful fp32 ORCA, once-rounded fp64 env arithmetic).
//
// One env step = pt_step_kernel -> ha_step_kernel -> MPC solve. The robot's motion depends only on the action, so the
// robot/path half runs FIRST; the crowd half and the MPC solve both depend on nothing but it. The crowd half, which
// produces 3/4 of the output bytes, goes next so that on the host path its observations travel while the solve runs.
//
//   pt_step_kernel : robot unicycle step (agent.py:169-176) + PTEnv.step (path_tracking_env.py:80-101): 5 intermediate
//                    poses, 2 localisations, PT rewards (reward_comp.py:20-116), optional path switch
//                    (HA_and_PT env.py:130-141), PT state (path_tracking_core.py:94-145).
//   ha_step_kernel : HAEnv.step / warm_start (reference human_avoidance_env.py:128-186, 79-125) for EPB envs per CTA:
//                    ORCA pass 1 (robot at its previous pose) -> human Euler integration -> history roll (+ sensor range)
//                    -> observation (:24-77; final here, it does not need the look-ahead) -> ORCA pass 2 (disturbance
//                    look-ahead, :245) -> collision / safety / disturbance / actuation rewards (:189-291) -> merge with
//                    the PT rewards / dones (HA_and_PT env.py:311-416), episode statistics -> goal resampling
//                    (:181-184, crowd_sim.py:114-144 with Philox).
//                    The env's agents live in a shared-memory fp32 SoA tile for both ORCA passes. For the host path the
//                    last CTA of every chunk of envs raises a flag in mapped host memory, so the D2H copy of a chunk's
//                    observations starts while later CTAs of the same launch are still computing.
//   cvmm_filter_kernel : cvmm_diverse_safety_pipeline / cvmm_safety_check (HA_and_PT env.py:422-513), one warp per env.
//   env_reset_kernel : reset_continuous_task (HA_and_PT env.py:263-280) + HAEnv.reset (:297-338) + spawning
//                    (crowd_sim.py:232-262 with Philox).
#include <cstdlib>
#include "orca_device.cuh"
#include "env_device.cuh"
#include "env_launch.cuh"

// ---------------------------------------------------------------------------------------------------------------
// shared-memory layout of ha_step_kernel
// ---------------------------------------------------------------------------------------------------------------
struct HaLayout {
    size_t off_lines, off_agents, off_dist, off_vnew, off_pos, off_pen, off_flags, off_stats, off_order, off_pref, bytes;
};
static HaLayout ha_layout(int Nh, int nA, int NG, int EPB, bool cutoff)
{
    HaLayout l;
    size_t o = 0;
    auto seg = [&o](size_t bytes) { const size_t at = o; o = (o + bytes + 15) / 16 * 16; return at; };   // 16 B aligned segments
    l.off_lines = seg((size_t)NG * 2 * nA * sizeof(OLine));
    l.off_pos = seg((size_t)EPB * Nh * sizeof(double2));          // new human positions (fp64)
    l.off_pen = seg((size_t)EPB * Nh * 3 * sizeof(double));       // pen_v, pen_th, closest_dist per human
    l.off_agents = seg((size_t)EPB * 5 * nA * sizeof(float));
    l.off_dist = seg((size_t)NG * (cutoff ? 2 : 1) * ((nA + 3) & ~3) * sizeof(float));   // squared distances (+ the agent indices of the compacted ones)
    l.off_vnew = seg((size_t)EPB * Nh * sizeof(float2));
    l.off_flags = seg((size_t)(EPB * 4 + 3) * sizeof(int));   // per-env flags + the two ORCA work counters + the CTA's LP3 count
    l.off_stats = seg(12 * sizeof(unsigned long long));        // CTA-level episode statistics: 9 counters + 3 reward sums
    l.off_order = seg((size_t)EPB * Nh * (sizeof(uint16_t) + 1) + 16 * sizeof(int) + 4);   // agent queue order, solve classes, sort counters
    l.off_pref = seg((size_t)EPB * Nh * (sizeof(float2) + 1));   // preferred velocities (fp32, like the Cython boundary) + static flags
    l.bytes = (o + 15) / 16 * 16;
    return l;
}

template <int G, bool REUSE, bool CUTOFF, bool PRUNE>
__device__ __noinline__ void orca_pass_tile(const EnvDev &D, TileGroup<G> &g, int gi, int NG, int EPB, int env0, int nA,
                               const float *agents, OLine *lines, float *dist, float2 *vout, const double2 *posd,
                               bool pos_from_smem, int *lp3_count, int *work_counter, bool reuse, const int *flags,
                               bool prune, const uint16_t *order, uint8_t *cls_out, const float2 *pref, const uint8_t *sstat)
{
    const int Nh = D.Nh;
    OLine *L = lines + (size_t)gi * 2 * nA;
    const int nA4 = (nA + 3) & ~3;
    float *sd = dist + (size_t)gi * (CUTOFF ? 2 : 1) * nA4;
    const float invTH = 1.0f / D.orca.time_horizon, invDt = 1.0f / D.orca.time_step;
    const int maxNb = D.orca.max_neighbors > 0 ? D.orca.max_neighbors : nA - 1;
    // LP3 (infeasible) solves cost an order of magnitude more than LP2 ones, so the groups of a CTA pull agents from a
    // shared counter instead of a static assignment
    for (;;) {
        int ag = 0;
        if (g.rank() == 0) ag = atomicAdd(work_counter, 1);
        ag = g.bcast(ag, 0);
        if (ag >= EPB * Nh) break;
        ag = order[ag];                 // expensive (LP3) solves first: see ha_build_order
        const int el = ag / Nh, i = ag - el * Nh, e = env0 + el;
        if (flags[el * 4 + 2] == 0) continue;
        const size_t gidx = (size_t)e * Nh + i;
        float ox = 0.f, oy = 0.f;
        if (PRUNE && prune) {
            // look-ahead pass with lookahead_prune: the disturbance reward (human_avoidance_env.py:247-273) reads the future
            // velocity only of humans within its radius of the robot; every other look-ahead solve is dead work (each ORCA
            // solve is a pure function of the crowd state). Same expression as the reward loop below -> same decision.
            const double2 P = posd[el * Nh + i];
            const double ddx = P.x - D.rpose[(size_t)e * 4], ddy = P.y - D.rpose[(size_t)e * 4 + 1];
            const double closest = sqrt(ddx * ddx + ddy * ddy) - D.human_radius - D.robot_radius;
            if (closest > D.dist_radius || closest <= 0.0) { g.sync(); continue; }
        }
        if (REUSE && reuse && D.cache_valid[e] && !D.goal_changed[gidx]) {
            // orca_dedup: this solve was done as the look-ahead pass of the previous step, on this very state
            const float2 c = D.hvel_next[gidx];
            ox = c.x; oy = c.y;
        } else if (!sstat[ag]) {
            const float2 pv = pref[ag];   // computed for the whole tile at once (ha_pref_velocity), not per lane-group here
            const float *A = agents + (size_t)el * 5 * nA;
            const int hard = orca_solve_group<CUTOFF>(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, pv.x, pv.y,
                             D.orca.max_speed_self, D.orca.neighbor_dist, maxNb, invTH, invDt, sd, sd + nA4, L, ox, oy, nullptr);
            if (g.rank() == 0) cls_out[ag] = (uint8_t)hard;
            if (lp3_count && g.rank() == 0 && hard >= 4) atomicAdd(lp3_count, 1);   // shared memory; one global atomic per CTA
        }
        if (g.rank() == 0) vout[el * Nh + i] = make_float2(ox, oy);
        g.sync();
    }
}

// Preferred velocity of a human (orca.py:97-100): towards the goal, normalised only beyond 1 m; computed in fp64 like the
// reference's numpy, then cast to fp32 like its Cython boundary.
__device__ __forceinline__ float2 ha_pref_velocity(double2 P, double2 Gl)
{
    const double dx = Gl.x - P.x, dy = Gl.y - P.y;
    const double speed = sqrt(dx * dx + dy * dy);
    double pvx = dx, pvy = dy;
    if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
    return make_float2((float)pvx, (float)pvy);
}

// Queue order of a pass: agents sorted by the cost class of their previous solve, most expensive first (a solve that needs
// linearProgram3 costs 2-3x an LP2-only one): the cheap ones fill the tail (longest-processing-time-first) and the
// lane-groups of a warp tend to be on the same class of solve at the same time. The class of an agent's previous solve predicts the next one almost always (the look-ahead
// pass of step t-1 and the movement pass of step t see the same crowd). Order only: results do not depend on it.
__device__ __forceinline__ void ha_build_order(uint16_t *order, const uint8_t *cls, int n, int *fill /*[16]*/)
{
    // counting sort by descending class (8 classes): count, prefix, place
    if (threadIdx.x < 16) fill[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) atomicAdd(&fill[cls[t] & 7], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int base = 0;
        for (int c = 7; c >= 0; --c) { fill[8 + c] = base; base += fill[c]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) order[atomicAdd(&fill[8 + (cls[t] & 7)], 1)] = (uint16_t)t;
    __syncthreads();
}

// -DDRE_HA_TRACE: thread 0 of every CTA adds the cycles it spent in each phase to stat_counters[12..15] + stat_sums
// (measurement builds only; see tools/probe_ha_trace.py)
#if defined(DRE_HA_TRACE)
#define DRE_HA_T(slot) do { if (threadIdx.x == 0 && D.stat_counters && mode == 0) { const long long _n = clock64(); \
        atomicAdd(D.stat_counters + (slot), (unsigned long long)(_n - tr_t)); tr_t = _n; } } while (0)
#else
#define DRE_HA_T(slot) do { } while (0)
#endif
#if defined(DRE_HA_TRACE) && DRE_HA_TRACE == 2   // second level: split the last phase instead
#define DRE_HA_T2(slot) do { if (threadIdx.x == 0 && D.stat_counters && mode == 0) { const long long _n = clock64(); \
        atomicAdd(D.stat_counters + (slot), (unsigned long long)(_n - tr_t)); tr_t = _n; } } while (0)
#undef DRE_HA_T
#define DRE_HA_T(slot) do { if (threadIdx.x == 0) tr_t = clock64(); } while (0)
#else
#define DRE_HA_T2(slot) do { } while (0)
#endif

// mode: 0 = full step, 1 = warm-start step (robot frozen, no rewards, no second pass), 2 = observation only
// DEDUP: the orca_dedup variant (reuses the previous step's look-ahead pass); a separate instantiation so that the default
// two-pass kernel carries none of its code
// 128-thread CTAs, register target for 9 resident CTAs per SM (56 registers, 24 B of spills). Measured on B200 at 16384 x 20
// (driver window / sparse-crowd ha_step): 7 CTAs 72 regs 1.047 / 0.587 ms, 8 CTAs 64 regs 1.045 / 0.582, 9 CTAs 56 regs
// 1.016 / 0.572, 10 CTAs 48 regs 1.016 / 0.569 (but the 50-human ORCA-only step 1.58 vs 1.49 ms); the unconstrained
// __launch_bounds__(512) this kernel had before also landed on 64 registers, with spills: 1.057 / 0.592.
#ifndef DRE_HA_MINB
#define DRE_HA_MINB 9
#endif
#define DRE_HA_BOUNDS 128, DRE_HA_MINB
template <int G, bool DEDUP, bool CUTOFF, bool PRUNE>
__global__ void __launch_bounds__(DRE_HA_BOUNDS)
ha_step_kernel(const __grid_constant__ EnvDev D, int EPB, int mode, int write_obs, const __grid_constant__ HaLayout lay)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nh = D.Nh, nA = Nh + (D.orca.robot_visible ? 1 : 0), H = D.LB + 1;
    const int NG = blockDim.x / G;
    OLine *lines = reinterpret_cast<OLine *>(smem_raw + lay.off_lines);
    double2 *posd = reinterpret_cast<double2 *>(smem_raw + lay.off_pos);
    double *pen = reinterpret_cast<double *>(smem_raw + lay.off_pen);
    float *agents = reinterpret_cast<float *>(smem_raw + lay.off_agents);
    float *dist = reinterpret_cast<float *>(smem_raw + lay.off_dist);
    float2 *vnew = reinterpret_cast<float2 *>(smem_raw + lay.off_vnew);
    int *flags = reinterpret_cast<int *>(smem_raw + lay.off_flags);
    float2 *pref = reinterpret_cast<float2 *>(smem_raw + lay.off_pref);
    uint8_t *sstat = reinterpret_cast<uint8_t *>(pref + (size_t)EPB * D.Nh);
    uint16_t *order = reinterpret_cast<uint16_t *>(smem_raw + lay.off_order);
    int *fill = reinterpret_cast<int *>(smem_raw + lay.off_order + (((size_t)EPB * D.Nh * sizeof(uint16_t) + 3) & ~(size_t)3));
    uint8_t *cls = reinterpret_cast<uint8_t *>(fill + 16);
    unsigned long long *s_cnt = reinterpret_cast<unsigned long long *>(smem_raw + lay.off_stats);   // [9]
    double *s_sum = reinterpret_cast<double *>(s_cnt + 9);                                            // [3]
    const int env0 = blockIdx.x * EPB;
    const double dt = D.dt;

    cg::thread_block blk = cg::this_thread_block();
    TileGroup<G> g(cg::tiled_partition<G>(blk));
    const int gi = threadIdx.x / G;
#ifdef DRE_HA_TRACE
    long long tr_t = clock64();
#endif

    // per-env flags: [0] a goal was reached, [1] visible humans (sensor restriction), [2] the env takes part in this launch
    // (in range and selected by the mask of a masked reset / observe), [3] spare; then the two ORCA work counters
    for (int t = threadIdx.x; t < EPB * 4 + 3; t += blockDim.x) {
        int v = 0;
        if (t < EPB * 4 && (t & 3) == 2) { const int e = env0 + (t >> 2); v = e < D.E && (D.env_mask == nullptr || D.env_mask[e] != 0); }
        flags[t] = v;
    }
    if (threadIdx.x < 12) s_cnt[threadIdx.x] = 0ull;   // 9 counters + 3 double sums (0.0 == all-zero bits)
    __syncthreads();
#define ENV_OFF(el) (flags[(el) * 4 + 2] == 0)

    if (mode != 2) {
        // ---- stage tile ----
        for (int t = threadIdx.x; t < EPB * nA; t += blockDim.x) {
            const int el = t / nA, a = t - el * nA, e = env0 + el;
            float *A = agents + (size_t)el * 5 * nA;
            float x = 0.f, y = 0.f, vx = 0.f, vy = 0.f;
            if (!ENV_OFF(el)) {
                if (a < Nh) {
                    const double2 q = D.hpos[(size_t)e * Nh + a];
                    const float2 v = D.hvel[(size_t)e * Nh + a];
                    x = (float)q.x; y = (float)q.y; vx = v.x; vy = v.y;
                } else {
                    // pt_step_kernel has already moved the robot: the humans act on its pose BEFORE the move
                    // (human_avoidance_env.py:129 precedes :132)
                    const double *rq = (mode == 0 ? D.rprev : D.rpose) + (size_t)e * 4;
                    x = (float)rq[0]; y = (float)rq[1];
                }
            }
            A[a] = x; A[nA + a] = y; A[2 * nA + a] = vx; A[3 * nA + a] = vy; A[4 * nA + a] = D.orca.radius;
        }
        __syncthreads();

        // the previous solve's class of every agent of this CTA (global state) -> queue order of pass 1
        for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x) {
            const int e = env0 + t / Nh;
            const size_t gidx = (size_t)env0 * Nh + t;
            const bool on = e < D.E;
            cls[t] = on ? D.lp_class[gidx] : 0;
            sstat[t] = on ? D.hstatic[gidx] : 1;
            pref[t] = on ? ha_pref_velocity(D.hpos[gidx], D.hgoal[gidx]) : make_float2(0.f, 0.f);   // pass 1: the humans where they are
        }
        __syncthreads();
        ha_build_order(order, cls, EPB * Nh, fill);
        // ---- ORCA pass 1: the humans' actions (human_avoidance_env.py:129 / :84) ----
        // (the DEDUP kernel uses ONE instantiation of the pass for both passes: two copies of the solve do not fit the
        // instruction cache together)
        orca_pass_tile<G, DEDUP, CUTOFF, PRUNE>(D, g, gi, NG, EPB, env0, nA, agents, lines, dist, vnew, posd, false,
                                 D.stat_counters ? &flags[EPB * 4 + 2] : nullptr, &flags[EPB * 4], mode == 0, flags, false, order, cls, pref, sstat);
        __syncthreads();
        if (threadIdx.x == 0 && D.stat_counters && flags[EPB * 4 + 2]) atomicAdd(D.stat_counters + 11, (unsigned long long)flags[EPB * 4 + 2]);
        if (mode == 1)   // no look-ahead pass in a warm-start step: these classes order the next pass
            for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x)
                if (env0 + t / Nh < D.E) D.lp_class[(size_t)env0 * Nh + t] = cls[t];

        DRE_HA_T(12);   // stage + pass 1
        // ---- integrate: robot (unicycle, agent.py:169-176) and humans (Euler, agent.py:164-168); histories ----
        for (int t = threadIdx.x; t < EPB * (Nh + 1); t += blockDim.x) {
            const int el = t / (Nh + 1), a = t - el * (Nh + 1), e = env0 + el;
            if (ENV_OFF(el)) continue;
            if (a < Nh) {
                const size_t gidx = (size_t)e * Nh + a;
                const double2 q = D.hpos[gidx];
                const float2 v = vnew[el * Nh + a];
                double2 np_;
                np_.x = q.x + (double)v.x * dt;
                np_.y = q.y + (double)v.y * dt;
                D.hpos[gidx] = np_;
                D.hvel[gidx] = v;
                posd[el * Nh + a] = np_;
                pref[el * Nh + a] = ha_pref_velocity(np_, D.hgoal[gidx]);   // pass 2: from the moved position
                float *A = agents + (size_t)el * 5 * nA;
                A[a] = (float)np_.x; A[nA + a] = (float)np_.y; A[2 * nA + a] = v.x; A[3 * nA + a] = v.y;
                // history roll (human_avoidance_env.py:149-158 / :92-100)
                double2 *hh = D.hhist + gidx * H;
                const int valid = D.hvalid[gidx];
                if (valid == H) {
                    for (int s = 0; s < H - 1; ++s) hh[s] = hh[s + 1];
                    hh[H - 1] = np_;
                } else {
                    hh[valid] = np_;
                    D.hvalid[gidx] = valid + 1 < H ? valid + 1 : H;
                }
                if (D.sensor_restriction) {   // out of the robot's sensor range: the history starts over (:102-105, :161-164)
                    double st, ct;
                    sincos(D.rpose[(size_t)e * 4 + 2], &st, &ct);
                    const double ddx = np_.x - D.rpose[(size_t)e * 4], ddy = np_.y - D.rpose[(size_t)e * 4 + 1];
                    const double lx = ct * ddx + st * ddy, ly = -st * ddx + ct * ddy;
                    if (sqrt(lx * lx + ly * ly) > D.sensor_range) D.hvalid[gidx] = 0;
                }
            } else {
                double *rp = D.rpose + (size_t)e * 4;
                double av = 0.0, aw = 0.0;
                if (mode == 0) {
                    av = D.out.actions_used[(size_t)e * 2]; aw = D.out.actions_used[(size_t)e * 2 + 1];   // set by pt_step_kernel
                    if (D.orca.robot_visible) {   // pass 2 sees the robot at its new pose (already in rpose)
                        float *A = agents + (size_t)el * 5 * nA;
                        A[Nh] = (float)rp[0]; A[nA + Nh] = (float)rp[1];
                    }
                }
                // robot history (human_avoidance_env.py:133-141 / :114-122)
                double *rh = D.rhist + (size_t)e * H * 4;
                double *pv = D.rpastv + (size_t)e * D.LB * 2;
                const int valid = D.rvalid[e];
                if (valid == H) {
                    for (int s = 0; s < (H - 1) * 4; ++s) rh[s] = rh[s + 4];
                    rh[(H - 1) * 4] = rp[0]; rh[(H - 1) * 4 + 1] = rp[1]; rh[(H - 1) * 4 + 2] = rp[2];
                    for (int s = 0; s < (D.LB - 1) * 2; ++s) pv[s] = pv[s + 2];
                    pv[(D.LB - 1) * 2] = av; pv[(D.LB - 1) * 2 + 1] = aw;
                } else {
                    rh[valid * 4] = rp[0]; rh[valid * 4 + 1] = rp[1]; rh[valid * 4 + 2] = rp[2];
                    if (valid - 1 >= 0 && valid - 1 < D.LB) { pv[(valid - 1) * 2] = av; pv[(valid - 1) * 2 + 1] = aw; }
                    D.rvalid[e] = valid + 1 < H ? valid + 1 : H;
                }
            }
        }
        __syncthreads();

    }

    // ---- observation (human_avoidance_env.py:24-77), frame = 'robot' ----
    if (write_obs) {
        const int W = 2 * H;
        // sensor restriction (:53-72): output row r shows the r-th VISIBLE human for r < nv and keeps human r's own row
        // beyond; one thread per env builds that row map (the reward scratch is free at this point of the kernel)
        int *src_row = reinterpret_cast<int *>(pen);
        if (D.sensor_restriction) {
            for (int el = threadIdx.x; el < EPB; el += blockDim.x) {
                const int e = env0 + el;
                if (ENV_OFF(el)) continue;
                int nv = 0;
                for (int i = 0; i < Nh; ++i) src_row[el * Nh + i] = i;
                for (int i = 0; i < Nh; ++i)
                    if (D.hvalid[(size_t)e * Nh + i] != 0) src_row[el * Nh + nv++] = i;
                flags[el * 4 + 1] = nv;
            }
            __syncthreads();
        }
        for (int t = threadIdx.x; t < EPB * Nh * H; t += blockDim.x) {
            const int el = t / (Nh * H), r = t - el * Nh * H, i = r / H, s = r - i * H, e = env0 + el;
            if (ENV_OFF(el)) continue;
            const int src = D.sensor_restriction ? src_row[el * Nh + i] : i;
            double st, ct;
            sincos(D.rpose[(size_t)e * 4 + 2], &st, &ct);
            const double rx = D.rpose[(size_t)e * 4], ry = D.rpose[(size_t)e * 4 + 1];
            const double2 q = D.hhist[((size_t)e * Nh + src) * H + s];
            const double dx = q.x - rx, dy = q.y - ry;
            double *o = D.out.spatial_edges + ((size_t)e * Nh + i) * W + 2 * s;
            o[0] = ct * dx + st * dy;
            o[1] = -st * dx + ct * dy;
            if (D.sensor_restriction && flags[el * 4 + 1] == 0 && i == 0) { o[0] = -10.0; o[1] = -10.0; }   // nobody in range: a fake far human (:66-68)
        }
        for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x) {
            const int el = t / Nh, i = t - el * Nh, e = env0 + el;
            if (ENV_OFF(el)) continue;
            double hv = (double)D.hvalid[(size_t)e * Nh + i];
            if (D.sensor_restriction) {
                const int nv = flags[el * 4 + 1];
                hv = nv == 0 ? 1.0 : (i < nv ? (double)D.hvalid[(size_t)e * Nh + src_row[el * Nh + i]] : 0.0);
            }
            D.out.human_valid[(size_t)e * Nh + i] = hv;
        }
        for (int t = threadIdx.x; t < EPB * D.LB; t += blockDim.x) {
            const int el = t / D.LB, s = t - el * D.LB, e = env0 + el;
            if (ENV_OFF(el)) continue;
            double st, ct;
            sincos(D.rpose[(size_t)e * 4 + 2], &st, &ct);
            const double rx = D.rpose[(size_t)e * 4], ry = D.rpose[(size_t)e * 4 + 1];
            const double *rh = D.rhist + ((size_t)e * H + s) * 4;
            const double dx = rh[0] - rx, dy = rh[1] - ry;
            D.out.robot_node[(size_t)e * 2 * D.LB + 2 * s] = ct * dx + st * dy;
            D.out.robot_node[(size_t)e * 2 * D.LB + 2 * s + 1] = -st * dx + ct * dy;
            D.out.past_robot_vel[(size_t)e * 2 * D.LB + 2 * s] = D.rpastv[((size_t)e * D.LB + s) * 2];
            D.out.past_robot_vel[(size_t)e * 2 * D.LB + 2 * s + 1] = D.rpastv[((size_t)e * D.LB + s) * 2 + 1];
        }
        for (int el = threadIdx.x; el < EPB; el += blockDim.x) {
            const int e = env0 + el;
            if (ENV_OFF(el)) continue;
            D.out.percent_episode[e] = D.gtime[e] / D.time_limit * 10.0 - 5.0;
            int det = Nh;
            if (D.sensor_restriction) det = flags[el * 4 + 1] > 0 ? flags[el * 4 + 1] : 1;
            D.out.detected_humans[e] = (double)det;
        }
        if (D.sensor_restriction) __syncthreads();   // the row map lives in the reward scratch
    }

    // ---- the observation of this CTA's envs is final: chunk flag for the pipelined host path (threadFenceReduction
    // pattern: the last CTA of a chunk sees every other CTA's fenced writes through the counter, then publishes
    // system-wide), raised BEFORE the look-ahead pass so the copy of a chunk overlaps the second half of its own CTAs ----
    if (D.chunk_flags != nullptr && mode == 0) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int ctas_per_chunk = D.chunk_envs / EPB;            // chunk_envs is a multiple of EPB
            const int c = blockIdx.x / ctas_per_chunk;
            const int first = c * ctas_per_chunk;
            const int total = (int)gridDim.x - first < ctas_per_chunk ? (int)gridDim.x - first : ctas_per_chunk;
            if (atomicAdd(D.chunk_count + c, 1) + 1 == total) {
                D.chunk_count[c] = 0;
                __threadfence_system();
                *reinterpret_cast<volatile int *>(D.chunk_flags + c) = D.step_seq;
            }
        }
    }

    DRE_HA_T(13);       // integrate + observation + chunk flag
    if (mode == 0) {
        // ---- ORCA pass 2 on the moved crowd: what the humans will do next (human_avoidance_env.py:245) ----
        float2 *vnext = vnew;   // pass-1 velocities now live in the tile / hvel
        ha_build_order(order, cls, EPB * Nh, fill);   // from the classes pass 1 has just produced
#ifndef DRE_HA_SKIP_PASS2   // measurement builds only: what does the kernel cost without the look-ahead pass?
        orca_pass_tile<G, DEDUP, CUTOFF, PRUNE>(D, g, gi, NG, EPB, env0, nA, agents, lines, dist, vnext, posd, true, nullptr, &flags[EPB * 4 + 1], false, flags, PRUNE, order, cls, pref, sstat);
#endif
        __syncthreads();
        for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x)   // the classes of the latest solves order the next step's pass 1
            if (env0 + t / Nh < D.E) D.lp_class[(size_t)env0 * Nh + t] = cls[t];
        if (DEDUP) {   // keep the look-ahead velocities: they are the next step's movement pass
            for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x) {
                const int el = t / Nh, i = t - el * Nh, e = env0 + el;
                if (ENV_OFF(el)) continue;
                D.hvel_next[(size_t)e * Nh + i] = vnext[el * Nh + i];
                D.goal_changed[(size_t)e * Nh + i] = 0;
            }
            for (int el = threadIdx.x; el < EPB; el += blockDim.x)
                if (!ENV_OFF(el)) D.cache_valid[env0 + el] = 1;
        }

        DRE_HA_T(14);   // pass 2
        // ---- per-human reward terms (human_avoidance_env.py:205-216, 247-273) ----
        for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x) {
            const int el = t / Nh, i = t - el * Nh, e = env0 + el;
            if (ENV_OFF(el)) continue;
            const double2 hp = posd[el * Nh + i];
            const double rx = D.rpose[(size_t)e * 4], ry = D.rpose[(size_t)e * 4 + 1];
            const double dx = hp.x - rx, dy = hp.y - ry;
            const double closest = sqrt(dx * dx + dy * dy) - D.human_radius - D.robot_radius;
            double pv_ = 0.0, pth_ = 0.0;
            if (!D.hstatic[(size_t)e * Nh + i] && !(closest > D.dist_radius || closest <= 0.0)) {
                const float *A = agents + (size_t)el * 5 * nA;
                const double cvx = (double)A[2 * nA + i], cvy = (double)A[3 * nA + i];
                const double fvx = (double)vnext[el * Nh + i].x, fvy = (double)vnext[el * Nh + i].y;
                const double curr_v = sqrt(cvx * cvx + cvy * cvy), fut_v = sqrt(fvx * fvx + fvy * fvy);
                const double curr_th = atan2(cvy, cvx), fut_th = atan2(fvy, fvx);
                const double dvm = fut_v - curr_v;
                const double pen_v_i = D.dist_factor_v * (dvm < 0.0 ? dvm : 0.0);
                const double th_diff = wrap_pi(fut_th - curr_th);
                const double pen_th_i = -D.dist_factor_th * fabs(th_diff);
                const double scaling = 1.0 - closest / D.dist_radius;
                pv_ = pen_v_i * scaling; pth_ = pen_th_i * scaling;
            }
            pen[(el * Nh + i) * 3] = pv_; pen[(el * Nh + i) * 3 + 1] = pth_; pen[(el * Nh + i) * 3 + 2] = closest;
        }
        __syncthreads();
        DRE_HA_T2(12);  // per-human reward terms

        // ---- per-env reward / done (one thread per env sums in human order, like the reference loop) ----
        for (int el = threadIdx.x; el < EPB; el += blockDim.x) {
            const int e = env0 + el;
            if (ENV_OFF(el)) continue;
            double dmin = INFINITY, pen_v = 0.0, pen_th = 0.0;
            bool collision = false;
            for (int i = 0; i < Nh; ++i) {
                const double c = pen[(el * Nh + i) * 3 + 2];
                if (c <= 0.0) collision = true;
                else if (c < dmin) dmin = c;
                pen_v += pen[(el * Nh + i) * 3]; pen_th += pen[(el * Nh + i) * 3 + 1];
            }
            if (collision) dmin = 0.0;
            double reward = 0.0;
            uint32_t bits = 0;
            if (collision) { reward += D.collision_penalty; bits |= DRE_DONE_COLLISION | DRE_DONE_HA; }
            if (dmin < D.safety_dist) { reward += D.safety_penalty; bits |= DRE_DONE_SAFETY_HUMAN | DRE_DONE_HA; }
            reward += (pen_v + pen_th);
            double total_v = 0.0;
            const double *pv = D.rpastv + (size_t)e * D.LB * 2;
            for (int s = 0; s < D.LB; ++s) total_v += fabs(pv[s * 2]);
            if (total_v < D.act_min_vel) { reward += D.act_penalty; bits |= DRE_DONE_ACTUATION | DRE_DONE_HA; }
            D.out.info[(size_t)e * 8 + 0] = pen_v; D.out.info[(size_t)e * 8 + 1] = pen_th; D.out.info[(size_t)e * 8 + 4] = dmin;
            // ---- merge with the PT part written by pt_step_kernel (HA_and_PT env.py:322-331, 391-414) ----
            const double r_pt = D.out.rewards[(size_t)e * 3];
            uint32_t all = D.out.done_bits[e] | bits;
            if (all & (DRE_DONE_HA | DRE_DONE_PT)) all |= DRE_DONE_ANY | DRE_DONE_FOR_REPLAY;
            D.out.rewards[(size_t)e * 3 + 1] = reward;
            D.out.rewards[(size_t)e * 3 + 2] = r_pt + reward;
            D.out.done_bits[e] = all;
            {   // unpacked 0/1 flags so the host mirror can hand out views without launching kernels
                const uint32_t order[12] = {DRE_DONE_PT, DRE_DONE_HA, DRE_DONE_ANY, DRE_DONE_FOR_REPLAY, DRE_DONE_GOAL, DRE_DONE_COLLISION,
                                            DRE_DONE_SAFETY_HUMAN, DRE_DONE_TIMEOUT, DRE_DONE_END_OF_PATH, DRE_DONE_CORRIDOR_HIT,
                                            DRE_DONE_SAFETY_CORRIDOR, DRE_DONE_ACTUATION};
                uint32_t w0 = 0, w1 = 0, w2 = 0;
                for (int q = 0; q < 4; ++q) {
                    w0 |= (uint32_t)((all & order[q]) != 0) << (8 * q);
                    w1 |= (uint32_t)((all & order[4 + q]) != 0) << (8 * q);
                    w2 |= (uint32_t)((all & order[8 + q]) != 0) << (8 * q);
                }
                uint32_t *df = reinterpret_cast<uint32_t *>(D.out.done_flags + (size_t)e * 16);
                df[0] = w0; df[1] = w1; df[2] = w2;   // df[3] (soft-reset flags) belongs to pt_step_kernel
            }
            if (D.stat_counters) {
                const uint32_t sb[8] = {DRE_DONE_GOAL, DRE_DONE_COLLISION, DRE_DONE_SAFETY_HUMAN, DRE_DONE_TIMEOUT, DRE_DONE_END_OF_PATH,
                                        DRE_DONE_CORRIDOR_HIT, DRE_DONE_SAFETY_CORRIDOR, DRE_DONE_ACTUATION};
                atomicAdd(&s_cnt[0], 1ull);
                for (int q = 0; q < 8; ++q)
                    if (all & sb[q]) atomicAdd(&s_cnt[1 + q], 1ull);
                atomicAdd(&s_sum[0], r_pt + reward); atomicAdd(&s_sum[1], r_pt); atomicAdd(&s_sum[2], reward);
            }
        }
        __syncthreads();
        if (D.stat_counters) {   // one global atomic per counter per CTA
            if (threadIdx.x < 9 && s_cnt[threadIdx.x]) atomicAdd(D.stat_counters + threadIdx.x, s_cnt[threadIdx.x]);
            if (threadIdx.x >= 9 && threadIdx.x < 12) atomicAdd(D.stat_sums + (threadIdx.x - 9), s_sum[threadIdx.x - 9]);
        }
    }

    DRE_HA_T2(13);      // per-env merge + statistics
    // ---- goal resampling for humans that reached their goal (human_avoidance_env.py:181-184 / :109-112) ----
    if (mode != 2 && D.end_goal_changing) {
        for (int t = threadIdx.x; t < EPB * Nh; t += blockDim.x) {
            const int el = t / Nh, i = t - el * Nh, e = env0 + el;
            if (ENV_OFF(el) || D.hstatic[(size_t)e * Nh + i]) continue;
            const double2 p = posd[el * Nh + i], gl = D.hgoal[(size_t)e * Nh + i];
            const double dx = gl.x - p.x, dy = gl.y - p.y;
            if (sqrt(dx * dx + dy * dy) < D.human_radius) atomicOr(&flags[el * 4], 1);
        }
        __syncthreads();
        DRE_HA_T2(14);  // goal-reached detection
        for (int el = threadIdx.x; el < EPB; el += blockDim.x) {
            const int e = env0 + el;
            if (ENV_OFF(el) || !flags[el * 4]) continue;
            const uint64_t gid = (uint64_t)(D.env_id_offset + e);
            const uint64_t stp = (uint64_t)D.step_count[e];
            const uint32_t epoch = (uint32_t)D.reset_epoch[e], kind = mode == 1 ? DRE_RNG_WARM : DRE_RNG_STEP, sub = mode == 1 ? (uint32_t)D.rvalid[e] : 0u;
            const double rgx = 0.0, rgy = D.circle_radius;   // robot goal, HA_and_PT env.py:293-299
            const double rpx = D.rpose[(size_t)e * 4], rpy = D.rpose[(size_t)e * 4 + 1];
            for (int i = 0; i < Nh; ++i) {
                if (D.hstatic[(size_t)e * Nh + i]) continue;
                const double2 p = posd[el * Nh + i];
                double2 gl = D.hgoal[(size_t)e * Nh + i];
                const double ddx = gl.x - p.x, ddy = gl.y - p.y;
                if (!(sqrt(ddx * ddx + ddy * ddy) < D.human_radius)) continue;
                const double md_r = D.human_radius + D.robot_radius + 0.2, md_h = D.human_radius + D.human_radius + 0.2;
                double gx = gl.x, gy = gl.y;
                for (uint32_t attempt = 0; attempt < 4096; ++attempt) {
                    double u0, u1, u2, u3;
                    env_uniform2(D.seed, epoch, gid, kind, sub, stp, (uint32_t)i, 2 * attempt, u0, u1);
                    env_uniform2(D.seed, epoch, gid, kind, sub, stp, (uint32_t)i, 2 * attempt + 1, u2, u3);
                    const double angle = u0 * DRE_PI * 2.0;
                    double sa, ca;
                    sincos(angle, &sa, &ca);
                    gx = D.circle_radius_humans * ca + (u1 - 0.5);
                    gy = D.circle_radius_humans * sa + (u2 - 0.5);
                    bool collide = false;
                    {
                        const double a = gx - rgx, b = gy - rgy, c = gx - rpx, d = gy - rpy;
                        if (sqrt(a * a + b * b) < md_r || sqrt(c * c + d * d) < md_r) collide = true;
                    }
                    for (int j = 0; j < Nh && !collide; ++j) {
                        if (j == i) continue;
                        const double2 op = posd[el * Nh + j], og = D.hgoal[(size_t)e * Nh + j];
                        const double a = gx - og.x, b = gy - og.y, c = gx - op.x, d = gy - op.y;
                        if (sqrt(a * a + b * b) < md_h || sqrt(c * c + d * d) < md_h) collide = true;
                    }
                    if (!collide) break;
                }
                D.hgoal[(size_t)e * Nh + i] = make_double2(gx, gy);
                if (DEDUP) D.goal_changed[(size_t)e * Nh + i] = 1;
            }
        }
    }
    DRE_HA_T(15);       // rewards + merge + goal resampling
    DRE_HA_T2(15);      // goal resampling
    // a warm-start step moves the crowd without a look-ahead pass: whatever was cached is stale
    if (DEDUP && mode == 1)
        for (int el = threadIdx.x; el < EPB; el += blockDim.x)
            if (!ENV_OFF(el)) D.cache_valid[env0 + el] = 0;
#undef ENV_OFF
}

// ---------------------------------------------------------------------------------------------------------------
// PT step: one thread per env
// ---------------------------------------------------------------------------------------------------------------
// mode 0: step; mode 1: observe only (reset / soft reset: localise + state, no rewards)
//
// PT_LANES = 8 lanes per env: the path scans (nearest two nodes of <= 125) and the 25-node PT state are shared by the
// lanes, the scalar logic in between runs redundantly on all of them (identical inputs -> identical results) and lane 0
// stores. The kernel sits on the critical path in front of both the crowd half and the MPC solve and is pure latency
// (one thread per env: 36 us; 8 lanes per env: see profiles/README.md). Shuffles use the full-warp mask: every thread of
// a warp stays in the kernel and takes the same branches per env group up to data-dependent `if`s without shuffles inside.
#define PT_LANES 8

// lexicographic (distance^2, index) order: what the sequential scan's strict `<` implements
__device__ __forceinline__ bool pt_key_lt(double d, int i, double e, int j) { return d < e || (d == e && i < j); }

// path_nearest_two (env_device.cuh) with the scan striped over the PT_LANES lanes of a group
__device__ __forceinline__ void path_nearest_two_lanes(const PathView &P, double qx, double qy, int lane, int &i0, int &i1)
{
    double d0 = INFINITY, d1 = INFINITY;
    int a0 = 0x7ffffff0, a1 = 0x7ffffff1;
    for (int i = lane; i < P.L; i += PT_LANES) {
        const double dx = P.x[i] - qx, dy = P.y[i] - qy;
        const double d = dx * dx + dy * dy;
        if (d < d0) { d1 = d0; a1 = a0; d0 = d; a0 = i; }
        else if (d < d1) { d1 = d; a1 = i; }
    }
#pragma unroll
    for (int o = 1; o < PT_LANES; o <<= 1) {
        const double e0 = __shfl_xor_sync(0xffffffffu, d0, o), e1 = __shfl_xor_sync(0xffffffffu, d1, o);
        const int b0 = __shfl_xor_sync(0xffffffffu, a0, o), b1 = __shfl_xor_sync(0xffffffffu, a1, o);
        if (pt_key_lt(d0, a0, e0, b0)) {                       // mine first: second = min(my second, partner's first)
            if (pt_key_lt(e0, b0, d1, a1)) { d1 = e0; a1 = b0; }
        } else {                                               // partner's first: second = min(partner's second, my first)
            const double f0 = d0; const int c0 = a0;
            d0 = e0; a0 = b0;
            if (pt_key_lt(e1, b1, f0, c0)) { d1 = e1; a1 = b1; } else { d1 = f0; a1 = c0; }
        }
    }
    i0 = a0; i1 = a1;
}

__device__ __forceinline__ Localization localize_lanes(const PathView &P, double qx, double qy, bool lts, int lane)
{
    int i0, i1;
    path_nearest_two_lanes(P, qx, qy, lane, i0, i1);
    return localize_from_nearest(P, qx, qy, lts, i0, i1);
}

// 8 CTAs per SM (64 registers, a few spilled words): 16384 envs x 8 lanes then fit in ONE wave (0.033 -> 0.026 ms; at 92
// registers the grid needed 1.4 waves)
__global__ void __launch_bounds__(128, 8) pt_step_kernel(EnvDev D, int mode)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = gt & (PT_LANES - 1);
    const bool in_range = (gt / PT_LANES) < D.E;
    // Surplus groups of the last warp (and envs a masked launch leaves alone) run the same code -- the shuffles below use the
    // full-warp mask -- and store nothing. Surplus groups work on constants: they must not read state another group writes.
    const int e = in_range ? gt / PT_LANES : 0;
    const bool valid = in_range && (D.env_mask == nullptr || D.env_mask[e] != 0);
    const bool writer = valid && lane == 0;
    const unsigned gshift = (threadIdx.x & 31u) & ~(unsigned)(PT_LANES - 1);

    int pid = in_range ? D.path_idx[e] : 0;
    PathView P = path_view(D.paths, D.path_stride, pid, D.lens);
    double *rp = D.rpose + (size_t)e * 4;
    double cx = 0.0, cy = 0.0, cth = 0.0;
    bool lts = true;
    uint32_t prev_bits = 0u;
    double av = 0.0, aw = 0.0;
    if (in_range) {
        cx = rp[0]; cy = rp[1]; cth = rp[2];
        lts = D.loc_to_start[e] != 0;
        prev_bits = D.prev_done_bits[e];
        if (mode == 0) { av = D.actions[(size_t)e * 2]; aw = D.actions[(size_t)e * 2 + 1]; }
    }
    __syncwarp();                                             // every lane has read the state lane 0 is about to overwrite
    Localization cur;
    if (mode == 0) {
        // ---- which action runs: the caller's, or the scripted one of an env recovering from a termination ----
        uint32_t sr_flags = 0;
        if (D.soft_reset) {
            double sv = av, sw = aw;
            if (lane == 0 && in_range) {
                SrState S = sr_unpack(D.sr_state + (size_t)e * 4);
                bool stuck;
                if (soft_reset_decide(S, prev_bits, P, cx, cy, cth, lts, sv, sw, stuck)) sr_flags |= 1u;
                if (stuck) sr_flags |= 1u << 8;
                if (writer) sr_pack(S, D.sr_state + (size_t)e * 4);
            }
            av = __shfl_sync(0xffffffffu, sv, gshift); aw = __shfl_sync(0xffffffffu, sw, gshift);
            sr_flags = __shfl_sync(0xffffffffu, sr_flags, gshift);
        }
        if (writer) {
            D.out.actions_used[(size_t)e * 2] = av; D.out.actions_used[(size_t)e * 2 + 1] = aw;
            reinterpret_cast<uint32_t *>(D.out.done_flags + (size_t)e * 16)[3] = sr_flags;
        }
        // ---- robot.step(action): exact-arc unicycle (agent.py:169-176 -> unicycle.py:89-126) ----
        const double px = cx, py = cy, pth = cth;
        unicycle_step(px, py, pth, av, aw, D.dt, cx, cy, cth);
        if (writer) {
            D.rprev[(size_t)e * 4] = px; D.rprev[(size_t)e * 4 + 1] = py; D.rprev[(size_t)e * 4 + 2] = pth;
            rp[0] = cx; rp[1] = cy; rp[2] = cth;
            D.gtime[e] += D.dt;
            D.step_count[e] += 1;
        }
        const Localization prv = localize_lanes(P, px, py, lts, lane);
        cur = localize_lanes(P, cx, cy, lts, lane);
        if (cur.index >= (double)(P.L / 3) && cur.index < (double)(2 * P.L / 3)) lts = false;
        // ---- rewards in strategy order (config_PT.py:12; reward_comp.py:20-32), each scaled by overall_scaling ----
        double reward = 0.0;
        uint32_t bits = 0;
        const double sc = D.pt_overall_scaling;
        {   // goal (reward_comp.py:34-54): the five intermediate poses, one per lane
            const double gx = P.x[P.L - 1], gy = P.y[P.L - 1], gth = P.th[P.L - 1];
            bool hit = false;
            for (int i = 1 + lane; i <= 5; i += PT_LANES) {
                double qx, qy, qth;
                unicycle_step(px, py, pth, av, aw, D.dt * (double)i / 5.0, qx, qy, qth);
                const double ddx = qx - gx, ddy = qy - gy;
                const double ad = fabs(np_mod(qth - gth + DRE_PI, 2.0 * DRE_PI) - DRE_PI);
                if (sqrt(ddx * ddx + ddy * ddy) < D.goal_radius && ad < D.goal_angle) hit = true;
            }
            hit = ((__ballot_sync(0xffffffffu, hit) >> gshift) & ((1u << PT_LANES) - 1u)) != 0u;
            if (hit) { reward += sc * D.goal_reward; bits |= DRE_DONE_GOAL | DRE_DONE_PT; }
            else reward += sc * 0.0;
        }
        double adv;
        {   // path_advancement (reward_comp.py:101-116)
            double pi_ = prv.index, ci_ = cur.index;
            if (fabs(ci_ - pi_) > 10.0) {
                const double half = (double)(P.L / 2);
                pi_ = np_mod(pi_ + half, (double)P.L);
                ci_ = np_mod(ci_ + half, (double)P.L);
            }
            adv = sc * (D.path_adv_scaling * (ci_ - pi_));
            reward += adv;
        }
        const double ddx = cx - cur.ix, ddy = cy - cur.iy;
        const double xy_dev = sqrt(ddx * ddx + ddy * ddy);
        double devr;
        {   // deviation (reward_comp.py:89-99)
            const double th_dev = np_mod(cth - cur.ith + DRE_PI, 2.0 * DRE_PI) - DRE_PI;
            const double th_reward = -D.dev_th * fabs(th_dev);
            const double xy_reward = -xy_dev * D.dev_xy;
            devr = sc * ((xy_reward + th_reward) * D.dev_scale);
            reward += devr;
        }
        if (P.L - 1 == (int)prv.index) { reward += sc * D.eop_penalty; bits |= DRE_DONE_END_OF_PATH | DRE_DONE_PT; }
        else reward += sc * 0.0;
        if (D.corridor_max_dev < xy_dev) { reward += sc * D.corridor_penalty; bits |= DRE_DONE_CORRIDOR_HIT | DRE_DONE_PT; }
        else reward += sc * 0.0;
        if (D.safety_corridor_max_dev < xy_dev) { reward += sc * D.safety_corridor_penalty; bits |= DRE_DONE_SAFETY_CORRIDOR | DRE_DONE_PT; }
        else reward += sc * 0.0;
        // the HA half (ha_step_kernel) merges rewards / dones (HA_and_PT env.py:322-331)
        if (writer) {
            D.out.rewards[(size_t)e * 3] = reward;
            D.out.done_bits[e] = bits;
            D.out.info[(size_t)e * 8 + 2] = adv; D.out.info[(size_t)e * 8 + 3] = devr;
            D.out.info[(size_t)e * 8 + 5] = cur.index; D.out.info[(size_t)e * 8 + 6] = xy_dev; D.out.info[(size_t)e * 8 + 7] = prv.index;
        }
        // ---- path switch on goal / end of path (what soft_reset does, HA_and_PT env.py:130-141) ----
        // (`bits` is the same on all lanes of the group and on all groups' surplus shadows, so the shuffles inside
        // localize_lanes are executed by whole groups; groups of a warp may differ, hence the explicit convergence)
        const bool sw = D.auto_path_switch && (bits & (DRE_DONE_GOAL | DRE_DONE_END_OF_PATH));
        const bool any_sw = __any_sync(0xffffffffu, sw);
        if (any_sw) {
            const int npid = sw ? (pid + 1) % D.num_paths : pid;
            const PathView NP = path_view(D.paths, D.path_stride, npid, D.lens);
            const Localization nl = localize_lanes(NP, cx, cy, sw ? true : lts, lane);   // all lanes of the warp take part
            if (sw) {
                pid = npid; P = NP; lts = true; cur = nl;
                if (writer) { D.path_idx[e] = pid; D.mpc_iteration0[e] = 1; }
            }
        }
        if (writer) D.loc_to_start[e] = lts ? 1 : 0;
    } else {
        cur = localize_lanes(P, cx, cy, lts, lane);
    }
    if (valid)
        pt_state_gen(P, cx, cy, cth, cur.index, D.lookahead, D.lookbehind, D.out.pt_state + (size_t)e * (4 * (D.lookahead + D.lookbehind)), 1,
                     lane, PT_LANES);
    if (writer) D.mpc_interp[e] = cur.index;
}

// MPC retry / give-up statistics after the solve
__global__ void mpc_stats_kernel(EnvDev D)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long r = 0, gup = 0;
    if (e < D.E && (D.env_mask == nullptr || D.env_mask[e] != 0)) { r = (unsigned long long)D.out.mpc_status[e].retries; gup = D.out.mpc_status[e].term < 0 ? 1 : 0; }
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); gup += __shfl_xor_sync(0xffffffffu, gup, o); }
    if ((threadIdx.x & 31) == 0) {
        if (r) atomicAdd(D.stat_counters + 9, r);
        if (gup) atomicAdd(D.stat_counters + 10, gup);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// reset: one thread per env
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) env_reset_kernel(EnvDev D, const double2 *hpos_init, const double2 *hgoal_init)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.E || (D.env_mask != nullptr && D.env_mask[e] == 0)) return;
    const int Nh = D.Nh, H = D.LB + 1;
    const double rx = 0.0, ry = -D.circle_radius, rth = DRE_PI;   // set_robot_reset(start_angle=pi), HA_and_PT env.py:264,286-299
    double *rp = D.rpose + (size_t)e * 4;
    rp[0] = rx; rp[1] = ry; rp[2] = rth; rp[3] = 0.0;
    double *rv = D.rprev + (size_t)e * 4;
    rv[0] = rx; rv[1] = ry; rv[2] = rth; rv[3] = 0.0;
    for (int s = 0; s < H; ++s) {
        double *rh = D.rhist + ((size_t)e * H + s) * 4;
        rh[0] = rx; rh[1] = ry; rh[2] = rth; rh[3] = 0.0;
    }
    for (int s = 0; s < D.LB * 2; ++s) D.rpastv[(size_t)e * D.LB * 2 + s] = 0.0;
    D.rvalid[e] = 1;
    D.gtime[e] = 0.0;
    D.step_count[e] = 0;
    D.reset_epoch[e] += 1;
    D.path_idx[e] = 0;
    D.loc_to_start[e] = 1;
    D.mpc_iteration0[e] = 1;
    for (int i = 0; i < 4; ++i) { D.sr_state[(size_t)e * 4 + i] = 0; reinterpret_cast<uint32_t *>(D.out.done_flags + (size_t)e * 16)[i] = 0u; }
    D.out.done_bits[e] = 0u;
    D.cache_valid[e] = 0;
    const uint64_t gid = (uint64_t)(D.env_id_offset + e);
    const uint32_t epoch = (uint32_t)D.reset_epoch[e];
    for (int i = 0; i < Nh; ++i) {
        const size_t gidx = (size_t)e * Nh + i;
        double px, py, gx, gy;
        const int si = i - (Nh - D.num_static);   // >= 0: static human (generate_static_humans, crowd_sim.py:171-183)
        if (si >= 0 && !hpos_init) {
            px = D.static_pos[2 * si]; py = D.static_pos[2 * si + 1]; gx = 0.0; gy = 0.0;   // human.set(px, py, 0, 0, ...)
        } else if (hpos_init) {
            px = hpos_init[gidx].x; py = hpos_init[gidx].y;
            gx = hgoal_init ? hgoal_init[gidx].x : -px; gy = hgoal_init ? hgoal_init[gidx].y : -py;
        } else {
            // generate_circle_crossing_human (crowd_sim.py:232-262)
            px = 0.0; py = 0.0;
            for (uint32_t attempt = 0; attempt < 4096; ++attempt) {
                double u0, u1, u2, u3;
                env_uniform2(D.seed, epoch, gid, DRE_RNG_RESET, 0u, 0ull, (uint32_t)i, 2 * attempt, u0, u1);
                env_uniform2(D.seed, epoch, gid, DRE_RNG_RESET, 0u, 0ull, (uint32_t)i, 2 * attempt + 1, u2, u3);
                double sa, ca;
                sincos(u0 * DRE_PI * 2.0, &sa, &ca);
                px = D.circle_radius_humans * ca + u1 * 1.0;
                py = D.circle_radius_humans * sa + u2 * 1.0;
                bool collide = false;
                {
                    const double a = px - rx, b = py - ry;
                    if (sqrt(a * a + b * b) < D.human_radius + D.robot_radius + 0.5) collide = true;
                }
                for (int j = 0; j < i && !collide; ++j) {   // earlier (dynamic) humans; statics are appended afterwards
                    const double2 q = D.hpos[(size_t)e * Nh + j];
                    const double a = px - q.x, b = py - q.y;
                    if (sqrt(a * a + b * b) < D.human_radius + D.human_radius + 0.2) collide = true;
                }
                if (!collide) break;
            }
            gx = -px; gy = -py;
        }
        D.hpos[gidx] = make_double2(px, py);
        D.hgoal[gidx] = make_double2(gx, gy);
        D.hvel[gidx] = make_float2(0.f, 0.f);
        D.hstatic[gidx] = si >= 0 ? 1 : 0;
        D.lp_class[gidx] = 0;
        int v0 = 1;
        if (D.sensor_restriction) {   // human_avoidance_env.py:331-334
            double st, ct;
            sincos(rth, &st, &ct);
            const double ddx = px - rx, ddy = py - ry;
            const double lx = ct * ddx + st * ddy, ly = -st * ddx + ct * ddy;
            if (sqrt(lx * lx + ly * ly) > D.sensor_range) v0 = 0;
        }
        D.hvalid[gidx] = v0;
        for (int s = 0; s < H; ++s) D.hhist[gidx * H + s] = make_double2(px, py);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// CVMM safety filter: HAAndPTEnv.cvmm_diverse_safety_pipeline / cvmm_safety_check (HA_and_PT env.py:422-513).
// One warp per env, lane c = candidate action c (0 = the policy's action, 1..A = the back-up set): constant-velocity
// roll-out of the humans against the exact-arc roll-out of the robot, plus the corridor test on the first two steps.
// ---------------------------------------------------------------------------------------------------------------
__device__ bool cvmm_safety_check(const EnvDev &D, int e, const PathView &P, bool lts, double v, double w, int num_steps, double &fde)
{
    double x = D.rpose[(size_t)e * 4], y = D.rpose[(size_t)e * 4 + 1], th = D.rpose[(size_t)e * 4 + 2];
    const double lim = D.robot_radius + D.human_radius + 0.22 + 1e-3;   // continuous task (:492-494)
    double dist_to_path = 0.0;
    fde = 1e10;
    for (int i = 0; i < num_steps; ++i) {
        double nx, ny, nth;
        unicycle_step(x, y, th, v, w, D.dt, nx, ny, nth);
        x = nx; y = ny; th = nth;
        const double tt = D.dt * (double)(i + 1);
        for (int j = 0; j < D.Nh; ++j) {
            const double2 hp = D.hpos[(size_t)e * D.Nh + j];
            const float2 hv = D.hvel[(size_t)e * D.Nh + j];
            const double hx = hp.x + (double)hv.x * tt, hy = hp.y + (double)hv.y * tt;
            const double dx = hx - x, dy = hy - y;
            if (sqrt(dx * dx + dy * dy) < lim) return false;
        }
        if (i < 2) {   // query_off_path (path_tracking_env.py:207-214)
            const Localization loc = localize_on_path(P, x, y, lts);
            const double ddx = x - loc.ix, ddy = y - loc.iy;
            dist_to_path = sqrt(ddx * ddx + ddy * ddy);
            if (D.safety_corridor_max_dev < dist_to_path) return false;
        }
    }
    fde = dist_to_path;
    return true;
}

// info[e] = {original_safe, num_safe among the back-up actions (0 if the original was safe), chosen back-up index
// (-1 = the original action), 1 if the choice was the random fall-back (no safe action)}
__global__ void __launch_bounds__(128) cvmm_filter_kernel(EnvDev D, const double *__restrict__ actions, int num_steps,
                                                          const double *__restrict__ backup, int A, double *__restrict__ safe_out,
                                                          int32_t *__restrict__ info, uint8_t *__restrict__ cand_safe)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, c = threadIdx.x & 31;
    if (e >= D.E) return;
    const PathView P = path_view(D.paths, D.path_stride, D.path_idx[e], D.lens);
    const bool lts = D.loc_to_start[e] != 0;
    const double tv = actions[(size_t)e * 2], tw = actions[(size_t)e * 2 + 1];
    double v = tv, w = tw;
    if (c >= 1 && c <= A) { v = backup[(size_t)(c - 1) * 2]; w = backup[(size_t)(c - 1) * 2 + 1]; }
    bool safe = false;
    double fde = 1e10;
    if (c <= A) safe = cvmm_safety_check(D, e, P, lts, v, w, num_steps, fde);
    if (cand_safe != nullptr && c <= A) cand_safe[(size_t)e * (A + 1) + c] = safe ? 1 : 0;
    const unsigned all = __ballot_sync(0xffffffffu, safe);
    const bool orig_safe = (all & 1u) != 0u;
    const unsigned bm = all & ~1u;                     // safe back-up actions
    double ov = tv, ow = tw;
    int chosen = -1, nsafe = 0, random = 0;
    if (!orig_safe) {
        nsafe = __popc(bm);
        if (nsafe == 0) {   // random back-up action (:459-464; np.random there, Philox here)
            double u0, u1;
            env_uniform2(D.seed, (uint32_t)D.reset_epoch[e], (uint64_t)(D.env_id_offset + e), DRE_RNG_CVMM, 0u, (uint64_t)D.step_count[e], 0xC7u, 0u, u0, u1);
            chosen = (int)(u0 * (double)A);
            if (chosen >= A) chosen = A - 1;
            random = 1;
        } else {
            // highest speed among the safe ones, then the angular velocity closest to the policy's (first occurrence) (:467-477)
            double vmax = (bm >> c) & 1u ? v : -INFINITY;
            for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            const bool at_max = ((bm >> c) & 1u) && v == vmax;
            double diff = at_max ? fabs(w - tw) : INFINITY;
            double dmin = diff;
            for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            const unsigned pick = __ballot_sync(0xffffffffu, at_max && diff == dmin);
            chosen = __ffs((int)pick) - 2;             // lane c holds back-up action c - 1
        }
        ov = backup[(size_t)chosen * 2]; ow = backup[(size_t)chosen * 2 + 1];
    }
    if (c == 0) {
        safe_out[(size_t)e * 2] = ov; safe_out[(size_t)e * 2 + 1] = ow;
        if (info != nullptr) { info[(size_t)e * 4] = orig_safe ? 1 : 0; info[(size_t)e * 4 + 1] = nsafe; info[(size_t)e * 4 + 2] = chosen; info[(size_t)e * 4 + 3] = random; }
    }
}

// HAAndPTEnv.cvmm_safety_check (HA_and_PT env.py:475-513) for one candidate action per env -> (safe, FDE_to_path)
__global__ void __launch_bounds__(128) cvmm_check_kernel(EnvDev D, const double *__restrict__ actions, int num_steps,
                                                         uint8_t *__restrict__ safe_out, double *__restrict__ dist_out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.E) return;
    const PathView P = path_view(D.paths, D.path_stride, D.path_idx[e], D.lens);
    double fde = 1e10;
    const bool safe = cvmm_safety_check(D, e, P, D.loc_to_start[e] != 0, actions[(size_t)e * 2], actions[(size_t)e * 2 + 1], num_steps, fde);
    safe_out[e] = safe ? 1 : 0;
    if (dist_out) dist_out[e] = fde;
}

int dre_env_launch_cvmm_check(const EnvDev &D, const double *actions, int num_steps, uint8_t *safe_out, double *dist_out, cudaStream_t s)
{
    if (num_steps < 1) return DRE_ERR_INVALID;
    cvmm_check_kernel<<<(D.E + 127) / 128, 128, 0, s>>>(D, actions, num_steps, safe_out, dist_out);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_env_launch_cvmm(const EnvDev &D, const double *actions, int num_steps, const double *backup, int A, double *safe_out,
                        int32_t *info, uint8_t *cand_safe, cudaStream_t s)
{
    if (A < 1 || A > 31 || num_steps < 1) return DRE_ERR_INVALID;
    const int threads = 128, envs_per_block = threads / 32;
    cvmm_filter_kernel<<<(D.E + envs_per_block - 1) / envs_per_block, threads, 0, s>>>(D, actions, num_steps, backup, A, safe_out, info, cand_safe);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
// threads per CTA of ha_step_kernel; DRE_HA_THREADS / DRE_HA_AGENTS_PER_GROUP override them for A/B measurements
static int dre_ha_threads()
{
    static int v = 0;
    if (v == 0) { const char *e = getenv("DRE_HA_THREADS"); v = e ? atoi(e) : 128; if (v != 64 && v != 128) v = 128; }   // the kernel is bounded to 128 threads
    return v;
}
static int dre_ha_agents_per_group()
{
    static int v = 0;
    if (v == 0) { const char *e = getenv("DRE_HA_AGENTS_PER_GROUP"); v = e ? atoi(e) : 8; if (v < 1 || v > 64) v = 8; }
    return v;
}

// the fused step kernel only has the lane-group passes: the stand-alone thread-per-agent variant (group_lanes = 1, a measurement
// aid of BatchedORCA) maps to the automatic choice here
static int dre_env_pick_group(const EnvDev &D)
{
    dre_orca_params q = D.orca;
    if ((q.group_lanes & 0xff) == 1) q.group_lanes = 0;
    return dre_orca_pick_group(q, D.Nh);
}

int dre_env_ha_envs_per_cta(const EnvDev &D)
{
    const int G = dre_env_pick_group(D);
    const int NG = dre_ha_threads() / G;
    const int cap = (D.E + 591) / 592;                       // keep >= 4 CTAs per SM in flight for small batches
    int base = (dre_ha_agents_per_group() * NG) / D.Nh;
    if (base > cap) base = cap;
    if (base < 1) base = 1;
    static const bool pinned = getenv("DRE_HA_AGENTS_PER_GROUP") != nullptr;   // measurement aid: no automatic choice
    if (pinned || base < 4) return base;
    // Wave quantisation: a grid just over a whole number of waves ends in an almost empty last wave. Measured on B200 at
    // 16384 x 20 (9 CTAs per SM = 1332 slots): 6 envs per CTA = 2731 CTAs = 2.05 waves: ha_step 0.658 ms (dense window) /
    // 0.482 (sparse); 7 envs = 1.76 waves: 0.632 / 0.474; 5 envs = 2.46 waves: 0.618 / 0.465; 8 envs (shared memory then
    // allows 8 CTAs per SM): slower. One env more or less per CTA when that fills the last wave clearly better.
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    }
    const int nA = D.Nh + (D.orca.robot_visible ? 1 : 0);
    const bool cut = !D.orca_dedup && dre_orca_use_cutoff(D.orca);
    auto fill = [&](int epb) {
        const size_t smem = ha_layout(D.Nh, nA, NG, epb, cut).bytes + 1024;     // + the per-CTA reservation
        int per_sm = (int)((size_t)227 * 1024 / smem);
        if (per_sm > DRE_HA_MINB) per_sm = DRE_HA_MINB;                        // the register bound (__launch_bounds__)
        if (per_sm < 1) per_sm = 1;
        const double waves = (double)((D.E + epb - 1) / epb) / ((double)sms * per_sm);
        return waves <= 1.0 ? 1.0 : waves / ceil(waves);
    };
    int best = base;
    double bf = fill(base);
    const int cand[2] = {base + 1, base - 1};
    for (int c : cand) {
        if (c < 1 || c > cap) continue;
        const double f = fill(c);
        if (f > bf + 0.05) { best = c; bf = f; }
    }
    return best;
}

template <int G, bool DEDUP, bool CUTOFF, bool PRUNE>
static int launch_ha(const EnvDev &D, int mode, int write_obs, cudaStream_t s)
{
    const int Nh = D.Nh, nA = Nh + (D.orca.robot_visible ? 1 : 0);
    // 128 threads = 128/G lane-groups and ~8 agents per group so the dynamic schedule can even out LP2 / LP3 solves; measured
    // on B200 at 16384 x 20: 128 threads 0.66 ms, 256 threads 0.71 ms, 512 threads 0.88 ms (CTA barriers cost less when more,
    // smaller CTAs share an SM)
    const int threads = dre_ha_threads();
    const int EPB = dre_env_ha_envs_per_cta(D);
    HaLayout lay = ha_layout(Nh, nA, threads / G, EPB, CUTOFF);
    {   // DRE_HA_SMEM_PAD_KB: extra dynamic shared memory per CTA (measurement aid: caps the CTAs per SM)
        static int pad = -1;
        if (pad < 0) { const char *e = getenv("DRE_HA_SMEM_PAD_KB"); pad = e ? atoi(e) : 0; }
        lay.bytes += (size_t)pad * 1024;
    }
    auto kern = ha_step_kernel<G, DEDUP, CUTOFF, PRUNE>;
    if (lay.bytes > 48 * 1024) {
        if (lay.bytes > 220 * 1024) return DRE_ERR_INVALID;
        DRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    }
    kern<<<(D.E + EPB - 1) / EPB, threads, lay.bytes, s>>>(D, EPB, mode, write_obs, lay);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_env_launch_ha(const EnvDev &D, int mode, int write_obs, cudaStream_t s)
{
    const int G = dre_env_pick_group(D);
    const bool cut = dre_orca_use_cutoff(D.orca);   // short neighbour range: cut-off + compaction before the ranking (orca_device.cuh)
    // kernel variants (one instantiation each, so the default two-pass kernel carries none of the others' code):
    // dedup | look-ahead pruning (+ cut-off) | cut-off | default
#define DRE_HA_CASE(g) case g: return D.orca_dedup ? launch_ha<g, true, false, false>(D, mode, write_obs, s) \
                                   : D.lookahead_prune ? (cut ? launch_ha<g, false, true, true>(D, mode, write_obs, s) : launch_ha<g, false, false, true>(D, mode, write_obs, s)) \
                                   : (cut ? launch_ha<g, false, true, false>(D, mode, write_obs, s) : launch_ha<g, false, false, false>(D, mode, write_obs, s))
    switch (G) {
    DRE_HA_CASE(2);
    DRE_HA_CASE(4);
    DRE_HA_CASE(8);
    DRE_HA_CASE(16);
    default:
    DRE_HA_CASE(32);
    }
#undef DRE_HA_CASE
}

// float64 -> float32, one rounding (the opt-in float32 observation output of the host path)
__global__ void cvt32_kernel(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const double2 v = *reinterpret_cast<const double2 *>(src + i);
        *reinterpret_cast<float2 *>(dst + i) = make_float2((float)v.x, (float)v.y);
    } else if (i < n) dst[i] = (float)src[i];
}

int dre_env_launch_cvt32(const double *src, float *dst, size_t n, cudaStream_t s)
{
    if (n == 0) return DRE_OK;
    const size_t pairs = (n + 1) / 2;
    cvt32_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(src, dst, n);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_env_launch_pt(const EnvDev &D, int mode, cudaStream_t s)
{
    const int threads = 128, envs_per_block = threads / 8;   // PT_LANES lanes per env
    pt_step_kernel<<<(D.E + envs_per_block - 1) / envs_per_block, threads, 0, s>>>(D, mode);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_env_launch_mpc_stats(const EnvDev &D, cudaStream_t s)
{
    if (!D.stat_counters) return DRE_OK;
    mpc_stats_kernel<<<(D.E + 255) / 256, 256, 0, s>>>(D);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// HAAndPTEnv.soft_reset's path switch (HA_and_PT env.py:130-141) for the masked envs: next path, localise to its start,
// MPC.reset(). The caller regenerates S with dre_env_observe_masked(mask, solve_mpc = 1) like PT_env.soft_reset (:141-148).
__global__ void switch_path_kernel(EnvDev D)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.E || (D.env_mask != nullptr && D.env_mask[e] == 0)) return;
    D.path_idx[e] = (D.path_idx[e] + 1) % D.num_paths;
    D.loc_to_start[e] = 1;
    D.mpc_iteration0[e] = 1;
}

int dre_env_launch_switch_path(const EnvDev &D, cudaStream_t s)
{
    switch_path_kernel<<<(D.E + 127) / 128, 128, 0, s>>>(D);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// The soft-reset machine's decision for the NEXT step of every env, from the latest step's done word: what a host-side
// caller that drives HAAndPTEnv.soft_reset's loop itself (one self.step(action, soft_reset=True) per decision) needs.
__global__ void sr_decide_kernel(EnvDev D, int commit, uint8_t *__restrict__ scripted, double *__restrict__ actions, uint8_t *__restrict__ stuck_out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.E) return;
    SrState S = sr_unpack(D.sr_state + (size_t)e * 4);
    const PathView P = path_view(D.paths, D.path_stride, D.path_idx[e], D.lens);
    double av = 0.0, aw = 0.0;
    bool stuck = false;
    const int sc = soft_reset_decide(S, D.out.done_bits[e], P, D.rpose[(size_t)e * 4], D.rpose[(size_t)e * 4 + 1], D.rpose[(size_t)e * 4 + 2],
                                     D.loc_to_start[e] != 0, av, aw, stuck);
    if (commit) sr_pack(S, D.sr_state + (size_t)e * 4);
    scripted[e] = (uint8_t)sc;
    actions[(size_t)e * 2] = av; actions[(size_t)e * 2 + 1] = aw;
    if (stuck_out) stuck_out[e] = stuck ? 1 : 0;
}

int dre_env_launch_sr_decide(const EnvDev &D, int commit, uint8_t *scripted, double *actions, uint8_t *stuck, cudaStream_t s)
{
    sr_decide_kernel<<<(D.E + 127) / 128, 128, 0, s>>>(D, commit, scripted, actions, stuck);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_env_launch_reset(const EnvDev &D, const double *hpos_init, const double *hgoal_init, cudaStream_t s)
{
    env_reset_kernel<<<(D.E + 127) / 128, 128, 0, s>>>(D, (const double2 *)hpos_init, (const double2 *)hgoal_init);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}
