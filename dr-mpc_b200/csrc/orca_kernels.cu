// orca_kernels.cu -- stand-alone batched ORCA pass (replaces CrowdSim.get_human_actions,
// reference crowd_sim.py:264-284 -> orca.py:64-117). Compile with -fmad=false (see orca_device.cuh).
//
// Grid: one CTA per EPB consecutive envs. The CTA stages its envs' agents (humans + robot) in shared
// memory as an fp32 structure-of-arrays tile (the fp64 -> fp32 cast is the Cython boundary of the
// reference), then every lane-group of G lanes solves one human's LP against the tile.
#include <cstdlib>
#include "orca_device.cuh"
#include "launch.cuh"

struct OrcaTileLayout {
    int nA, NG, EPB;
    size_t off_lines, off_agents, off_dist, bytes;
};

static OrcaTileLayout orca_layout(int Nh, int robot_visible, int G, int threads, int EPB, bool cutoff)
{
    OrcaTileLayout l;
    l.nA = Nh + (robot_visible ? 1 : 0);
    l.NG = threads / G;
    l.EPB = EPB;
    l.off_lines = 0;
    l.off_agents = (size_t)l.NG * 2 * l.nA * sizeof(OLine);
    l.off_dist = (l.off_agents + (size_t)EPB * 5 * l.nA * sizeof(float) + 15) / 16 * 16;
    l.bytes = (l.off_dist + (size_t)l.NG * (cutoff ? 2 : 1) * ((l.nA + 3) & ~3) * sizeof(float) + 15) / 16 * 16 + 16;   // distances (+ indices), + work counter
    return l;
}

template <int G, bool CUTOFF>
__global__ void __launch_bounds__(128, 9)
orca_predict_kernel(dre_orca_params p, int E, int Nh, int EPB, const double2 *__restrict__ hpos,
                    const float2 *__restrict__ hvel, const double2 *__restrict__ goal,
                    const double2 *__restrict__ rpos, const uint8_t *__restrict__ is_static,
                    float2 *__restrict__ vnew, dre_orca_stats *__restrict__ stats,
                    size_t off_agents, size_t off_dist, size_t off_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nA = Nh + (p.robot_visible ? 1 : 0);
    const int NG = blockDim.x / G;
    OLine *lines = reinterpret_cast<OLine *>(smem_raw);
    float *agents = reinterpret_cast<float *>(smem_raw + off_agents);
    float *dist = reinterpret_cast<float *>(smem_raw + off_dist);
    int *work_counter = reinterpret_cast<int *>(smem_raw + off_counter);
    const int env0 = blockIdx.x * EPB;
    if (threadIdx.x == 0) *work_counter = 0;

    // stage the tile: agents[e_local][field][a]
    for (int t = threadIdx.x; t < EPB * nA; t += blockDim.x) {
        const int el = t / nA, a = t - el * nA, e = env0 + el;
        float *A = agents + (size_t)el * 5 * nA;
        float x = 0.f, y = 0.f, vx = 0.f, vy = 0.f;
        if (e < E) {
            if (a < Nh) {
                const double2 q = hpos[(size_t)e * Nh + a];
                const float2 v = hvel[(size_t)e * Nh + a];
                x = (float)q.x; y = (float)q.y; vx = v.x; vy = v.y;
            } else {
                const double2 q = rpos[e];
                x = (float)q.x; y = (float)q.y;   // robot seen at rest (agent.py:43,96-97)
            }
        }
        A[a] = x; A[nA + a] = y; A[2 * nA + a] = vx; A[3 * nA + a] = vy; A[4 * nA + a] = p.radius;
    }
    __syncthreads();

    cg::thread_block blk = cg::this_thread_block();
    TileGroup<G> g(cg::tiled_partition<G>(blk));
    const int gi = threadIdx.x / G;
    OLine *L = lines + (size_t)gi * 2 * nA;
    const int nA4 = (nA + 3) & ~3;
    float *sd = dist + (size_t)gi * (CUTOFF ? 2 : 1) * nA4;
    const float invTH = 1.0f / p.time_horizon, invDt = 1.0f / p.time_step;
    const int maxNb = p.max_neighbors > 0 ? p.max_neighbors : nA - 1;

    const bool static_sched = (p.group_lanes & 0x200) != 0;   // measurement aid: agent = group + k * NG instead of the queue
    int next_static = gi;
    for (;;) {   // groups pull agents from a CTA-wide counter: LP3 solves are ~10x dearer than LP2 ones
        int ag = 0;
        if (static_sched) { ag = next_static; next_static += NG; }
        else {
            if (g.rank() == 0) ag = atomicAdd(work_counter, 1);
            ag = g.bcast(ag, 0);
        }
        if (ag >= EPB * Nh) break;
        const int el = ag / Nh, i = ag - el * Nh, e = env0 + el;
        if (e >= E) continue;
        const size_t gidx = (size_t)e * Nh + i;
        float ox = 0.f, oy = 0.f;
        if (is_static != nullptr && is_static[gidx]) {
            if (g.rank() == 0 && stats) { dre_orca_stats z = {0, 0, 0, 0, 0u, 0}; stats[gidx] = z; }
        } else {
            // preferred velocity in fp64, then cast (orca.py:97-100,108)
            const double2 P = hpos[gidx], Gl = goal[gidx];
            const double dx = Gl.x - P.x, dy = Gl.y - P.y;
            const double speed = sqrt(dx * dx + dy * dy);
            double pvx = dx, pvy = dy;
            if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
            const float *A = agents + (size_t)el * 5 * nA;
            orca_solve_group<CUTOFF>(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, (float)pvx, (float)pvy,
                             p.max_speed_self, p.neighbor_dist, maxNb, invTH, invDt, sd, sd + nA4, L, ox, oy,
                             stats ? &stats[gidx] : nullptr, (p.group_lanes & 0x100) != 0);
        }
        if (g.rank() == 0) vnew[gidx] = make_float2(ox, oy);
        g.sync();
    }
}

template <int G, bool CUTOFF>
static int launch_orca(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                       const double *rpos, const uint8_t *is_static, float *vnew, dre_orca_stats *stats, cudaStream_t s)
{
    // same CTA shape as the fused env kernel (env_kernels.cu: launch_ha): 128 threads, ~8 agents per lane-group;
    // DRE_ORCA_THREADS / DRE_ORCA_AGENTS_PER_GROUP override it for A/B measurements
    static int threads = 0, apg = 0;
    if (threads == 0) {
        const char *e = getenv("DRE_ORCA_THREADS"); threads = e ? atoi(e) : 128;
        if (threads != 64 && threads != 128) threads = 128;   // the kernel is bounded to 128 threads
        e = getenv("DRE_ORCA_AGENTS_PER_GROUP"); apg = e ? atoi(e) : 8;
        if (apg < 1 || apg > 64) apg = 8;
    }
    const int NGl = threads / G;
    int EPB = (apg * NGl) / Nh;
    if (EPB > (E + 591) / 592) EPB = (E + 591) / 592;   // keep >= 4 CTAs per SM in flight for small batches
    if (EPB < 1) EPB = 1;
    const OrcaTileLayout l = orca_layout(Nh, p.robot_visible, G, threads, EPB, CUTOFF);
    auto kern = orca_predict_kernel<G, CUTOFF>;
    if (l.bytes > 48 * 1024) {
        if (l.bytes > 220 * 1024) return DRE_ERR_INVALID;
        DRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.bytes));
    }
    const int blocks = (E + EPB - 1) / EPB;
    kern<<<blocks, threads, l.bytes, s>>>(p, E, Nh, EPB, (const double2 *)hpos, (const float2 *)hvel, (const double2 *)goal,
                                           (const double2 *)rpos, is_static, (float2 *)vnew, stats, l.off_agents, l.off_dist, l.bytes - 16);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// ---- thread-per-agent variant (group_lanes = 1): one thread solves one agent with the sequential RVO2 control flow; the
// per-thread scratch arrays are interleaved across the CTA's threads (conflict-free 16 B accesses). P_GLOBAL: LP3's
// projected lines live in a global scratch buffer instead of shared memory (halves the footprint -> 4 CTAs per SM).
#define DRE_TP_THREADS 128
template <bool P_GLOBAL>
__global__ void __launch_bounds__(DRE_TP_THREADS)
orca_thread_kernel(dre_orca_params p, int E, int Nh, int EPB, const double2 *__restrict__ hpos,
                   const float2 *__restrict__ hvel, const double2 *__restrict__ goal,
                   const double2 *__restrict__ rpos, const uint8_t *__restrict__ is_static,
                   float2 *__restrict__ vnew, dre_orca_stats *__restrict__ stats, OLine *__restrict__ pscratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nA = Nh + (p.robot_visible ? 1 : 0), nA4 = (nA + 3) & ~3;
    OLine *lines = reinterpret_cast<OLine *>(smem_raw);                                   // [nA or 2 nA][threads]
    const size_t nl = (size_t)(P_GLOBAL ? nA : 2 * nA) * DRE_TP_THREADS;
    float4 *dist = reinterpret_cast<float4 *>(smem_raw + nl * sizeof(OLine));              // [2][nA4/4][threads]: distances, indices
    float *agents = reinterpret_cast<float *>(smem_raw + nl * sizeof(OLine) + (size_t)2 * (nA4 / 4) * DRE_TP_THREADS * sizeof(float4));
    const int env0 = blockIdx.x * EPB;
    for (int t = threadIdx.x; t < EPB * nA; t += blockDim.x) {
        const int el = t / nA, a = t - el * nA, e = env0 + el;
        float *A = agents + (size_t)el * 5 * nA;
        float x = 0.f, y = 0.f, vx = 0.f, vy = 0.f;
        if (e < E) {
            if (a < Nh) {
                const double2 q = hpos[(size_t)e * Nh + a];
                const float2 v = hvel[(size_t)e * Nh + a];
                x = (float)q.x; y = (float)q.y; vx = v.x; vy = v.y;
            } else {
                const double2 q = rpos[e];
                x = (float)q.x; y = (float)q.y;
            }
        }
        A[a] = x; A[nA + a] = y; A[2 * nA + a] = vx; A[3 * nA + a] = vy; A[4 * nA + a] = p.radius;
    }
    __syncthreads();
    const int ag = threadIdx.x;
    if (ag >= EPB * Nh) return;
    const int el = ag / Nh, i = ag - el * Nh, e = env0 + el;
    if (e >= E) return;
    const size_t gidx = (size_t)e * Nh + i;
    float ox = 0.f, oy = 0.f;
    if (is_static != nullptr && is_static[gidx]) {
        if (stats) { dre_orca_stats z = {0, 0, 0, 0, 0u, 0}; stats[gidx] = z; }
    } else {
        const double2 P = hpos[gidx], Gl = goal[gidx];
        const double dx = Gl.x - P.x, dy = Gl.y - P.y;
        const double speed = sqrt(dx * dx + dy * dy);
        double pvx = dx, pvy = dy;
        if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
        const float *A = agents + (size_t)el * 5 * nA;
        const float invTH = 1.0f / p.time_horizon, invDt = 1.0f / p.time_step;
        const int maxNb = p.max_neighbors > 0 ? p.max_neighbors : nA - 1;
        StridedDist sd = {dist + threadIdx.x, DRE_TP_THREADS};
        StridedDist si = {dist + (size_t)(nA4 / 4) * DRE_TP_THREADS + threadIdx.x, DRE_TP_THREADS};
        SerialGroup g;
        if (P_GLOBAL) {
            // L in shared memory, P in the global scratch: one accessor type -> generic addressing for both
            TwoSpaceLines L = {lines + threadIdx.x, DRE_TP_THREADS,
                               pscratch + (size_t)blockIdx.x * nA * DRE_TP_THREADS + threadIdx.x, DRE_TP_THREADS, nA, 0};
            orca_solve_group(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, (float)pvx, (float)pvy, p.max_speed_self,
                             p.neighbor_dist, maxNb, invTH, invDt, sd, si, L, ox, oy, stats ? &stats[gidx] : nullptr, false);
        } else {
            StridedLines L = {lines + threadIdx.x, DRE_TP_THREADS};
            orca_solve_group(g, nA, A, A + nA, A + 2 * nA, A + 3 * nA, A + 4 * nA, i, (float)pvx, (float)pvy, p.max_speed_self,
                             p.neighbor_dist, maxNb, invTH, invDt, sd, si, L, ox, oy, stats ? &stats[gidx] : nullptr, false);
        }
    }
    vnew[gidx] = make_float2(ox, oy);
}

static int launch_orca_thread(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                              const double *rpos, const uint8_t *is_static, float *vnew, dre_orca_stats *stats, cudaStream_t s)
{
    const int nA = Nh + (p.robot_visible ? 1 : 0), nA4 = (nA + 3) & ~3;
    int EPB = DRE_TP_THREADS / Nh;
    if (EPB < 1) return DRE_ERR_INVALID;
    const size_t fixed = (size_t)2 * (nA4 / 4) * DRE_TP_THREADS * sizeof(float4) + (size_t)EPB * 5 * nA * sizeof(float);
    // large crowds: both line arrays do not fit in shared memory -> projected lines in the global scratch
    const bool pglobal = (p.group_lanes & 0x400) != 0 || (size_t)2 * nA * DRE_TP_THREADS * sizeof(OLine) + fixed > 200 * 1024;
    const size_t bytes = (size_t)(pglobal ? nA : 2 * nA) * DRE_TP_THREADS * sizeof(OLine) + fixed;
    if (bytes > 220 * 1024) return DRE_ERR_INVALID;
    const int blocks = (E + EPB - 1) / EPB;
    static OLine *scratch = nullptr; static size_t scratch_n = 0;
    if (pglobal) {
        const size_t need = (size_t)blocks * nA * DRE_TP_THREADS;
        if (need > scratch_n) { if (scratch) cudaFree(scratch); DRE_CUDA_TRY(cudaMalloc(&scratch, need * sizeof(OLine))); scratch_n = need; }
        DRE_CUDA_TRY(cudaFuncSetAttribute(orca_thread_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        orca_thread_kernel<true><<<blocks, DRE_TP_THREADS, bytes, s>>>(p, E, Nh, EPB, (const double2 *)hpos, (const float2 *)hvel, (const double2 *)goal,
                                                                        (const double2 *)rpos, is_static, (float2 *)vnew, stats, scratch);
    } else {
        DRE_CUDA_TRY(cudaFuncSetAttribute(orca_thread_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        orca_thread_kernel<false><<<blocks, DRE_TP_THREADS, bytes, s>>>(p, E, Nh, EPB, (const double2 *)hpos, (const float2 *)hvel, (const double2 *)goal,
                                                                         (const double2 *)rpos, is_static, (float2 *)vnew, stats, nullptr);
    }
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Joint-state solve: ORCA.predict(JointState) (reference scripts/policy/orca.py:64-117) for a batch of INDEPENDENT joint
// states -- agent 0 = the ego agent (own radius, own max speed, preferred velocity), agents 1..n = the others in the order
// of state.human_states, each with its own radius -- i.e. exactly what one private rvo2 simulator sees. A lane-group
// stages its joint state as a private tile and solves agent 0. This is the literal boundary-B entry point; the crowd
// pass above is the same solve with the tile shared by an env's humans.
// ---------------------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(128)
orca_joint_kernel(dre_orca_params p, int N, int n_others, const float *__restrict__ self8, const float *__restrict__ others5,
                  float2 *__restrict__ vnew, dre_orca_stats *__restrict__ stats, int *__restrict__ queue)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nA = n_others + 1, nA4 = (nA + 3) & ~3;
    const int NG = blockDim.x / G, gi = threadIdx.x / G;
    const size_t per_group = (size_t)2 * nA * sizeof(OLine) + (size_t)(5 * nA4 + 2 * nA4) * sizeof(float);
    unsigned char *base = smem_raw + per_group * gi;
    OLine *L = reinterpret_cast<OLine *>(base);
    float *A = reinterpret_cast<float *>(base + (size_t)2 * nA * sizeof(OLine));   // [5][nA4]
    float *sd = A + 5 * nA4;
    cg::thread_block blk = cg::this_thread_block();
    TileGroup<G> g(cg::tiled_partition<G>(blk));
    const float invTH = 1.0f / p.time_horizon, invDt = 1.0f / p.time_step;
    const int maxNb = p.max_neighbors > 0 ? p.max_neighbors : n_others;
    (void)NG;
    for (;;) {
        int n = 0;
        if (g.rank() == 0) n = atomicAdd(queue, 1);
        n = g.bcast(n, 0);
        if (n >= N) break;
        const float *S = self8 + (size_t)n * 8;
        const float *O = others5 + (size_t)n * n_others * 5;
        for (int a = g.rank(); a < nA; a += G) {
            float x, y, vx, vy, r;
            if (a == 0) { x = S[0]; y = S[1]; vx = S[2]; vy = S[3]; r = S[6]; }
            else { const float *o = O + (size_t)(a - 1) * 5; x = o[0]; y = o[1]; vx = o[2]; vy = o[3]; r = o[4]; }
            A[a] = x; A[nA4 + a] = y; A[2 * nA4 + a] = vx; A[3 * nA4 + a] = vy; A[4 * nA4 + a] = r;
        }
        g.sync();
        float ox = 0.f, oy = 0.f;
        orca_solve_group(g, nA, A, A + nA4, A + 2 * nA4, A + 3 * nA4, A + 4 * nA4, 0, S[4], S[5], S[7], p.neighbor_dist, maxNb,
                         invTH, invDt, sd, sd + nA4, L, ox, oy, stats ? &stats[n] : nullptr, false);
        if (g.rank() == 0) vnew[n] = make_float2(ox, oy);
        g.sync();
    }
}

template <int G>
static int launch_orca_joint(const dre_orca_params &p, int N, int n_others, const float *self8, const float *others5, float *vnew,
                             dre_orca_stats *stats, int *queue, cudaStream_t s)
{
    const int nA = n_others + 1, nA4 = (nA + 3) & ~3, threads = 128;
    const size_t bytes = ((size_t)2 * nA * sizeof(OLine) + (size_t)7 * nA4 * sizeof(float)) * (threads / G);
    if (bytes > 220 * 1024) return DRE_ERR_INVALID;
    auto kern = orca_joint_kernel<G>;
    if (bytes > 48 * 1024) DRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int blocks = (N + threads / G - 1) / (threads / G);
    if (blocks > 148 * 8) blocks = 148 * 8;
    DRE_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int), s));
    kern<<<blocks, threads, bytes, s>>>(p, N, n_others, self8, others5, (float2 *)vnew, stats, queue);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_orca_joint_launch(const dre_orca_params &p, int N, int n_others, const float *self8, const float *others5, float *vnew,
                          dre_orca_stats *stats, int *queue, cudaStream_t s)
{
    if (N <= 0 || n_others < 0 || n_others > DRE_MAX_HUMANS) return DRE_ERR_INVALID;
    if (n_others + 1 <= 8) return launch_orca_joint<4>(p, N, n_others, self8, others5, vnew, stats, queue, s);
    return launch_orca_joint<8>(p, N, n_others, self8, others5, vnew, stats, queue, s);
}

// Short neighbour range (below the reference crowd's radius of 5 m): most candidates are out of range, the cut-off +
// compaction variant of the neighbour search wins; group_lanes bits 0x800 / 0x1000 force it on / off (A/B measurements).
bool dre_orca_use_cutoff(const dre_orca_params &p)
{
    if (p.group_lanes & 0x800) return true;
    if (p.group_lanes & 0x1000) return false;
    return p.neighbor_dist < 5.0f;
}

int dre_orca_pick_group(const dre_orca_params &p, int Nh)
{
    const int gl = p.group_lanes & 0xff;
    if (gl == 1) return 1;
    if (gl == 2 || gl == 4 || gl == 8 || gl == 16 || gl == 32) return gl;
    if (p.group_lanes == 2 || p.group_lanes == 4 || p.group_lanes == 8 || p.group_lanes == 16 || p.group_lanes == 32) return p.group_lanes;
    const int nA = Nh + (p.robot_visible ? 1 : 0);
    return nA <= 8 ? 4 : 8;   // measured on B200: 8 lanes per agent is fastest from 10 to 50 humans (profiles/)
}

int dre_orca_launch(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                    const double *rpos, const uint8_t *is_static, float *vnew, dre_orca_stats *stats, cudaStream_t s)
{
    if (E <= 0 || Nh <= 0 || Nh > DRE_MAX_HUMANS) return DRE_ERR_INVALID;
    const bool cut = dre_orca_use_cutoff(p);
    switch (dre_orca_pick_group(p, Nh)) {
    case 1: return launch_orca_thread(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    case 2: return cut ? launch_orca<2, true>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s) : launch_orca<2, false>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    case 4: return cut ? launch_orca<4, true>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s) : launch_orca<4, false>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    case 8: return cut ? launch_orca<8, true>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s) : launch_orca<8, false>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    case 16: return cut ? launch_orca<16, true>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s) : launch_orca<16, false>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    default: return cut ? launch_orca<32, true>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s) : launch_orca<32, false>(p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, s);
    }
}
