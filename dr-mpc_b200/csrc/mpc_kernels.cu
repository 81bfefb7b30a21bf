// mpc_kernels.cu -- batched MPC.act (reference environment/path_tracking/utils/mpc.py:58-206).
//
// Two kernels:
//   mpc_coop_kernel<G> : a group of G >= K lanes per problem (4, 8, 16, 32, or a pair of warps for K > 32), lane k = horizon
//                        step k (mpc_coop.cuh). This is the product path for every K <= DRE_MAX_HORIZON.
//   mpc_act_kernel<KC> : one thread per problem with the system in local memory: DRE_MPC_SERIAL=1 only (A/B measurements,
//                        and the host-compilable semantic reference of the tests).
// Data layout: warm-start poses/controls are problem-major ([N][K] double4 / double2), so the lanes of a warp
// (G-lane groups of consecutive problems) read/write one contiguous run; per-problem inputs and outputs likewise.
#include <cstdlib>
#include <cstdio>
#include "mpc_coop.cuh"
#include "launch.cuh"

// MINB = minimum resident CTAs per SM the register allocation is capped for. The linearisation lives in shared memory
// (mpc_coop.cuh), which lets 3 CTAs = 12 warps per SM hide the fp64 dependency chains without spills.
template <int G, int MINB, bool TRACE = false>
__global__ void __launch_bounds__(DRE_MPC_CTA, MINB * (128 / DRE_MPC_CTA))
mpc_coop_kernel(dre_mpc_params prm, CoopBatch B)
{
    extern __shared__ __align__(16) double coop_smem[];   // CS_SLOTS columns of DRE_MPC_CTA doubles
    LaneGroup<G> g;
    g.attach(coop_smem + CS_SLOTS * DRE_MPC_CTA);   // exchange area of the two-warp groups (K > 32)
    mpc_coop_persistent<G, TRACE>(g, prm, B, coop_smem);
}

// DRE_MPC_TRACE=1: run the instrumented K <= 4 kernel and print where a lane-group's cycles go (stderr; synchronises)
static bool dre_mpc_trace()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DRE_MPC_TRACE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// DRE_MPC_MINB=2|3|4 overrides the occupancy target of the K <= 4 kernel (A/B measurements)
static int dre_mpc_minb()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DRE_MPC_MINB"); v = e ? atoi(e) : 0; }
    return v;
}

template <int G, int MINB>
static int launch_mpc_coop(const dre_mpc_params &p, const MpcDeviceState &st, const int32_t *path_id, const double *interp_index,
                           const double *pose, int pose_stride, double *controls, int controls_stride, int controls_count,
                           dre_mpc_status *status, const uint8_t *active_mask, cudaStream_t s)
{
    const int threads = DRE_MPC_CTA;
    const size_t smem = ((size_t)CS_SLOTS * DRE_MPC_CTA + (G == 64 ? 64 : 0)) * sizeof(double);
    static int resident_blocks = 0;           // persistent grid: as many CTAs as fit on the device at once
    if (resident_blocks == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        DRE_CUDA_TRY(cudaGetDevice(&dev));
        DRE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        DRE_CUDA_TRY(cudaFuncSetAttribute(mpc_coop_kernel<G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (G == 4 && MINB == 3)
            DRE_CUDA_TRY(cudaFuncSetAttribute(mpc_coop_kernel<4, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpc_coop_kernel<G, MINB>, threads, smem));
        // DRE_MPC_CTAS_PER_SM caps the persistent grid (measurement aid: leaves register file to a co-running kernel)
        const char *cap = getenv("DRE_MPC_CTAS_PER_SM");
        if (cap && atoi(cap) > 0 && atoi(cap) < per_sm) per_sm = atoi(cap);
        resident_blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    // small batches: one problem per warp rather than filling the warps of a few CTAs (see the fetch in mpc_coop.cuh)
    const int warps_per_block = G > 32 ? threads / G : threads / 32;
    int blocks = (st.N + warps_per_block - 1) / warps_per_block;
    if (blocks > resident_blocks) blocks = resident_blocks;
    CoopBatch B;
    B.N = st.N; B.path_id = path_id; B.interp_index = interp_index; B.pose = pose; B.pose_stride = pose_stride;
    B.controls = controls; B.controls_stride = controls_stride; B.controls_count = controls_count; B.status = status;
    B.active_mask = active_mask; B.paths = st.paths; B.lens = st.lens; B.num_paths = st.num_paths; B.path_stride = st.path_stride;
    B.warm_T = reinterpret_cast<double4 *>(st.warm_T); B.warm_u = reinterpret_cast<double2 *>(st.warm_u);
    B.iteration0 = st.iteration0; B.queue = st.queue;
    DRE_CUDA_TRY(cudaMemsetAsync(st.queue, 0, sizeof(int), s));
    B.trace = nullptr;
    if (G == 4 && MINB == 3 && dre_mpc_trace()) {
        static unsigned long long *d_tr = nullptr;
        if (!d_tr) DRE_CUDA_TRY(cudaMalloc(&d_tr, 16 * sizeof(unsigned long long)));
        DRE_CUDA_TRY(cudaMemsetAsync(d_tr, 0, 16 * sizeof(unsigned long long), s));
        B.trace = d_tr;
        mpc_coop_kernel<4, 3, true><<<blocks, threads, smem, s>>>(p, B);
        DRE_CUDA_TRY(cudaGetLastError());
        unsigned long long h[16];
        DRE_CUDA_TRY(cudaMemcpyAsync(h, d_tr, sizeof h, cudaMemcpyDeviceToHost, s));
        DRE_CUDA_TRY(cudaStreamSynchronize(s));
        const double it = (double)(h[6] ? h[6] : 1), gr = (double)(h[9] ? h[9] : 1);
        fprintf(stderr, "[mpc trace] groups %llu iters/group %.1f (max %llu) live groups/warp %.2f | cycles per iteration: fetch %.0f cost0 %.0f "
                        "build %.0f solve %.0f trials %.0f book %.0f | longest group %.0f us at 1.965 GHz\n",
                h[9], it / gr, h[8], (double)h[7] / it, h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[4] / it, h[5] / it, h[10] / 1965.0);
        return DRE_OK;
    }
    mpc_coop_kernel<G, MINB><<<blocks, threads, smem, s>>>(p, B);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

template <int KC>
__global__ void __launch_bounds__(128)
mpc_act_kernel(dre_mpc_params prm, MpcDeviceState st, const int32_t *__restrict__ path_id,
               const double *__restrict__ interp_index, const double *__restrict__ pose, int pose_stride,
               double *__restrict__ controls, int controls_stride, int controls_count,
               dre_mpc_status *__restrict__ status, const uint8_t *__restrict__ active_mask)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= st.N) return;
    if (active_mask != nullptr && !active_mask[n]) return;
    int pid = path_id ? path_id[n] : 0;
    if (pid < 0 || pid >= st.num_paths) pid = 0;
    const double *path = st.paths + (size_t)pid * 3 * st.path_stride;
    MpcSystem<KC> S;
    mpc_act_one<KC>(prm, path, st.path_stride, st.lens[pid], interp_index[n], pose[(size_t)n * pose_stride],
                    pose[(size_t)n * pose_stride + 1], pose[(size_t)n * pose_stride + 2],
                    reinterpret_cast<double4 *>(st.warm_T) + (size_t)n * prm.horizon,
                    reinterpret_cast<double2 *>(st.warm_u) + (size_t)n * prm.horizon, (size_t)1,
                    st.iteration0 + n, controls + (size_t)n * controls_stride, controls_count,
                    status ? status + n : nullptr, S);
}

__global__ void mpc_reset_kernel(uint8_t *it0, const uint8_t *mask, int N)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N && (mask == nullptr || mask[n])) it0[n] = 1;
}

template <int KC>
static int launch_mpc(const dre_mpc_params &p, const MpcDeviceState &st, const int32_t *path_id, const double *interp_index,
                      const double *pose, int pose_stride, double *controls, int controls_stride, int controls_count,
                      dre_mpc_status *status, const uint8_t *active_mask, cudaStream_t s)
{
    const int threads = KC <= 4 ? 64 : 32;   // small CTAs spread few problems over all 148 SMs
    const int blocks = (st.N + threads - 1) / threads;
    mpc_act_kernel<KC><<<blocks, threads, 0, s>>>(p, st, path_id, interp_index, pose, pose_stride, controls, controls_stride,
                                                  controls_count, status, active_mask);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

// DRE_MPC_SERIAL=1 selects the one-thread-per-problem kernel for every K (A/B measurements, tests)
static bool dre_mpc_force_serial()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DRE_MPC_SERIAL"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int dre_mpc_launch(const dre_mpc_params &p, const MpcDeviceState &st, const int32_t *path_id, const double *interp_index,
                   const double *pose, int pose_stride, double *controls, int controls_stride, int controls_count,
                   dre_mpc_status *status, const uint8_t *active_mask, cudaStream_t s)
{
    const int K = p.horizon;
    if (K < 1 || K > DRE_MAX_HORIZON || st.N <= 0 || st.paths == nullptr) return DRE_ERR_INVALID;
#define DRE_MPC_ARGS p, st, path_id, interp_index, pose, pose_stride, controls, controls_stride, controls_count, status, active_mask, s
    if (!dre_mpc_force_serial()) {
        if (K <= 4) {
            switch (dre_mpc_minb()) {
            case 2: return launch_mpc_coop<4, 2>(DRE_MPC_ARGS);
            case 4: return launch_mpc_coop<4, 4>(DRE_MPC_ARGS);
            default: return launch_mpc_coop<4, 3>(DRE_MPC_ARGS);
            }
        }
        if (K <= 8) return launch_mpc_coop<8, 3>(DRE_MPC_ARGS);
        if (K <= 16) return launch_mpc_coop<16, 3>(DRE_MPC_ARGS);
        if (K <= 32) return launch_mpc_coop<32, 3>(DRE_MPC_ARGS);
        return launch_mpc_coop<64, 4>(DRE_MPC_ARGS);   // one problem per pair of warps; 16 warps per SM (measured +11 % at K = 40 despite the spills)
    }
    if (K <= 4) return launch_mpc<4>(DRE_MPC_ARGS);
    if (K <= 10) return launch_mpc<10>(DRE_MPC_ARGS);
    if (K <= 20) return launch_mpc<20>(DRE_MPC_ARGS);
    return launch_mpc<40>(DRE_MPC_ARGS);
#undef DRE_MPC_ARGS
}

// accuracy self-test of the kernel's own reciprocal (tests/test_gpu_mpc.py)
__global__ void mpc_math_selftest_kernel(int n, const double *__restrict__ x, double *__restrict__ rc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rc[i] = rcp_fast(x[i]);
}

// fn: 0 = rcp_fast, 1 = barrier_val_fast (compact ln with the reference's clamps), 2 / 3 = sin / cos of sincos_small
__global__ void mpc_math_selftest_fn_kernel(int fn, int n, const double *__restrict__ x, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sn, cs;
    switch (fn) {
    case 0: out[i] = rcp_fast(x[i]); break;
    case 1: out[i] = barrier_val_fast(x[i]); break;
    default: sincos_small(x[i], &sn, &cs); out[i] = fn == 2 ? sn : cs; break;
    }
}

extern "C" int dre_mpc_math_selftest_fn(int32_t fn, int32_t n, const double *x_dev, double *out_dev, void *stream)
{
    if (fn < 0 || fn > 3 || n <= 0 || !x_dev || !out_dev) return DRE_ERR_INVALID;
    mpc_math_selftest_fn_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fn, n, x_dev, out_dev);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

extern "C" int dre_mpc_math_selftest(int32_t n, const double *x_dev, double *rcp_out_dev, void *stream)
{
    if (n <= 0 || !x_dev || !rcp_out_dev) return DRE_ERR_INVALID;
    mpc_math_selftest_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, x_dev, rcp_out_dev);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}

int dre_mpc_reset_launch(const MpcDeviceState &st, const uint8_t *mask, cudaStream_t s)
{
    mpc_reset_kernel<<<(st.N + 255) / 256, 256, 0, s>>>(st.iteration0, mask, st.N);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}
