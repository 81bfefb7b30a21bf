// mpc_kernels.cu -- batched MPC.act: one thread per problem runs the whole LM solve (mpc_device.cuh).
//
// Data layout: warm-start poses/controls are horizon-major ([K][N]) so that the 32 problems of a warp
// read/write consecutive 32 B / 16 B elements (coalesced double4 / double2 accesses); per-problem
// inputs (path id, interpolation index, pose) and outputs are problem-major.
#include "mpc_device.cuh"
#include "launch.cuh"

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
                    reinterpret_cast<double4 *>(st.warm_T) + n, reinterpret_cast<double2 *>(st.warm_u) + n, (size_t)st.N,
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

int dre_mpc_launch(const dre_mpc_params &p, const MpcDeviceState &st, const int32_t *path_id, const double *interp_index,
                   const double *pose, int pose_stride, double *controls, int controls_stride, int controls_count,
                   dre_mpc_status *status, const uint8_t *active_mask, cudaStream_t s)
{
    const int K = p.horizon;
    if (K < 1 || K > DRE_MAX_HORIZON || st.N <= 0 || st.paths == nullptr) return DRE_ERR_INVALID;
#define DRE_MPC_CASE(KC) return launch_mpc<KC>(p, st, path_id, interp_index, pose, pose_stride, controls, controls_stride, controls_count, status, active_mask, s)
    if (K == 4) DRE_MPC_CASE(4);
    if (K <= 10) DRE_MPC_CASE(10);
    if (K <= 20) DRE_MPC_CASE(20);
    if (K <= 30) DRE_MPC_CASE(30);
    DRE_MPC_CASE(40);
#undef DRE_MPC_CASE
}

int dre_mpc_reset_launch(const MpcDeviceState &st, const uint8_t *mask, cudaStream_t s)
{
    mpc_reset_kernel<<<(st.N + 255) / 256, 256, 0, s>>>(st.iteration0, mask, st.N);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}
