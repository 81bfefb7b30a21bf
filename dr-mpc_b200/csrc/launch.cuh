// launch.cuh -- host-side helpers shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "dre_common.cuh"

void dre_set_cuda_error(cudaError_t e, const char *file, int line);
void dre_set_error_text(const char *msg);   // what dre_last_cuda_error() reports for a non-CUDA failure behind DRE_ERR_CUDA

#define DRE_CUDA_TRY(call)                                      \
    do {                                                        \
        cudaError_t _e = (call);                                \
        if (_e != cudaSuccess) {                                \
            dre_set_cuda_error(_e, __FILE__, __LINE__);         \
            return DRE_ERR_CUDA;                                \
        }                                                       \
    } while (0)

// ---- orca_kernels.cu ----
int dre_orca_pick_group(const dre_orca_params &p, int Nh);
bool dre_orca_use_cutoff(const dre_orca_params &p);
int dre_orca_launch(const dre_orca_params &p, int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                    const double *rpos, const uint8_t *is_static, float *vnew, dre_orca_stats *stats, cudaStream_t s);

int dre_orca_joint_launch(const dre_orca_params &p, int N, int n_others, const float *self8, const float *others5, float *vnew,
                          dre_orca_stats *stats, int *queue, cudaStream_t s);

// ---- mpc_kernels.cu ----
struct MpcDeviceState {
    int N, K;
    double *warm_T;      // [N][K][4]  (c,s,x,y) of the inverse-pose guesses T_1..T_K
    double *warm_u;      // [N][K][2]
    uint8_t *iteration0; // [N]
    int *queue;          // work counter of the persistent MPC kernel
    const double *paths; // [P][3][stride]
    const int32_t *lens; // [P]
    int num_paths, path_stride;
};
int dre_mpc_launch(const dre_mpc_params &p, const MpcDeviceState &st, const int32_t *path_id, const double *interp_index,
                   const double *pose, int pose_stride, double *controls, int controls_stride, int controls_count,
                   dre_mpc_status *status, const uint8_t *active_mask, cudaStream_t s);
int dre_mpc_reset_launch(const MpcDeviceState &st, const uint8_t *mask, cudaStream_t s);
