// tests/host_harness/mpc_host.cpp -- TEST INFRASTRUCTURE. Compiles the PRODUCT's device header
// (dr-mpc_b200/csrc/mpc_device.cuh) for the host so the CPU-only test suite can check the kernel's
// per-problem logic against the oracle without a GPU. Never shipped, never a fallback.
#include <cuda_runtime.h>
#include <cstring>
#include "../../dr-mpc_b200/csrc/mpc_device.cuh"

template <int KC>
static void run(const dre_mpc_params *prm, int N, const double *paths, int stride, const int32_t *lens,
                const int32_t *path_id, const double *interp, const double *pose, double *warm_T, double *warm_u,
                uint8_t *it0, double *controls, dre_mpc_status *status)
{
    static MpcSystem<KC> S;
    const int K = prm->horizon;
    for (int n = 0; n < N; ++n) {
        const int pid = path_id ? path_id[n] : 0;
        mpc_act_one<KC>(*prm, paths + (size_t)pid * 3 * stride, stride, lens[pid], interp[n], pose[3 * n], pose[3 * n + 1],
                        pose[3 * n + 2], reinterpret_cast<double4 *>(warm_T) + (size_t)n * K, reinterpret_cast<double2 *>(warm_u) + (size_t)n * K,
                        (size_t)1, it0 + n, controls + (size_t)n * 2 * K, 2 * K, status ? status + n : nullptr, S);
    }
}

// warm_T: [N][K][4], warm_u: [N][K][2] (the device layout)
extern "C" int host_mpc_act(const dre_mpc_params *prm, int N, const double *paths, int stride, const int32_t *lens,
                            const int32_t *path_id, const double *interp, const double *pose, double *warm_T,
                            double *warm_u, uint8_t *it0, double *controls, dre_mpc_status *status)
{
    const int K = prm->horizon;
    if (K == 4) run<4>(prm, N, paths, stride, lens, path_id, interp, pose, warm_T, warm_u, it0, controls, status);
    else if (K <= 10) run<10>(prm, N, paths, stride, lens, path_id, interp, pose, warm_T, warm_u, it0, controls, status);
    else if (K <= 40) run<40>(prm, N, paths, stride, lens, path_id, interp, pose, warm_T, warm_u, it0, controls, status);
    else return -1;
    return 0;
}
