// capi.cu -- extern "C" entry points of libdre.so for the ORCA pass and the MPC (include/dre.h).
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "launch.cuh"

static thread_local char g_cuda_err[512] = "";

void dre_set_cuda_error(cudaError_t e, const char *file, int line)
{
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

void dre_set_error_text(const char *msg) { snprintf(g_cuda_err, sizeof g_cuda_err, "%s", msg); }

extern "C" const char *dre_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char *dre_strerror(int code)
{
    switch (code) {
    case DRE_OK: return "ok";
    case DRE_ERR_INVALID: return "invalid argument or unsupported size";
    case DRE_ERR_CUDA: return "CUDA runtime error (see dre_last_cuda_error)";
    case DRE_ERR_NOMEM: return "out of memory";
    case DRE_ERR_STATE: return "call order violated";
    default: return "unknown error";
    }
}

extern "C" int dre_version(void) { return 100; }

// ------------------------------------------------------------------------------------------------
// CUDA-core peak microbenchmark (the roofline denominators SURVEY 8(d) asks the builder to measure)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T *out, int iters)
{
    T a0 = (T)threadIdx.x * (T)1e-6, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3, a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999999, c = (T)1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
            a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == (T)-1) out[0] = a0;
}

extern "C" int dre_measure_peak(int32_t kind, double *tflops)
{
    if (!tflops || (kind != 0 && kind != 1)) return DRE_ERR_INVALID;
    int dev = 0, sms = 0;
    DRE_CUDA_TRY(cudaGetDevice(&dev));
    DRE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void *buf = nullptr;
    DRE_CUDA_TRY(cudaMalloc(&buf, 64));
    cudaEvent_t e0, e1;
    DRE_CUDA_TRY(cudaEventCreate(&e0));
    DRE_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = kind == 0 ? 4096 : 2048;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        DRE_CUDA_TRY(cudaEventRecord(e0, 0));
        if (kind == 0) fma_chain_kernel<float><<<blocks, threads>>>((float *)buf, iters);
        else fma_chain_kernel<double><<<blocks, threads>>>((double *)buf, iters);
        DRE_CUDA_TRY(cudaEventRecord(e1, 0));
        DRE_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        DRE_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / ((double)best * 1e-3) * 1e-12;
    return DRE_OK;
}

// ------------------------------------------------------------------------------------------------
// ORCA
// ------------------------------------------------------------------------------------------------
extern "C" int dre_orca_predict(const dre_orca_params *p, int32_t E, int32_t Nh, const double *hpos, const float *hvel,
                                const double *goal, const double *rpos, const uint8_t *is_static, float *vnew,
                                dre_orca_stats *stats, void *stream)
{
    if (!p || !hpos || !hvel || !goal || !vnew || (p->robot_visible && !rpos)) return DRE_ERR_INVALID;
    return dre_orca_launch(*p, E, Nh, hpos, hvel, goal, rpos, is_static, vnew, stats, (cudaStream_t)stream);
}

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? DRE_OK : DRE_ERR_NOMEM; }
};
}  // namespace

extern "C" int dre_orca_predict_host(const dre_orca_params *p, int32_t E, int32_t Nh, const double *hpos, const float *hvel,
                                     const double *goal, const double *rpos, const uint8_t *is_static, float *vnew,
                                     dre_orca_stats *stats)
{
    if (!p || E <= 0 || Nh <= 0) return DRE_ERR_INVALID;
    const size_t n = (size_t)E * Nh;
    DevBuf dpos, dvel, dgoal, drp, dst, dv, dstats;
    if (dpos.alloc(n * 16) || dvel.alloc(n * 8) || dgoal.alloc(n * 16) || drp.alloc((size_t)E * 16) || dv.alloc(n * 8)) return DRE_ERR_NOMEM;
    if (is_static && dst.alloc(n)) return DRE_ERR_NOMEM;
    if (stats && dstats.alloc(n * sizeof(dre_orca_stats))) return DRE_ERR_NOMEM;
    cudaStream_t s = 0;
    DRE_CUDA_TRY(cudaMemcpyAsync(dpos.p, hpos, n * 16, cudaMemcpyHostToDevice, s));
    DRE_CUDA_TRY(cudaMemcpyAsync(dvel.p, hvel, n * 8, cudaMemcpyHostToDevice, s));
    DRE_CUDA_TRY(cudaMemcpyAsync(dgoal.p, goal, n * 16, cudaMemcpyHostToDevice, s));
    if (rpos) DRE_CUDA_TRY(cudaMemcpyAsync(drp.p, rpos, (size_t)E * 16, cudaMemcpyHostToDevice, s));
    if (is_static) DRE_CUDA_TRY(cudaMemcpyAsync(dst.p, is_static, n, cudaMemcpyHostToDevice, s));
    const int rc = dre_orca_launch(*p, E, Nh, (const double *)dpos.p, (const float *)dvel.p, (const double *)dgoal.p,
                                   (const double *)drp.p, is_static ? (const uint8_t *)dst.p : nullptr, (float *)dv.p,
                                   stats ? (dre_orca_stats *)dstats.p : nullptr, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaMemcpyAsync(vnew, dv.p, n * 8, cudaMemcpyDeviceToHost, s));
    if (stats) DRE_CUDA_TRY(cudaMemcpyAsync(stats, dstats.p, n * sizeof(dre_orca_stats), cudaMemcpyDeviceToHost, s));
    DRE_CUDA_TRY(cudaStreamSynchronize(s));
    return DRE_OK;
}

extern "C" int dre_orca_solve_host(const dre_orca_params *p, int32_t N, int32_t n_others, const float *self8, const float *others5,
                                   float *vnew, dre_orca_stats *stats)
{
    if (!p || N <= 0 || n_others < 0 || !self8 || (n_others > 0 && !others5) || !vnew) return DRE_ERR_INVALID;
    DevBuf dself, doth, dv, dstats, dq;
    const size_t no = (size_t)N * (n_others > 0 ? n_others : 1) * 5 * sizeof(float);
    if (dself.alloc((size_t)N * 8 * sizeof(float)) || doth.alloc(no) || dv.alloc((size_t)N * 2 * sizeof(float)) || dq.alloc(sizeof(int))) return DRE_ERR_NOMEM;
    if (stats && dstats.alloc((size_t)N * sizeof(dre_orca_stats))) return DRE_ERR_NOMEM;
    cudaStream_t s = 0;
    DRE_CUDA_TRY(cudaMemcpyAsync(dself.p, self8, (size_t)N * 8 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (n_others > 0) DRE_CUDA_TRY(cudaMemcpyAsync(doth.p, others5, no, cudaMemcpyHostToDevice, s));
    const int rc = dre_orca_joint_launch(*p, N, n_others, (const float *)dself.p, (const float *)doth.p, (float *)dv.p,
                                         stats ? (dre_orca_stats *)dstats.p : nullptr, (int *)dq.p, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaMemcpyAsync(vnew, dv.p, (size_t)N * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (stats) DRE_CUDA_TRY(cudaMemcpyAsync(stats, dstats.p, (size_t)N * sizeof(dre_orca_stats), cudaMemcpyDeviceToHost, s));
    DRE_CUDA_TRY(cudaStreamSynchronize(s));
    return DRE_OK;
}

// ------------------------------------------------------------------------------------------------
// MPC
// ------------------------------------------------------------------------------------------------
struct dre_mpc {
    dre_mpc_params prm;
    MpcDeviceState st;
    double *d_paths = nullptr;
    int32_t *d_lens = nullptr;
    // staging for the *_host entry point
    int32_t *d_pid = nullptr;
    double *d_interp = nullptr, *d_pose = nullptr, *d_ctrl = nullptr;
    dre_mpc_status *d_status = nullptr;
};

extern "C" int dre_mpc_create(const dre_mpc_params *p, int32_t N, dre_mpc **out)
{
    if (!p || !out || N <= 0 || p->horizon < 1 || p->horizon > DRE_MAX_HORIZON) return DRE_ERR_INVALID;
    if (!(p->time_step > 0) || !(p->v_max > 0) || !(p->w_max > 0) || !(p->node_separation > 0)) return DRE_ERR_INVALID;
    dre_mpc *h = new (std::nothrow) dre_mpc();
    if (!h) return DRE_ERR_NOMEM;
    h->prm = *p;
    if (h->prm.max_retries <= 0) h->prm.max_retries = 20;
    if (h->prm.max_iterations <= 0) h->prm.max_iterations = 1500;
    memset(&h->st, 0, sizeof h->st);
    h->st.N = N; h->st.K = p->horizon;
    const size_t K = (size_t)p->horizon;
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&h->st.warm_T, K * N * 4 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&h->st.warm_u, K * N * 2 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&h->st.iteration0, (size_t)N)) != cudaSuccess ||
        (e = cudaMalloc(&h->st.queue, 128)) != cudaSuccess || (e = cudaMemset(h->st.queue, 0, 128)) != cudaSuccess ||

        (e = cudaMemset(h->st.warm_T, 0, K * N * 4 * sizeof(double))) != cudaSuccess ||
        (e = cudaMemset(h->st.warm_u, 0, K * N * 2 * sizeof(double))) != cudaSuccess ||
        (e = cudaMemset(h->st.iteration0, 1, (size_t)N)) != cudaSuccess) {
        dre_set_cuda_error(e, __FILE__, __LINE__);
        dre_mpc_destroy(h);
        return DRE_ERR_CUDA;
    }
    *out = h;
    return DRE_OK;
}

extern "C" int dre_mpc_destroy(dre_mpc *h)
{
    if (!h) return DRE_OK;
    cudaFree(h->st.warm_T); cudaFree(h->st.warm_u); cudaFree(h->st.iteration0); cudaFree(h->st.queue);
    cudaFree(h->d_paths); cudaFree(h->d_lens);
    cudaFree(h->d_pid); cudaFree(h->d_interp); cudaFree(h->d_pose); cudaFree(h->d_ctrl); cudaFree(h->d_status);
    delete h;
    return DRE_OK;
}

extern "C" int dre_mpc_set_paths(dre_mpc *h, const double *paths_host, int32_t num_paths, int32_t path_stride,
                                 const int32_t *lens_host)
{
    if (!h || !paths_host || !lens_host || num_paths <= 0 || path_stride <= 1) return DRE_ERR_INVALID;
    for (int i = 0; i < num_paths; ++i)
        if (lens_host[i] < 2 || lens_host[i] > path_stride) return DRE_ERR_INVALID;
    cudaFree(h->d_paths); cudaFree(h->d_lens);
    h->d_paths = nullptr; h->d_lens = nullptr;
    const size_t nb = (size_t)num_paths * 3 * path_stride * sizeof(double);
    DRE_CUDA_TRY(cudaMalloc(&h->d_paths, nb));
    DRE_CUDA_TRY(cudaMalloc(&h->d_lens, (size_t)num_paths * sizeof(int32_t)));
    DRE_CUDA_TRY(cudaMemcpy(h->d_paths, paths_host, nb, cudaMemcpyHostToDevice));
    DRE_CUDA_TRY(cudaMemcpy(h->d_lens, lens_host, (size_t)num_paths * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->st.paths = h->d_paths; h->st.lens = h->d_lens; h->st.num_paths = num_paths; h->st.path_stride = path_stride;
    return DRE_OK;
}

extern "C" int dre_mpc_reset(dre_mpc *h, const uint8_t *mask, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    return dre_mpc_reset_launch(h->st, mask, (cudaStream_t)stream);
}

extern "C" int dre_mpc_act(dre_mpc *h, const int32_t *path_id, const double *interp_index, const double *pose,
                           double *controls, dre_mpc_status *status, void *stream)
{
    if (!h || !interp_index || !pose || !controls) return DRE_ERR_INVALID;
    if (!h->st.paths) return DRE_ERR_STATE;
    return dre_mpc_launch(h->prm, h->st, path_id, interp_index, pose, 3, controls, 2 * h->prm.horizon, 2 * h->prm.horizon,
                          status, nullptr, (cudaStream_t)stream);
}

extern "C" int dre_mpc_act_host(dre_mpc *h, const int32_t *path_id, const double *interp_index, const double *pose,
                                double *controls, dre_mpc_status *status)
{
    if (!h || !interp_index || !pose || !controls) return DRE_ERR_INVALID;
    if (!h->st.paths) return DRE_ERR_STATE;
    const size_t N = (size_t)h->st.N, K2 = 2 * (size_t)h->prm.horizon;
    if (!h->d_interp) {
        DRE_CUDA_TRY(cudaMalloc(&h->d_pid, N * sizeof(int32_t)));
        DRE_CUDA_TRY(cudaMalloc(&h->d_interp, N * sizeof(double)));
        DRE_CUDA_TRY(cudaMalloc(&h->d_pose, N * 3 * sizeof(double)));
        DRE_CUDA_TRY(cudaMalloc(&h->d_ctrl, N * K2 * sizeof(double)));
        DRE_CUDA_TRY(cudaMalloc(&h->d_status, N * sizeof(dre_mpc_status)));
    }
    cudaStream_t s = 0;
    if (path_id) DRE_CUDA_TRY(cudaMemcpyAsync(h->d_pid, path_id, N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    DRE_CUDA_TRY(cudaMemcpyAsync(h->d_interp, interp_index, N * sizeof(double), cudaMemcpyHostToDevice, s));
    DRE_CUDA_TRY(cudaMemcpyAsync(h->d_pose, pose, N * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    const int rc = dre_mpc_launch(h->prm, h->st, path_id ? h->d_pid : nullptr, h->d_interp, h->d_pose, 3, h->d_ctrl, (int)K2,
                                  (int)K2, status ? h->d_status : nullptr, nullptr, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaMemcpyAsync(controls, h->d_ctrl, N * K2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (status) DRE_CUDA_TRY(cudaMemcpyAsync(status, h->d_status, N * sizeof(dre_mpc_status), cudaMemcpyDeviceToHost, s));
    DRE_CUDA_TRY(cudaStreamSynchronize(s));
    return DRE_OK;
}

extern "C" int dre_mpc_get_warm(dre_mpc *h, double *warm_T, double *warm_u, uint8_t *iteration0)
{
    if (!h) return DRE_ERR_INVALID;
    const size_t N = (size_t)h->st.N, K = (size_t)h->st.K;
    DRE_CUDA_TRY(cudaDeviceSynchronize());
    if (warm_T) DRE_CUDA_TRY(cudaMemcpy(warm_T, h->st.warm_T, N * K * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (warm_u) DRE_CUDA_TRY(cudaMemcpy(warm_u, h->st.warm_u, N * K * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (iteration0) DRE_CUDA_TRY(cudaMemcpy(iteration0, h->st.iteration0, N, cudaMemcpyDeviceToHost));
    return DRE_OK;
}

extern "C" int dre_mpc_set_warm(dre_mpc *h, const double *warm_T, const double *warm_u, const uint8_t *iteration0)
{
    if (!h) return DRE_ERR_INVALID;
    const size_t N = (size_t)h->st.N, K = (size_t)h->st.K;
    DRE_CUDA_TRY(cudaDeviceSynchronize());
    if (warm_T) DRE_CUDA_TRY(cudaMemcpy(h->st.warm_T, warm_T, N * K * 4 * sizeof(double), cudaMemcpyHostToDevice));
    if (warm_u) DRE_CUDA_TRY(cudaMemcpy(h->st.warm_u, warm_u, N * K * 2 * sizeof(double), cudaMemcpyHostToDevice));
    if (iteration0) DRE_CUDA_TRY(cudaMemcpy(h->st.iteration0, iteration0, N, cudaMemcpyHostToDevice));
    return DRE_OK;
}

// accessors used by env_capi.cu
const dre_mpc_params &dre_mpc_params_of(const dre_mpc *h) { return h->prm; }
const MpcDeviceState &dre_mpc_state_of(const dre_mpc *h) { return h->st; }
