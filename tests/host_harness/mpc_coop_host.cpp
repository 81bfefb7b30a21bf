// tests/host_harness/mpc_coop_host.cpp -- TEST INFRASTRUCTURE. Runs the PRODUCT's cooperative MPC kernel source
// (dr-mpc_b200/csrc/mpc_coop.cuh: mpc_coop_persistent -- the lane-per-horizon-step LM state machine the B200 executes) on the
// host through the SIMT emulator of simt_emu.h: one CTA of DRE_MPC_CTA threads = that many coroutines, every warp / CTA collective
// of the source (__shfl_*_sync, __ballot_sync, __syncthreads_or, the named barrier of the two-warp groups) an exchange between
// two barriers. The lane-groups of a warp therefore run their own LM paths and re-converge exactly as the source prescribes.
//
// The header is compiled from a generated copy (tests/test_host_mpc_coop.py writes it next to this file) in which exactly three
// lines differ from the product source: the `#if defined(__CUDACC__)` guard (-> 1) and the two inline-PTX statements
//   asm("rcp.approx.ftz.f64 %0, %1;" ...)     -> r = dre_emul_rcp_approx(x);      (a reciprocal seed good to 2^-20, like the hardware's 2^-23)
//   asm volatile("bar.sync %0, 64;" ...)      -> dre_emul_bar_sync(bar_id, 64);
// Never shipped, never a fallback: the product path is the CUDA kernel.
#include "simt_emu.h"

#include "_gen_mpc_coop.cuh"   // generated from dr-mpc_b200/csrc/mpc_coop.cuh (see the header of this file)

// ---------------------------------------------------------------------------------------------------------------
// the "kernel launch": one CTA of DRE_MPC_CTA emulated threads running mpc_coop_kernel's body (mpc_kernels.cu)
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <int G> void run_cta(const dre_mpc_params &prm, const CoopBatch &B)
{
    const size_t smem = ((size_t)CS_SLOTS * DRE_MPC_CTA + (G == 64 ? 64 : 0)) * sizeof(double);
    emu::Launcher(1, DRE_MPC_CTA, smem)([&] {
        double *coop_smem = reinterpret_cast<double *>(emu::shared_base());
        LaneGroup<G> g;
        g.attach(coop_smem + CS_SLOTS * DRE_MPC_CTA);   // exchange area of the two-warp groups (K > 32)
        mpc_coop_persistent<G, false>(g, prm, B, coop_smem);
    });
}
}  // namespace

// Same contract as host_mpc_act (mpc_host.cpp): warm_T [N][K][4], warm_u [N][K][2], it0 [N] are read and updated.
extern "C" int host_mpc_coop_act(const dre_mpc_params *prm, int N, const double *paths, int stride, const int32_t *lens, int num_paths,
                                 const int32_t *path_id, const double *interp, const double *pose, double *warm_T, double *warm_u,
                                 uint8_t *it0, const uint8_t *active_mask, double *controls, dre_mpc_status *status)
{
    const int K = prm->horizon;
    if (K < 1 || K > DRE_MAX_HORIZON || N <= 0) return -1;
    int queue = 0;
    CoopBatch B;
    B.N = N; B.path_id = path_id; B.interp_index = interp; B.pose = pose; B.pose_stride = 3;
    B.controls = controls; B.controls_stride = 2 * K; B.controls_count = 2 * K; B.status = status; B.active_mask = active_mask;
    B.paths = paths; B.lens = lens; B.num_paths = num_paths; B.path_stride = stride;
    B.warm_T = reinterpret_cast<double4 *>(warm_T); B.warm_u = reinterpret_cast<double2 *>(warm_u);
    B.iteration0 = it0; B.queue = &queue; B.trace = nullptr;
    // the product's dispatch (mpc_kernels.cu: dre_mpc_launch)
    if (K <= 4) run_cta<4>(*prm, B);
    else if (K <= 8) run_cta<8>(*prm, B);
    else if (K <= 16) run_cta<16>(*prm, B);
    else if (K <= 32) run_cta<32>(*prm, B);
    else run_cta<64>(*prm, B);
    return 0;
}

extern "C" int host_mpc_coop_cta_threads() { return DRE_MPC_CTA; }
