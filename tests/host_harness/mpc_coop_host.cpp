// tests/host_harness/mpc_coop_host.cpp -- TEST INFRASTRUCTURE. Runs the PRODUCT's cooperative MPC kernel source
// (dr-mpc_b200/csrc/mpc_coop.cuh: mpc_coop_persistent -- the lane-per-horizon-step LM state machine the B200 executes) on the
// host: one CTA of DRE_MPC_CTA threads is emulated by that many coroutines (ucontext) scheduled round-robin on ONE OS thread;
// every warp / CTA collective of the source (__shfl_*_sync, __ballot_sync, __syncthreads_or, the named barrier of the two-warp
// groups) is an exchange through per-lane slots between two barriers, a barrier being "yield until everybody named in the mask
// has arrived". The lane-groups of a warp therefore run their own LM paths and re-converge exactly as the source prescribes.
//
// The header is compiled from a generated copy (tests/test_host_mpc_coop.py writes it next to this file) in which exactly three
// lines differ from the product source: the `#if defined(__CUDACC__)` guard (-> 1) and the two inline-PTX statements
//   asm("rcp.approx.ftz.f64 %0, %1;" ...)     -> r = dre_emul_rcp_approx(x);      (a reciprocal seed good to 2^-20, like the hardware's 2^-23)
//   asm volatile("bar.sync %0, 64;" ...)      -> dre_emul_bar_sync(bar_id, 64);
// Never shipped, never a fallback: the product path is the CUDA kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// SIMT emulation: coroutines = threads of one CTA
// ---------------------------------------------------------------------------------------------------------------
namespace emu {
constexpr int MAXT = 128;
struct Lane {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};
struct Bar {
    int arrived = 0;
    unsigned gen = 0;
};
static Lane lanes[MAXT];
static ucontext_t main_ctx;
static int cur = 0, NT = MAXT;
static uint64_t slot[MAXT];
static Bar group_bar[MAXT], named_bar[16], cta_bar;
static int or_acc[2] = {0, 0};
static uint3 tid3_v, bid3_v;
static dim3 gdim_v(1, 1, 1), bdim_v(MAXT, 1, 1);

inline void yield() { swapcontext(&lanes[cur].ctx, &main_ctx); }
inline void barrier(Bar &b, int count)
{
    const unsigned gen = b.gen;
    if (++b.arrived == count) { b.arrived = 0; ++b.gen; }
    else while (b.gen == gen) yield();
}
inline int lane_id() { return cur & 31; }
inline Bar &bar_of(unsigned mask) { return group_bar[(cur & ~31) + __builtin_ctz(mask)]; }
inline uint64_t exchange(unsigned mask, uint64_t v, int src_lane)
{
    Bar &b = bar_of(mask);
    const int n = __builtin_popcount(mask);
    slot[cur] = v;
    barrier(b, n);
    const uint64_t o = slot[(cur & ~31) + src_lane];
    barrier(b, n);
    return o;
}
template <class T> inline uint64_t bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
inline const uint3 &tid3() { tid3_v.x = (unsigned)cur; tid3_v.y = tid3_v.z = 0; return tid3_v; }
}  // namespace emu

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = emu::lane_id();
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), (lane & ~(width - 1)) | (src & (width - 1))));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane & ~(width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane + (int)delta;
    if (s > (lane | (width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane ^ o;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
static inline unsigned __ballot_sync(unsigned mask, bool p)
{
    emu::Bar &b = emu::bar_of(mask);
    const int n = __builtin_popcount(mask);
    emu::slot[emu::cur] = p ? 1u : 0u;
    emu::barrier(b, n);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) out |= (unsigned)emu::slot[(emu::cur & ~31) + l] << l;
    emu::barrier(b, n);
    return out;
}
static inline int __syncthreads_or(int x)
{
    const unsigned gen = emu::cta_bar.gen;
    const int idx = (int)(gen & 1u);
    emu::or_acc[idx] |= x;
    if (++emu::cta_bar.arrived == emu::NT) { emu::cta_bar.arrived = 0; emu::or_acc[idx ^ 1] = 0; ++emu::cta_bar.gen; }
    else while (emu::cta_bar.gen == gen) emu::yield();
    return emu::or_acc[idx];
}
static inline void dre_emul_bar_sync(int id, int count) { emu::barrier(emu::named_bar[id & 15], count); }
// rcp.approx.ftz.f64: a reciprocal seed whose low 32 bits are zero (PTX ISA); the two Newton steps behind it square the error twice
static inline double dre_emul_rcp_approx(double x)
{
    uint64_t u = emu::bits(1.0 / x);
    u &= 0xffffffff00000000ull;
    return emu::unbits<double>(u);
}
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __double2hiint(double x) { return (int)(emu::bits(x) >> 32); }
static inline int __double2loint(double x) { return (int)(emu::bits(x) & 0xffffffffull); }
static inline double __hiloint2double(int hi, int lo) { return emu::unbits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))   // crt/host_defines.h leaves it to nvcc under GCC
#endif
#define threadIdx (emu::tid3())
#define blockIdx (emu::bid3_v)
#define gridDim (emu::gdim_v)
#define blockDim (emu::bdim_v)

#include "_gen_mpc_coop.cuh"   // generated from dr-mpc_b200/csrc/mpc_coop.cuh (see the header of this file)

// ---------------------------------------------------------------------------------------------------------------
// the "kernel launch": one CTA, every thread a coroutine running mpc_coop_kernel's body
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Launch {
    const dre_mpc_params *prm;
    const CoopBatch *B;
    double *smem;
};
Launch g_launch;

template <int G> void lane_entry()
{
    LaneGroup<G> g;
    g.attach(g_launch.smem + CS_SLOTS * DRE_MPC_CTA);
    mpc_coop_persistent<G, false>(g, *g_launch.prm, *g_launch.B, g_launch.smem);
    emu::lanes[emu::cur].done = true;
    for (;;) emu::yield();
}

template <int G> void run_cta(const dre_mpc_params &prm, const CoopBatch &B)
{
    static_assert(DRE_MPC_CTA <= emu::MAXT, "CTA larger than the emulator");
    emu::NT = DRE_MPC_CTA;
    emu::bdim_v = dim3(DRE_MPC_CTA, 1, 1);
    emu::gdim_v = dim3(1, 1, 1);
    emu::bid3_v.x = emu::bid3_v.y = emu::bid3_v.z = 0;
    emu::cta_bar = emu::Bar();
    emu::or_acc[0] = emu::or_acc[1] = 0;
    for (auto &b : emu::group_bar) b = emu::Bar();
    for (auto &b : emu::named_bar) b = emu::Bar();
    std::vector<double> smem((size_t)CS_SLOTS * DRE_MPC_CTA + 64, 0.0);
    g_launch.prm = &prm; g_launch.B = &B; g_launch.smem = smem.data();
    for (int t = 0; t < emu::NT; ++t) {
        emu::Lane &L = emu::lanes[t];
        L.done = false;
        L.stack.assign(512 * 1024, 0);
        getcontext(&L.ctx);
        L.ctx.uc_stack.ss_sp = L.stack.data();
        L.ctx.uc_stack.ss_size = L.stack.size();
        L.ctx.uc_link = &emu::main_ctx;
        makecontext(&L.ctx, (void (*)())lane_entry<G>, 0);
    }
    for (bool any = true; any;) {
        any = false;
        for (int t = 0; t < emu::NT; ++t) {
            if (emu::lanes[t].done) continue;
            any = true;
            emu::cur = t;
            swapcontext(&emu::main_ctx, &emu::lanes[t].ctx);
        }
    }
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
