// tests/host_harness/simt_emu.h -- TEST INFRASTRUCTURE. A small SIMT emulator that lets the CPU-only test suite EXECUTE the
// product's CUDA kernel sources: the threads of one CTA are coroutines (ucontext) scheduled round-robin on one OS thread, the
// CTAs of a grid run one after the other, and every collective the kernels use -- __shfl_*_sync, __ballot_sync, __any_sync,
// __syncwarp, __syncthreads(_or), cooperative-groups tiles, named barriers -- is an exchange through per-thread slots between
// two barriers, a barrier being "yield until every live thread named in the mask has arrived". Divergent lane-groups of a warp
// therefore follow their own paths and re-converge exactly where the source says. Atomics and fences are plain operations (one
// OS thread). Nothing here is shipped and nothing here is a fallback: the product path is the CUDA build.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>

// (C++17 inline variables: ONE emulator state for all translation units of the emulated library)
namespace emu {
constexpr int MAXT = 1024;
constexpr size_t STACK_BYTES = 512 * 1024;
struct Lane {
    ucontext_t ctx;
    void *sp = nullptr;         // x86-64 fast path: the suspended coroutine's stack pointer
    char *stack = nullptr;
    bool done = false;
};
struct Bar {
    int arrived = 0;
    unsigned gen = 0;
};
inline Lane lanes[MAXT];
inline ucontext_t main_ctx;
inline int cur = 0, NT = 0, alive = 0;
inline uint64_t slot[MAXT];
inline unsigned exited[MAXT / 32];          // per warp: lanes that have returned from the kernel
inline Bar group_bar[MAXT], named_bar[16], cta_bar;
inline int or_acc[3] = {0, 0, 0};
inline uint3 tid3_v, bid3_v;
inline dim3 gdim_v(1, 1, 1), bdim_v(1, 1, 1);
inline unsigned char *dyn_smem = nullptr;
inline const std::function<void()> *body = nullptr;

inline void *main_sp = nullptr;
#if defined(__x86_64__) && !defined(DRE_EMU_UCONTEXT)
// A coroutine switch is a dozen instructions: push the callee-saved registers, swap the stack pointer, pop, return. swapcontext
// does the same plus two sigprocmask system calls per switch, which was a third of the suite's run time.
#define DRE_EMU_FAST_SWITCH 1
extern "C" void dre_emu_switch(void **save_sp, void *load_sp);
__asm__(".text\n"
        ".weak dre_emu_switch\n"
        ".type dre_emu_switch,@function\n"
        "dre_emu_switch:\n"
        "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
        "    movq %rsp, (%rdi)\n"
        "    movq %rsi, %rsp\n"
        "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
        "    ret\n"
        ".size dre_emu_switch, .-dre_emu_switch\n");
inline void yield() { dre_emu_switch(&lanes[cur].sp, main_sp); }
inline void resume(int t) { dre_emu_switch(&main_sp, lanes[t].sp); }
inline void prepare(Lane &L, void (*entry)())
{
    // [r15 r14 r13 r12 rbx rbp | entry | 0]: after the six pops and the ret, rsp = 8 mod 16 as at any function entry
    void **top = reinterpret_cast<void **>((reinterpret_cast<uintptr_t>(L.stack) + STACK_BYTES) & ~(uintptr_t)15);
    void **sp = top - 8;
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;
    sp[6] = reinterpret_cast<void *>(entry);
    sp[7] = nullptr;
    L.sp = sp;
}
#else
inline void yield() { swapcontext(&lanes[cur].ctx, &main_ctx); }
inline void resume(int t) { swapcontext(&main_ctx, &lanes[t].ctx); }
inline void prepare(Lane &L, void (*entry)())
{
    getcontext(&L.ctx);
    L.ctx.uc_stack.ss_sp = L.stack;
    L.ctx.uc_stack.ss_size = STACK_BYTES;
    L.ctx.uc_link = &main_ctx;
    makecontext(&L.ctx, entry, 0);
}
#endif
// every waiter re-evaluates the release condition: threads that exit while others wait are never waited for
template <class Expected> inline void barrier(Bar &b, Expected expected)
{
    const unsigned gen = b.gen;
    ++b.arrived;
    for (;;) {
        if (b.gen != gen) return;
        if (b.arrived >= expected()) { b.arrived = 0; ++b.gen; return; }
        yield();
    }
}
inline int lane_id() { return cur & 31; }
inline int warp_base() { return cur & ~31; }
inline unsigned warp_members()              // lanes of this thread's warp that exist in the CTA
{
    const int n = NT - warp_base();
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}
inline Bar &bar_of(unsigned mask) { return group_bar[warp_base() + __builtin_ctz(mask)]; }
inline void warp_barrier(unsigned mask)
{
    mask &= warp_members();
    const int w = cur >> 5;
    barrier(bar_of(mask), [mask, w] { return __builtin_popcount(mask & ~exited[w]); });
}
inline uint64_t exchange(unsigned mask, uint64_t v, int src_lane)
{
    slot[cur] = v;
    warp_barrier(mask);
    const uint64_t o = slot[warp_base() + src_lane];
    warp_barrier(mask);
    return o;
}
inline void cta_barrier() { barrier(cta_bar, [] { return alive; }); }
template <class T> inline uint64_t bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
inline const uint3 &tid3() { tid3_v.x = (unsigned)cur; tid3_v.y = tid3_v.z = 0; return tid3_v; }

inline void lane_entry()
{
    (*body)();
    lanes[cur].done = true;
    exited[cur >> 5] |= 1u << (cur & 31);
    --alive;
    for (;;) yield();
}

// kernel<<<grid, block, smem, stream>>>(args...) becomes  emu::Launcher(grid, block, smem, stream)([&] { kernel(args...); });
struct Launcher {
    dim3 grid, block;
    size_t smem;
    Launcher(dim3 g, dim3 b, size_t s = 0, const void * = nullptr) : grid(g), block(b), smem(s) {}
    void operator()(const std::function<void()> &f) const
    {
        const int nt = (int)(block.x * block.y * block.z);
        static const bool trace = getenv("DRE_EMU_TRACE") != nullptr;
        if (trace) {
            fprintf(stderr, "[emu] launch grid %u block %d smem %zu\n", grid.x, nt, smem);
            static bool handler = false;
            if (!handler) {
                handler = true;
                static char alt[64 * 1024];
                stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof alt; ss.ss_flags = 0;
                sigaltstack(&ss, nullptr);
                struct sigaction sa; memset(&sa, 0, sizeof sa);
                sa.sa_flags = SA_ONSTACK | SA_SIGINFO;
                sa.sa_sigaction = [](int, siginfo_t *si, void *) {
                    fprintf(stderr, "[emu] SIGSEGV at address %p, emulated thread %d of block %u\n", si->si_addr, cur, bid3_v.x);
                    void *bt[48]; const int n = backtrace(bt, 48); backtrace_symbols_fd(bt, n, 2); _exit(139);
                };
                sigaction(SIGSEGV, &sa, nullptr);
            }
        }
        if (nt < 1 || nt > MAXT) abort();
        void *sm = nullptr;                     // an exact-size block: an overrun of the launch's shared memory is a heap overrun (ASan build)
        if (posix_memalign(&sm, 64, smem ? smem : 1) != 0) abort();
        memset(sm, 0, smem ? smem : 1);
        dyn_smem = static_cast<unsigned char *>(sm);
        body = &f;
        NT = nt;
        bdim_v = block;
        gdim_v = grid;
        for (unsigned b = 0; b < grid.x; ++b) {
            bid3_v.x = b; bid3_v.y = bid3_v.z = 0;
            cta_bar = Bar();
            or_acc[0] = or_acc[1] = or_acc[2] = 0;
            for (auto &x : group_bar) x = Bar();
            for (auto &x : named_bar) x = Bar();
            memset(exited, 0, sizeof exited);
            alive = nt;
            for (int t = 0; t < nt; ++t) {
                Lane &L = lanes[t];
                if (!L.stack) L.stack = (char *)malloc(STACK_BYTES);
                L.done = false;
                prepare(L, lane_entry);
            }
            for (bool any = true; any;) {
                any = false;
                for (int t = 0; t < nt; ++t) {
                    if (lanes[t].done) continue;
                    any = true;
                    cur = t;
                    resume(t);
                }
            }
        }
        body = nullptr;
        dyn_smem = nullptr;
        free(sm);
    }
};
}  // namespace emu

// ---- warp collectives -----------------------------------------------------------------------------------------------------
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
    emu::slot[emu::cur] = p ? 1u : 0u;
    emu::warp_barrier(mask);
    unsigned out = 0;
    const unsigned live = mask & emu::warp_members() & ~emu::exited[emu::cur >> 5];
    for (int l = 0; l < 32; ++l)
        if ((live >> l) & 1u) out |= (unsigned)emu::slot[emu::warp_base() + l] << l;
    emu::warp_barrier(mask);
    return out;
}
static inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
static inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, !p) == 0u; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
static inline void __syncthreads() { emu::cta_barrier(); }
static inline int __syncthreads_or(int x)
{
    // three rotating accumulators: arrivals at barrier g clear the one of barrier g + 1 (nobody can be there before g releases),
    // while threads that have not yet resumed from barrier g - 1 still read theirs
    const int idx = (int)(emu::cta_bar.gen % 3u);
    emu::or_acc[idx] |= x;
    emu::or_acc[(idx + 1) % 3] = 0;
    emu::cta_barrier();
    return emu::or_acc[idx];
}
static inline void dre_emul_bar_sync(int id, int count) { emu::barrier(emu::named_bar[id & 15], [count] { return count; }); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

// ---- atomics, loads, bit helpers (one OS thread: plain operations) -------------------------------------------------------------
template <class T, class U> static inline T atomicAdd(T *p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { const T o = *p; *p = (T)v; return o; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __activemask() { return emu::warp_members() & ~emu::exited[emu::cur >> 5]; }
static inline unsigned __float_as_uint(float f) { return (unsigned)emu::bits(f); }
static inline float __uint_as_float(unsigned u) { return emu::unbits<float>(u); }
static inline int __float_as_int(float f) { return (int)emu::bits(f); }
static inline float __int_as_float(int u) { return emu::unbits<float>((uint64_t)(uint32_t)u); }
static inline int __double2hiint(double x) { return (int)(emu::bits(x) >> 32); }
static inline int __double2loint(double x) { return (int)(emu::bits(x) & 0xffffffffull); }
static inline double __hiloint2double(int hi, int lo) { return emu::unbits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
// rcp.approx.ftz.f64: a reciprocal seed whose low 32 bits are zero (PTX ISA); the Newton steps behind it square the error twice
static inline double dre_emul_rcp_approx(double x)
{
    uint64_t u = emu::bits(1.0 / x);
    u &= 0xffffffff00000000ull;
    return emu::unbits<double>(u);
}

// ---- cooperative groups: the subset the kernels use ---------------------------------------------------------------------------
namespace cooperative_groups {
struct thread_block {
    void sync() const { __syncthreads(); }
    unsigned thread_rank() const { return (unsigned)emu::cur; }
};
inline thread_block this_thread_block() { return thread_block(); }
template <int G> struct thread_block_tile {
    unsigned mask() const { return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (emu::lane_id() & ~(G - 1))); }
    unsigned thread_rank() const { return (unsigned)(emu::lane_id() & (G - 1)); }
    unsigned meta_group_rank() const { return (unsigned)(emu::cur / G); }
    unsigned size() const { return G; }
    unsigned ballot(bool p) const { return (__ballot_sync(mask(), p) >> (emu::lane_id() & ~(G - 1))) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u)); }
    bool any(bool p) const { return ballot(p) != 0u; }
    bool all(bool p) const { return ballot(!p) == 0u; }
    template <class T> T shfl(T v, int src) const { return __shfl_sync(mask(), v, src, G); }
    template <class T> T shfl_xor(T v, int m) const { return __shfl_xor_sync(mask(), v, m, G); }
    template <class T> T shfl_up(T v, int d) const { return __shfl_up_sync(mask(), v, (unsigned)d, G); }
    template <class T> T shfl_down(T v, int d) const { return __shfl_down_sync(mask(), v, (unsigned)d, G); }
    void sync() const { __syncwarp(mask()); }
};
template <int G> inline thread_block_tile<G> tiled_partition(const thread_block &) { return thread_block_tile<G>(); }
}  // namespace cooperative_groups

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))   // crt/host_defines.h leaves it to nvcc under GCC
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif
#define threadIdx (emu::tid3())
#define blockIdx (emu::bid3_v)
#define gridDim (emu::gdim_v)
#define blockDim (emu::bdim_v)
// __shared__ variables of a kernel: one CTA runs at a time, so a function-level static is the CTA's shared memory
#undef __shared__
#define __shared__ static
