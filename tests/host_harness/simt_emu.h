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

// Race check (tools/emu_racecheck.sh): built with -fsanitize=thread, every emulated CUDA thread is a ThreadSanitizer FIBER that is
// concurrent with all others (switches carry no synchronisation); only what CUDA's memory model orders is annotated as ordering:
// __syncthreads / __syncthreads_or / __syncwarp / tile.sync() / the named barrier (release on arrival, acquire on exit, among the
// participants), kernel boundaries (host -> kernel -> host), and atomics (real __atomic operations). Shuffles and ballots exchange
// VALUES but order no memory, exactly like shfl.sync / vote.sync. The emulator's own bookkeeping is excluded from instrumentation.
#if defined(__SANITIZE_THREAD__)
#include <sanitizer/tsan_interface.h>
#define DRE_EMU_TSAN 1
#include <sys/mman.h>
#define DRE_EMU_NOTSAN __attribute__((no_sanitize("thread"), noinline))
#else
#define DRE_EMU_TSAN 0
#define DRE_EMU_NOTSAN
#endif

// (C++17 inline variables: ONE emulator state for all translation units of the emulated library)
namespace emu {
constexpr int MAXT = 1024;
constexpr size_t STACK_BYTES = 512 * 1024;
struct Bar {
    int arrived = 0;
    unsigned gen = 0;
};
struct Lane {
    ucontext_t ctx;
    void *sp = nullptr;         // x86-64 fast path: the suspended coroutine's stack pointer
    void *fiber = nullptr;      // ThreadSanitizer build: the fiber that stands for this CUDA thread
    // what a thread blocked in a barrier last saw: the scheduler resumes it only when that can have changed
    Bar *wait = nullptr;
    unsigned wait_gen = 0;
    int wait_arrived = 0, wait_exits = 0;
    char *stack = nullptr;
    bool done = false;
};
inline Lane lanes[MAXT];
inline ucontext_t main_ctx;
inline int cur = 0, NT = 0, alive = 0, exits = 0;
inline uint64_t slot[MAXT];
inline unsigned exited[MAXT / 32];          // per warp: lanes that have returned from the kernel
inline Bar group_bar[MAXT], named_bar[16], cta_bar;
inline int or_acc[3] = {0, 0, 0};
inline uint3 tid3_v, bid3_v;
inline dim3 gdim_v(1, 1, 1), bdim_v(1, 1, 1);
inline unsigned char *dyn_smem = nullptr;
inline const std::function<void()> *body = nullptr;

inline void *main_sp = nullptr;
inline void *main_fiber = nullptr;
inline char launch_token, done_token;      // ThreadSanitizer build: kernel boundaries (host writes -> kernel, kernel writes -> host)
DRE_EMU_NOTSAN inline void tsan_to_lane(int t)
{
#if DRE_EMU_TSAN
    __tsan_switch_to_fiber(lanes[t].fiber, __tsan_switch_to_fiber_no_sync);
#else
    (void)t;
#endif
}
DRE_EMU_NOTSAN inline void tsan_to_main()
{
#if DRE_EMU_TSAN
    __tsan_switch_to_fiber(main_fiber, __tsan_switch_to_fiber_no_sync);
#endif
}
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
DRE_EMU_NOTSAN inline void yield() { tsan_to_main(); dre_emu_switch(&lanes[cur].sp, main_sp); }
DRE_EMU_NOTSAN inline void resume(int t) { tsan_to_lane(t); dre_emu_switch(&main_sp, lanes[t].sp); }
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
DRE_EMU_NOTSAN inline void yield() { tsan_to_main(); swapcontext(&lanes[cur].ctx, &main_ctx); }
DRE_EMU_NOTSAN inline void resume(int t) { tsan_to_lane(t); swapcontext(&main_ctx, &lanes[t].ctx); }
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
// `orders_memory`: a CUDA barrier (bar.sync / __syncwarp) -- the participants' earlier accesses happen before their later ones;
// false for the rendezvous inside a shuffle / vote, which moves values but orders nothing.
template <class Expected> DRE_EMU_NOTSAN inline void barrier(Bar &b, Expected expected, bool orders_memory)
{
    const unsigned gen = b.gen;
#if DRE_EMU_TSAN
    if (orders_memory) __tsan_release(&b);
#endif
    ++b.arrived;
    for (;;) {
        if (b.gen != gen) break;
        if (b.arrived >= expected()) { b.arrived = 0; ++b.gen; break; }
        Lane &me = lanes[cur];
        me.wait = &b; me.wait_gen = gen; me.wait_arrived = b.arrived; me.wait_exits = exits;
        yield();
    }
    lanes[cur].wait = nullptr;
#if DRE_EMU_TSAN
    if (orders_memory) __tsan_acquire(&b);
#endif
    (void)orders_memory;
}
DRE_EMU_NOTSAN inline int lane_id() { return cur & 31; }
DRE_EMU_NOTSAN inline int warp_base() { return cur & ~31; }
DRE_EMU_NOTSAN inline int warp_index() { return cur >> 5; }
DRE_EMU_NOTSAN inline unsigned warp_members()              // lanes of this thread's warp that exist in the CTA
{
    const int n = NT - (cur & ~31);
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}
DRE_EMU_NOTSAN inline unsigned warp_exited() { return exited[cur >> 5]; }
DRE_EMU_NOTSAN inline void warp_barrier(unsigned mask, bool orders_memory)
{
    // positive control of the race check: DRE_EMU_DROP_SYNCWARP=1 makes __syncwarp / tile.sync() wait without ordering memory, as if
    // the kernels had none -- the accesses they separate must then be reported
    static const bool drop = getenv("DRE_EMU_DROP_SYNCWARP") != nullptr;
    if (drop) orders_memory = false;
    const int n = NT - (cur & ~31);
    mask &= n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
    const int w = cur >> 5;
    barrier(group_bar[(cur & ~31) + __builtin_ctz(mask)], [mask, w] { return __builtin_popcount(mask & ~exited[w]); }, orders_memory);
}
DRE_EMU_NOTSAN inline uint64_t exchange(unsigned mask, uint64_t v, int src_lane)
{
    slot[cur] = v;
    warp_barrier(mask, false);
    const uint64_t o = slot[(cur & ~31) + src_lane];
    warp_barrier(mask, false);
    return o;
}
DRE_EMU_NOTSAN inline unsigned vote(unsigned mask, bool p)
{
    slot[cur] = p ? 1u : 0u;
    warp_barrier(mask, false);
    unsigned out = 0;
    const int n = NT - (cur & ~31);
    const unsigned live = mask & (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) & ~exited[cur >> 5];
    for (int l = 0; l < 32; ++l)
        if ((live >> l) & 1u) out |= (unsigned)slot[(cur & ~31) + l] << l;
    warp_barrier(mask, false);
    return out;
}
DRE_EMU_NOTSAN inline void cta_barrier() { barrier(cta_bar, [] { return alive; }, true); }
DRE_EMU_NOTSAN inline int cta_barrier_or(int x)
{
    // three rotating accumulators: arrivals at barrier g clear the one of barrier g + 1 (nobody can be there before g releases),
    // while threads that have not yet resumed from barrier g - 1 still read theirs
    const int idx = (int)(cta_bar.gen % 3u);
    or_acc[idx] |= x;
    or_acc[(idx + 1) % 3] = 0;
    barrier(cta_bar, [] { return alive; }, true);
    return or_acc[idx];
}
DRE_EMU_NOTSAN inline void named_barrier(int id, int count) { barrier(named_bar[id & 15], [count] { return count; }, true); }
template <class T> inline uint64_t bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
DRE_EMU_NOTSAN inline uint3 tid3() { uint3 v; v.x = (unsigned)cur; v.y = v.z = 0; return v; }
DRE_EMU_NOTSAN inline uint3 bid3() { return bid3_v; }
DRE_EMU_NOTSAN inline dim3 gdim() { return gdim_v; }
DRE_EMU_NOTSAN inline dim3 bdim() { return bdim_v; }
DRE_EMU_NOTSAN inline unsigned char *shared_base() { return dyn_smem; }

DRE_EMU_NOTSAN inline void lane_entry()
{
#if DRE_EMU_TSAN
    __tsan_acquire(&launch_token);
#endif
    (*body)();
#if DRE_EMU_TSAN
    __tsan_release(&done_token);
#endif
    lanes[cur].done = true;
    exited[cur >> 5] |= 1u << (cur & 31);
    --alive;
    ++exits;
    for (;;) yield();
}

// kernel<<<grid, block, smem, stream>>>(args...) becomes  emu::Launcher(grid, block, smem, stream)([&] { kernel(args...); });
struct Launcher {
    dim3 grid, block;
    size_t smem;
    Launcher(dim3 g, dim3 b, size_t s = 0, const void * = nullptr) : grid(g), block(b), smem(s) {}
    DRE_EMU_NOTSAN void operator()(const std::function<void()> &f) const
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
        body = &f;
        NT = nt;
        bdim_v = block;
        gdim_v = grid;
#if DRE_EMU_TSAN
        main_fiber = __tsan_get_current_fiber();
        __tsan_release(&launch_token);          // everything the host did so far happens before the kernel's threads
#endif
        std::vector<void *> held;               // race-check build: blocks are freed after the kernel boundary, not between concurrent CTAs
        // The hardware starts CTAs in no particular order: DRE_EMU_CTA_ORDER=reverse | shuffle runs them last-to-first / in a
        // pseudo-random order, so that the suite also shows that no result depends on the order (last-CTA patterns, chunk flags,
        // the persistent kernels' static first assignment).
        std::vector<unsigned> order(grid.x);
        for (unsigned b = 0; b < grid.x; ++b) order[b] = b;
        static const char *ord = getenv("DRE_EMU_CTA_ORDER");
        if (ord && ord[0] == 'r') { for (unsigned b = 0; b < grid.x / 2; ++b) { const unsigned t = order[b]; order[b] = order[grid.x - 1 - b]; order[grid.x - 1 - b] = t; } }
        if (ord && ord[0] == 's') {
            static uint64_t lcg = 0x9E3779B97F4A7C15ull;
            for (unsigned b = grid.x; b > 1; --b) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                const unsigned j = (unsigned)((lcg >> 33) % b), t = order[b - 1];
                order[b - 1] = order[j]; order[j] = t;
            }
        }
        for (unsigned bi = 0; bi < grid.x; ++bi) {
            const unsigned b = order[bi];
            // an exact-size block per CTA: an overrun of the shared memory is a heap overrun (ASan build), and the CTAs of a grid,
            // which are concurrent for the race check, do not share it
            void *sm = nullptr;
            if (posix_memalign(&sm, 64, smem ? smem : 1) != 0) abort();
            memset(sm, 0xA5, smem ? smem : 1);   // shared memory is NOT zero on the GPU: poison it, a kernel that relies on zeros must fail here
            dyn_smem = static_cast<unsigned char *>(sm);
            bid3_v.x = b; bid3_v.y = bid3_v.z = 0;
            cta_bar = Bar();
            or_acc[0] = or_acc[1] = or_acc[2] = 0;
            for (auto &x : group_bar) x = Bar();
            for (auto &x : named_bar) x = Bar();
            memset(exited, 0, sizeof exited);
            alive = nt;
            for (int t = 0; t < nt; ++t) {
                Lane &L = lanes[t];
#if DRE_EMU_TSAN
                if (!L.stack) L.stack = (char *)mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                if (L.stack == (char *)MAP_FAILED) abort();
#else
                if (!L.stack) L.stack = (char *)malloc(STACK_BYTES);
#endif
                L.done = false;
                L.wait = nullptr;
                prepare(L, lane_entry);
#if DRE_EMU_TSAN
                L.fiber = __tsan_create_fiber(0);
#endif
            }
            static const char *lord = getenv("DRE_EMU_THREAD_ORDER");      // "reverse": the scheduler visits threads last-to-first
            const bool rev = lord && lord[0] == 'r';
            for (bool any = true; any;) {
                any = false;
                for (int ti = 0; ti < nt; ++ti) {
                    const int t = rev ? nt - 1 - ti : ti;
                    Lane &L = lanes[t];
                    if (L.done) continue;
                    any = true;
                    // blocked in a barrier that nobody has reached or left since it last looked, and no thread has exited: still blocked
                    if (L.wait && L.wait->gen == L.wait_gen && L.wait->arrived == L.wait_arrived && exits == L.wait_exits) continue;
                    cur = t;
                    resume(t);
                }
            }
#if DRE_EMU_TSAN
            for (int t = 0; t < nt; ++t) {
                __tsan_destroy_fiber(lanes[t].fiber);
                lanes[t].fiber = nullptr;
                munmap(lanes[t].stack, STACK_BYTES);   // the next CTA's thread t gets a fresh mapping: ThreadSanitizer forgets the accesses
                lanes[t].stack = nullptr;              // of this CTA's thread t to the same addresses (the CTAs are concurrent for it)
            }
#endif
            dyn_smem = nullptr;
#if DRE_EMU_TSAN
            held.push_back(sm);
#else
            free(sm);
#endif
        }
#if DRE_EMU_TSAN
        __tsan_acquire(&done_token);            // ... and everything the kernel's threads did happens before what the host does next
#endif
        for (void *q : held) free(q);
        body = nullptr;
    }
};
}  // namespace emu

// ---- warp collectives -----------------------------------------------------------------------------------------------------
template <class T> DRE_EMU_NOTSAN static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = emu::lane_id();
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), (lane & ~(width - 1)) | (src & (width - 1))));
}
template <class T> DRE_EMU_NOTSAN static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane & ~(width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
template <class T> DRE_EMU_NOTSAN static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane + (int)delta;
    if (s > (lane | (width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
template <class T> DRE_EMU_NOTSAN static inline T __shfl_xor_sync(unsigned mask, T v, int o, int width = 32)
{
    const int lane = emu::lane_id();
    int s = lane ^ o;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu::unbits<T>(emu::exchange(mask, emu::bits(v), s));
}
static inline unsigned __ballot_sync(unsigned mask, bool p) { return emu::vote(mask, p); }
static inline bool __any_sync(unsigned mask, bool p) { return emu::vote(mask, p) != 0u; }
static inline bool __all_sync(unsigned mask, bool p) { return emu::vote(mask, !p) == 0u; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask, true); }
static inline void __syncthreads() { emu::cta_barrier(); }
static inline int __syncthreads_or(int x) { return emu::cta_barrier_or(x); }
static inline void dre_emul_bar_sync(int id, int count) { emu::named_barrier(id, count); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- atomics (real atomic operations: the race check must see them as such), loads, bit helpers ----------------------------------
template <class T> struct dre_emu_word { typedef T type; };
template <> struct dre_emu_word<float> { typedef uint32_t type; };
template <> struct dre_emu_word<double> { typedef uint64_t type; };
template <class T, class F> static inline T dre_emu_rmw(T *p, F f)
{
    typedef typename dre_emu_word<T>::type W;
    static_assert(sizeof(W) == sizeof(T), "atomic word");
    W *w = reinterpret_cast<W *>(p);
    W old = __atomic_load_n(w, __ATOMIC_ACQUIRE), neu;
    T o;
    do {
        memcpy(&o, &old, sizeof(T));
        const T n = f(o);
        memcpy(&neu, &n, sizeof(T));
    } while (!__atomic_compare_exchange_n(w, &old, neu, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
    return o;
}
template <class T, class U> static inline T atomicAdd(T *p, U v) { return dre_emu_rmw(p, [v](T o) { return (T)(o + (T)v); }); }
template <class T, class U> static inline T atomicMax(T *p, U v) { return dre_emu_rmw(p, [v](T o) { return (T)v > o ? (T)v : o; }); }
template <class T, class U> static inline T atomicMin(T *p, U v) { return dre_emu_rmw(p, [v](T o) { return (T)v < o ? (T)v : o; }); }
template <class T, class U> static inline T atomicOr(T *p, U v) { return dre_emu_rmw(p, [v](T o) { return (T)(o | (T)v); }); }
template <class T, class U> static inline T atomicExch(T *p, U v) { return dre_emu_rmw(p, [v](T) { return (T)v; }); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __activemask() { return emu::warp_members() & ~emu::warp_exited(); }
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
    unsigned thread_rank() const { return emu::tid3().x; }
};
inline thread_block this_thread_block() { return thread_block(); }
template <int G> struct thread_block_tile {
    unsigned mask() const { return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (emu::lane_id() & ~(G - 1))); }
    unsigned thread_rank() const { return (unsigned)(emu::lane_id() & (G - 1)); }
    unsigned meta_group_rank() const { return emu::tid3().x / G; }
    unsigned size() const { return G; }
    unsigned ballot(bool p) const { return (__ballot_sync(mask(), p) >> (emu::lane_id() & ~(G - 1))) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u)); }
    bool any(bool p) const { return ballot(p) != 0u; }
    bool all(bool p) const { return ballot(!p) == 0u; }
    template <class T> T shfl(T v, int src) const { return __shfl_sync(mask(), v, src, G); }
    template <class T> T shfl_xor(T v, int m) const { return __shfl_xor_sync(mask(), v, m, G); }
    template <class T> T shfl_up(T v, int d) const { return __shfl_up_sync(mask(), v, (unsigned)d, G); }
    template <class T> T shfl_down(T v, int d) const { return __shfl_down_sync(mask(), v, (unsigned)d, G); }
    void sync() const { emu::warp_barrier(mask(), true); }
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
#define blockIdx (emu::bid3())
#define gridDim (emu::gdim())
#define blockDim (emu::bdim())
// __shared__ variables of a kernel: one CTA runs at a time, so a function-level static is the CTA's shared memory
#undef __shared__
#define __shared__ static
