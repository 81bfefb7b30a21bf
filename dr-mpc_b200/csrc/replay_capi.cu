// replay_capi.cu -- device-resident replay ring fed straight from the env's double-buffered output slab
// (include/dre.h: dre_replay_*). Replaces ReplayBuffer.insert / sample_from_buffer (reference scripts/utils/storage.py:54-98,
// 101-133) and the numpy -> torch -> cuda conversion in front of it (scripts/online_continuous_task.py:136,148,187;
// scripts/utils/misc.py:7-21) for a batch of E envs: ONE selection kernel + ONE gather kernel per step instead of a
// roll of every tensor plus one row write per key.
//
// Layout: the reference's, one array [capacity][...] per observation key for S and for S_prime plus rewards [N][1],
// actions [N][2], done [N][1] (storage.py:21-35), in float64 (the reference's) or float32 (each float64 rounded once; the
// networks cast to float32 anyway, scripts/models/model.py). The ring overwrites the oldest rows instead of rolling
// (storage.py:56-75); chronological index j of the reference = slot (head + j) % capacity once full.
//
// Both kernels are pure copies: HBM-bound, 8-byte coalesced accesses along each row, no reuse.
#include <new>
#include "env_launch.cuh"

namespace {
enum { RP_OBS = 8, RP_FIELDS = RP_OBS * 2 + 3, RP_THREADS = 256, RP_UNROLL = 2, RP_ROWS = 4, RP_SEL = 1024 };

struct RingState {          // device-resident bookkeeping: an insert never synchronises with the host
    long long head;         // slot the next row goes to
    long long valid;        // rows holding data (<= capacity)
    long long total;        // rows ever inserted
    int n_last;             // rows selected by the latest insert
    int first_last;         // rank of the first selected row that fits (n_last > capacity keeps the newest)
    long long base_last;    // head before the latest insert
    unsigned arrived;       // CTAs of the running selection that have published their keep words
};

// one copy job per field: dst[dst_row * w + k] = src[src_row * w + k], k < w. A row is cut into UNITS: pairs of elements
// (one 16-byte access for float64) in the fields of even width, single elements in the others.
struct FieldTable {
    const void *src[RP_FIELDS];
    void *dst[RP_FIELDS];
    int w[RP_FIELDS];
    int uoff[RP_FIELDS + 1];  // prefix sums of the units per field: unit j belongs to field f with uoff[f] <= j < uoff[f+1]
    int nf;
};

// keep[e] = (mask == nullptr || mask[e]) && !(skip_scripted && done_flags[e][12]); rows[] = the kept envs in env order.
// Every CTA publishes the keep bits of its 1024 envs as 32 ballot words; the LAST CTA to arrive scans the words (2 KB for
// 16384 envs instead of 256 KB of strided flag bytes), writes rows[] and advances the ring.
__global__ void __launch_bounds__(RP_SEL) replay_select_kernel(int E, long long capacity, const uint8_t *mask, const uint8_t *done_flags,
                                                               int skip_scripted, unsigned *words, int *rows, RingState *st)
{
    __shared__ int warp_sum[32];
    __shared__ int running;
    __shared__ bool last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int e = blockIdx.x * RP_SEL + tid;
    bool keep = e < E;
    if (keep && mask) keep = mask[e] != 0;
    if (keep && skip_scripted) keep = done_flags[(size_t)e * 16 + 12] == 0;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) words[e >> 5] = b;
    __syncthreads();
    if (tid == 0) {                       // one cumulative fence per CTA (the barrier ordered the CTA's word stores before it)
        __threadfence();
        last = atomicAdd(&st->arrived, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (!last) return;
    if (tid == 0) running = 0;
    __syncthreads();
    const int nwords = gridDim.x * 32;
    for (int wbase = 0; wbase < nwords; wbase += RP_SEL) {
        unsigned w = wbase + tid < nwords ? __ldcg(&words[wbase + tid]) : 0u;
        const int mine = __popc(w);
        int incl = mine;                                       // inclusive scan over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int before = running + incl - mine;
        for (int q = 0; q < wid; ++q) before += warp_sum[q];
        const int e0 = (wbase + tid) * 32;
        while (w) {
            rows[before++] = e0 + __ffs(w) - 1;
            w &= w - 1;
        }
        __syncthreads();
        if (tid == RP_SEL - 1) running = before;               // the last thread's end = everything kept so far
        __syncthreads();
    }
    if (tid == 0) {
        const int n = running;
        const int first = n > capacity ? (int)(n - capacity) : 0;   // only the newest `capacity` rows can survive
        const long long m = n - first;
        st->n_last = n;
        st->first_last = first;
        st->base_last = st->head;
        st->head = (st->head + m) % capacity;
        st->valid = st->valid + m > capacity ? capacity : st->valid + m;
        st->total += m;
        st->arrived = 0;
    }
}

template <class T> struct alignas(2 * sizeof(T)) RpPair { T x, y; };

// MODE 0 (insert): row r < n_last - first_last: src row = rows[first_last + r], dst row = (base_last + r) mod capacity;
//                  the three transition scalars come from their own source layouts.
// MODE 1 (sample): row r < n: src row = ring slot of chronological index idx[r], dst row = r.
// Unit j of a row (all fields of S and S' laid end to end) is copied by thread j mod 256 for EVERY row: the (field, offset)
// of a thread's RP_UNROLL units is looked up once per tile, outside the row loop; one trip then issues
// RP_ROWS * RP_UNROLL independent 16-byte loads per thread before the first store -- the copy is bound by memory latency
// x bytes in flight, not by instructions.
template <class Tsrc, class Tdst, int MODE>
__global__ void __launch_bounds__(RP_THREADS, 3) replay_gather_kernel(const __grid_constant__ FieldTable T, long long capacity,
                                                                      const RingState *st, const int *rows, const long long *idx,
                                                                      long long n_rows, const double *act_src, const double *rew_src,
                                                                      const uint8_t *done_src, Tdst *act_dst, Tdst *rew_dst, Tdst *done_dst)
{
    const int nf = T.nf, U = T.uoff[nf];
    long long n = n_rows, head = 0, valid = 0, base = 0;
    int first = 0;
    if (MODE == 0) {
        first = st->first_last;
        n = st->n_last - first;
        base = st->base_last;
    } else {
        head = st->head;
        valid = st->valid;
    }
    constexpr int TILE = RP_UNROLL * RP_THREADS;
    for (int t0 = 0; t0 < U; t0 += TILE) {
        const Tsrc *sp[RP_UNROLL];
        Tdst *dp[RP_UNROLL];
        int w[RP_UNROLL];       // field width; 0 = no unit; negative = a single-element unit
#pragma unroll
        for (int u = 0; u < RP_UNROLL; ++u) {
            const int j = t0 + u * RP_THREADS + (int)threadIdx.x;
            sp[u] = nullptr; dp[u] = nullptr; w[u] = 0;
            if (j < U) {
                int f = 0;
                for (int q = 1; q < nf; ++q) f += j >= T.uoff[q] ? 1 : 0;
                const int wf = T.w[f];
                const bool pair = (wf & 1) == 0;
                const int k = (j - T.uoff[f]) * (pair ? 2 : 1);
                sp[u] = (const Tsrc *)T.src[f] + k;
                dp[u] = (Tdst *)T.dst[f] + k;
                w[u] = pair ? wf : -wf;
            }
        }
        for (long long r0 = blockIdx.x; r0 < n; r0 += RP_ROWS * (long long)gridDim.x) {
            long long srow[RP_ROWS], drow[RP_ROWS];
#pragma unroll
            for (int h = 0; h < RP_ROWS; ++h) {
                const long long r = r0 + h * (long long)gridDim.x;
                srow[h] = -1; drow[h] = 0;
                if (r >= n) continue;
                if (MODE == 0) {
                    srow[h] = rows[first + r];
                    drow[h] = base + r;                     // base < capacity and r < capacity
                    if (drow[h] >= capacity) drow[h] -= capacity;
                } else {
                    srow[h] = idx[r];
                    if (valid >= capacity) {                // full ring: chronological 0 = the slot at head
                        srow[h] += head;
                        if (srow[h] >= capacity) srow[h] -= capacity;
                    }
                    drow[h] = r;
                }
            }
            RpPair<Tsrc> v[RP_ROWS][RP_UNROLL];
#pragma unroll
            for (int h = 0; h < RP_ROWS; ++h)
#pragma unroll
                for (int u = 0; u < RP_UNROLL; ++u) {
                    if (srow[h] < 0 || w[u] == 0) continue;
                    if (w[u] > 0) v[h][u] = *(const RpPair<Tsrc> *)(sp[u] + srow[h] * w[u]);
                    else v[h][u].x = sp[u][srow[h] * -w[u]];
                }
#pragma unroll
            for (int h = 0; h < RP_ROWS; ++h)
#pragma unroll
                for (int u = 0; u < RP_UNROLL; ++u) {
                    if (srow[h] < 0 || w[u] == 0) continue;
                    if (w[u] > 0) {
                        RpPair<Tdst> o;
                        o.x = (Tdst)v[h][u].x; o.y = (Tdst)v[h][u].y;
                        *(RpPair<Tdst> *)(dp[u] + drow[h] * w[u]) = o;
                    } else {
                        dp[u][drow[h] * -w[u]] = (Tdst)v[h][u].x;
                    }
                }
            if (MODE == 0 && t0 == 0 && threadIdx.x >= RP_THREADS - 4 * RP_ROWS) {
                // A [E][2], R = rewards[e][2] (HA_and_PT env.py:413), done_for_replay = flags[e][3]
                const int q = threadIdx.x - (RP_THREADS - 4 * RP_ROWS), t = q & 3, hh = q >> 2;
                long long sr = -1, dr = 0;
#pragma unroll
                for (int h = 0; h < RP_ROWS; ++h)
                    if (h == hh) { sr = srow[h]; dr = drow[h]; }
                if (sr >= 0) {
                    if (t < 2) act_dst[dr * 2 + t] = (Tdst)act_src[sr * 2 + t];
                    else if (t == 2) rew_dst[dr] = (Tdst)rew_src[sr * 3 + 2];
                    else done_dst[dr] = (Tdst)(done_src[sr * 16 + 3] ? 1.0 : 0.0);
                }
            }
        }
    }
}

struct ObsWidths { int w[RP_OBS]; };

void *const *obs_ptrs(const dre_replay_fields &f) { return (void *const *)&f; }
}  // namespace

struct dre_replay {
    dre_env *env = nullptr;
    long long capacity = 0;
    int dtype = 0;                  // DRE_REPLAY_F64 | DRE_REPLAY_F32
    int E = 0;
    ObsWidths ow;
    void *block = nullptr;          // the ring: one allocation
    size_t block_bytes = 0;
    dre_replay_view ring;
    RingState *st = nullptr;        // device
    int *rows = nullptr;            // device [E]
    unsigned *words = nullptr;      // device [ceil(E / 1024) * 32] keep bits of the running selection
};

namespace {
size_t elem_size(int dtype) { return dtype == DRE_REPLAY_F32 ? 4 : 8; }

void obs_widths(const dre_env_config &c, int n_mpc_act, ObsWidths &ow)
{
    const int LB = c.lookback, Nh = c.num_humans;
    // order of dre_replay_fields: pt_state, mpc_actions, robot_node, past_robot_vel, percent_episode, spatial_edges,
    // human_valid, detected_humans
    ow.w[0] = 4 * (c.lookahead + c.lookbehind); ow.w[1] = n_mpc_act; ow.w[2] = 2 * LB; ow.w[3] = 2 * LB; ow.w[4] = 1;
    ow.w[5] = Nh * 2 * (LB + 1); ow.w[6] = Nh; ow.w[7] = 1;
}

void slab_obs(const dre_step_out &o, const void *p[RP_OBS])
{
    p[0] = o.pt_state; p[1] = o.mpc_actions; p[2] = o.robot_node; p[3] = o.past_robot_vel; p[4] = o.percent_episode;
    p[5] = o.spatial_edges; p[6] = o.human_valid; p[7] = o.detected_humans;
}

void carve_view(char *base, size_t &off, long long rows, size_t es, const ObsWidths &ow, dre_replay_view &v)
{
    auto take = [&](size_t n) {
        off = (off + 255) / 256 * 256;
        void *p = base ? base + off : nullptr;
        off += n * es;
        return p;
    };
    void **s = (void **)&v.S, **sp = (void **)&v.S_prime;
    for (int f = 0; f < RP_OBS; ++f) s[f] = take((size_t)rows * ow.w[f]);
    for (int f = 0; f < RP_OBS; ++f) sp[f] = take((size_t)rows * ow.w[f]);
    v.rewards = take((size_t)rows);
    v.actions = take((size_t)rows * 2);
    v.done = take((size_t)rows);
    off = (off + 255) / 256 * 256;
}

// units per field + alignment of the paired (even-width) fields: base pointers on 16 bytes (row strides then are too)
bool finish_table(FieldTable &T)
{
    T.uoff[0] = 0;
    for (int f = 0; f < RP_FIELDS; ++f) {
        const bool pair = (T.w[f] & 1) == 0;
        T.uoff[f + 1] = T.uoff[f] + (f < T.nf ? (pair ? T.w[f] / 2 : T.w[f]) : 0);
        if (f < T.nf && pair && T.w[f] > 0 && ((((uintptr_t)T.src[f]) | ((uintptr_t)T.dst[f])) & 15)) return false;
    }
    return true;
}

template <int MODE, class Tsrc, class Tdst>
int launch_gather(const FieldTable &T, const dre_replay *rb, const long long *idx, long long n_rows, const double *act,
                  const double *rew, const uint8_t *done, void *act_dst, void *rew_dst, void *done_dst, cudaStream_t s)
{
    // persistent grid: exactly the CTAs that are resident at once, every row a grid-stride item (no partial second wave)
    static int resident = 0;   // per instantiation
    if (!resident) {
        int dev = 0, sms = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        DRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, replay_gather_kernel<Tsrc, Tdst, MODE>, RP_THREADS, 0));
        resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    long long grid = resident;
    if (grid > n_rows) grid = n_rows;
    if (grid < 1) return DRE_OK;
    replay_gather_kernel<Tsrc, Tdst, MODE><<<(unsigned)grid, RP_THREADS, 0, s>>>(T, rb->capacity, rb->st, rb->rows, idx, n_rows, act, rew,
                                                                               done, (Tdst *)act_dst, (Tdst *)rew_dst, (Tdst *)done_dst);
    DRE_CUDA_TRY(cudaGetLastError());
    return DRE_OK;
}
}  // namespace

extern "C" int dre_replay_create(dre_env *env, int64_t capacity, int32_t dtype, dre_replay **out)
{
    if (!env || !out || capacity < 1 || (dtype != DRE_REPLAY_F64 && dtype != DRE_REPLAY_F32)) return DRE_ERR_INVALID;
    dre_replay *rb = new (std::nothrow) dre_replay;
    if (!rb) return DRE_ERR_NOMEM;
    const dre_env_config &c = dre_env_config_of(env);
    rb->env = env; rb->capacity = capacity; rb->dtype = dtype; rb->E = c.num_envs;
    obs_widths(c, dre_env_mpc_actions_width(env), rb->ow);
    size_t bytes = 0;
    carve_view(nullptr, bytes, capacity, elem_size(dtype), rb->ow, rb->ring);
    cudaError_t e;
    if ((e = cudaMalloc(&rb->block, bytes)) != cudaSuccess || (e = cudaMemset(rb->block, 0, bytes)) != cudaSuccess ||
        (e = cudaMalloc((void **)&rb->st, sizeof(RingState))) != cudaSuccess || (e = cudaMemset(rb->st, 0, sizeof(RingState))) != cudaSuccess ||
        (e = cudaMalloc((void **)&rb->rows, (size_t)rb->E * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc((void **)&rb->words, (size_t)((rb->E + RP_SEL - 1) / RP_SEL) * 32 * sizeof(unsigned))) != cudaSuccess) {
        dre_set_cuda_error(e, __FILE__, __LINE__);
        dre_replay_destroy(rb);
        return e == cudaErrorMemoryAllocation ? DRE_ERR_NOMEM : DRE_ERR_CUDA;
    }
    rb->block_bytes = bytes;
    size_t off = 0;
    carve_view((char *)rb->block, off, capacity, elem_size(dtype), rb->ow, rb->ring);
    rb->ring.rows = capacity; rb->ring.dtype = dtype;
    *out = rb;
    return DRE_OK;
}

extern "C" int dre_replay_destroy(dre_replay *rb)
{
    if (!rb) return DRE_OK;
    cudaDeviceSynchronize();
    cudaFree(rb->block); cudaFree(rb->st); cudaFree(rb->rows); cudaFree(rb->words);
    delete rb;
    return DRE_OK;
}

extern "C" int dre_replay_storage(dre_replay *rb, dre_replay_view *view, int32_t *widths8, size_t *bytes)
{
    if (!rb) return DRE_ERR_INVALID;
    if (view) *view = rb->ring;
    if (widths8) for (int f = 0; f < RP_OBS; ++f) widths8[f] = rb->ow.w[f];
    if (bytes) *bytes = rb->block_bytes;
    return DRE_OK;
}

extern "C" int dre_replay_insert(dre_replay *rb, const double *actions, const uint8_t *keep_mask, int32_t skip_scripted, void *stream)
{
    if (!rb) return DRE_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int cur = dre_env_current_slab(rb->env);
    dre_step_out now, before;
    int rc = dre_env_outputs_at(rb->env, cur, &now);
    if (rc) return rc;
    if ((rc = dre_env_outputs_at(rb->env, cur ^ 1, &before))) return rc;
    dre_env_note_stream(rb->env, s);   // the gather reads both slabs: a dre_env_step_host that follows (own streams) must wait for it
    replay_select_kernel<<<(rb->E + RP_SEL - 1) / RP_SEL, RP_SEL, 0, s>>>(rb->E, rb->capacity, keep_mask, now.done_flags, skip_scripted ? 1 : 0,
                                                                        rb->words, rb->rows, rb->st);
    DRE_CUDA_TRY(cudaGetLastError());
    FieldTable T;
    const void *ps[RP_OBS], *pn[RP_OBS];
    slab_obs(before, ps);   // S  = the observation the action was chosen on: the previous step's slab
    slab_obs(now, pn);      // S' = the latest outputs
    T.nf = 2 * RP_OBS;
    for (int f = 0; f < RP_OBS; ++f) {
        T.src[f] = ps[f]; T.dst[f] = obs_ptrs(rb->ring.S)[f]; T.w[f] = rb->ow.w[f];
        T.src[RP_OBS + f] = pn[f]; T.dst[RP_OBS + f] = obs_ptrs(rb->ring.S_prime)[f]; T.w[RP_OBS + f] = rb->ow.w[f];
    }
    for (int f = T.nf; f < RP_FIELDS; ++f) { T.src[f] = nullptr; T.dst[f] = nullptr; T.w[f] = 0; }
    if (!finish_table(T)) return DRE_ERR_INVALID;
    const double *act = actions ? actions : now.actions_used;
    const long long n_max = rb->E < rb->capacity ? rb->E : rb->capacity;
    if (rb->dtype == DRE_REPLAY_F32)
        return launch_gather<0, double, float>(T, rb, nullptr, n_max, act, now.rewards, now.done_flags, rb->ring.actions, rb->ring.rewards,
                                               rb->ring.done, s);
    return launch_gather<0, double, double>(T, rb, nullptr, n_max, act, now.rewards, now.done_flags, rb->ring.actions, rb->ring.rewards,
                                            rb->ring.done, s);
}

extern "C" int dre_replay_count(dre_replay *rb, int64_t *valid, int64_t *head, int64_t *last_inserted, void *stream)
{
    if (!rb) return DRE_ERR_INVALID;
    RingState h;
    DRE_CUDA_TRY(cudaMemcpyAsync(&h, rb->st, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DRE_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (valid) *valid = h.valid;
    if (head) *head = h.head;
    if (last_inserted) *last_inserted = h.n_last - h.first_last;
    return DRE_OK;
}

extern "C" int dre_replay_clear(dre_replay *rb, void *stream)
{
    if (!rb) return DRE_ERR_INVALID;
    DRE_CUDA_TRY(cudaMemsetAsync(rb->st, 0, sizeof(RingState), (cudaStream_t)stream));
    return DRE_OK;
}

extern "C" int dre_replay_sample(dre_replay *rb, const int64_t *indices, int64_t n, int32_t out_dtype, const dre_replay_view *batch,
                                 void *stream)
{
    if (!rb || !indices || !batch || n < 0 || (out_dtype != DRE_REPLAY_F64 && out_dtype != DRE_REPLAY_F32)) return DRE_ERR_INVALID;
    if (out_dtype == DRE_REPLAY_F64 && rb->dtype == DRE_REPLAY_F32) return DRE_ERR_INVALID;   // a float32 ring cannot give float64 back
    if (n == 0) return DRE_OK;
    FieldTable T;
    T.nf = RP_FIELDS;
    for (int f = 0; f < RP_OBS; ++f) {
        T.src[f] = obs_ptrs(rb->ring.S)[f]; T.dst[f] = obs_ptrs(batch->S)[f]; T.w[f] = rb->ow.w[f];
        T.src[RP_OBS + f] = obs_ptrs(rb->ring.S_prime)[f]; T.dst[RP_OBS + f] = obs_ptrs(batch->S_prime)[f]; T.w[RP_OBS + f] = rb->ow.w[f];
    }
    const int a = 2 * RP_OBS;
    T.src[a] = rb->ring.rewards; T.dst[a] = batch->rewards; T.w[a] = 1;
    T.src[a + 1] = rb->ring.actions; T.dst[a + 1] = batch->actions; T.w[a + 1] = 2;
    T.src[a + 2] = rb->ring.done; T.dst[a + 2] = batch->done; T.w[a + 2] = 1;
    for (int f = 0; f < T.nf; ++f)
        if (!T.dst[f]) return DRE_ERR_INVALID;
    if (!finish_table(T)) return DRE_ERR_INVALID;   // the paired fields of `batch` must be 16-byte aligned
    cudaStream_t s = (cudaStream_t)stream;
    const long long *idx = (const long long *)indices;
    if (rb->dtype == DRE_REPLAY_F32) return launch_gather<1, float, float>(T, rb, idx, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
    if (out_dtype == DRE_REPLAY_F32) return launch_gather<1, double, float>(T, rb, idx, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
    return launch_gather<1, double, double>(T, rb, idx, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
}
