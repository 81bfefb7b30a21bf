// env_capi.cu -- extern "C" entry points of the batched environment (include/dre.h: dre_env_*).
// Replaces HAAndPTEnv.reset/step (reference environment/HA_and_PT/human_avoidance_and_path_tracking_env.py:243-416).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <vector>
#include <dlfcn.h>
#include "env_launch.cuh"

namespace {
struct NcclApi {
    // nccl.h: ncclUniqueId is a 128-byte struct passed BY VALUE to ncclCommInitRank
    struct UniqueId { char internal[128]; };
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return api;
    api.GetUniqueId = (int (*)(NcclApi::UniqueId *))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, NcclApi::UniqueId, int))dlsym(lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
    return api;
}
int nccl_fail(int rc, int line)
{
    const NcclApi &A = nccl_api();
    char msg[256];
    snprintf(msg, sizeof msg, "NCCL error %d (%s) at %s:%d", rc, A.GetErrorString ? A.GetErrorString(rc) : "?", __FILE__, line);
    dre_set_error_text(msg);
    return DRE_ERR_CUDA;
}
enum { kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };   // nccl.h: ncclDataType_t / ncclRedOp_t
}  // namespace

struct dre_env {
    dre_env_config cfg;
    EnvDev D;
    dre_mpc *mpc = nullptr;
    void *state_block = nullptr;   // one allocation for the whole SoA state
    // Outputs of a step: one contiguous slab, DOUBLE-BUFFERED -- step t writes slab t & 1, so the S a caller got from
    // step t-1 (zero-copy views) is still intact while it holds S' of step t: `insert(S, a, S'); S = S'` needs no clone.
    void *slabs[2] = {nullptr, nullptr};
    dre_step_out outs[2];
    EnvOutDev outdev[2];
    int cur = 0;                   // slab holding the latest outputs
    void *slab = nullptr;          // = slabs[cur]
    size_t slab_bytes = 0;
    dre_step_out out;              // = outs[cur]
    double *d_actions = nullptr;   // staging for step_host
    cudaStream_t own_stream = nullptr;
    // step_host pipeline: one crowd-half launch; the last CTA of every env chunk raises a flag in mapped host memory,
    // the host thread polls the flags and queues that chunk's D2H copy on the copy stream
    static const int MAX_CHUNKS = 32;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_front = nullptr, ev_pt = nullptr, ev_done = nullptr;
    // the MPC solve runs on its own stream next to the crowd half: both depend on pt_step only, and each has a long,
    // thinly occupied tail (the solve's slow problems, the crowd half's last partial wave) the other one fills
    cudaStream_t mpc_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_sync = nullptr; // orders step_host (own streams) after earlier device-path calls on the caller's stream
    cudaStream_t last_stream = nullptr;
    bool last_stream_set = false;
    int *h_flags = nullptr;        // mapped pinned host memory [MAX_CHUNKS]
    int *d_flags = nullptr;        // its device alias
    int *d_chunk_count = nullptr;  // device [MAX_CHUNKS]
    int step_seq = 0;
    int host_chunks = 8;
    int host_mode = 0;             // dre_env_set_host_outputs: 0 whole float64 slab, 1 observations as float32, 2 control fields only
    float *d_stage32 = nullptr;    // mode 1: float32 images of the observation fields, at the fields' own byte offsets
    bool was_reset = false;
    void *stats_comm = nullptr;    // ncclComm_t of dre_env_stats_comm_init (the only collective of the engine)
    // per-kernel timing (bench.py's roofline leg)
    bool profiling = false;
    static const int PROF_RING = 64;
    cudaEvent_t ev[PROF_RING][5];
    int ev_used = 0;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_steps = 0;
};

static int prof_drain(dre_env *h)
{
    for (int i = 0; i < h->ev_used; ++i) {
        DRE_CUDA_TRY(cudaEventSynchronize(h->ev[i][4]));
        static const int slot[4] = {1, 0, 2, 3};   // launch order pt, ha, mpc, stats -> {ha, pt, mpc, stats}
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            DRE_CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[i][k], h->ev[i][k + 1]));
            h->prof_ms[slot[k]] += ms;
        }
        h->prof_steps++;
    }
    h->ev_used = 0;
    return DRE_OK;
}

namespace {
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base((char *)b) {}
    template <class T> T *take(size_t n)
    {
        off = (off + 255) / 256 * 256;
        T *p = base ? (T *)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

void carve_state(EnvDev &D, void *base, size_t &bytes)
{
    Carver c(base);
    const size_t E = D.E, Nh = D.Nh, H = D.LB + 1, LB = D.LB;
    D.hpos = c.take<double2>(E * Nh);
    D.hvel = c.take<float2>(E * Nh);
    D.hgoal = c.take<double2>(E * Nh);
    D.hstatic = c.take<uint8_t>(E * Nh);
    D.hhist = c.take<double2>(E * Nh * H);
    D.hvalid = c.take<int32_t>(E * Nh);
    D.rpose = c.take<double>(E * 4);
    D.rprev = c.take<double>(E * 4);
    D.rhist = c.take<double>(E * H * 4);
    D.rpastv = c.take<double>(E * LB * 2);
    D.rvalid = c.take<int32_t>(E);
    D.gtime = c.take<double>(E);
    D.path_idx = c.take<int32_t>(E);
    D.loc_to_start = c.take<uint8_t>(E);
    D.step_count = c.take<int64_t>(E);
    D.reset_epoch = c.take<int32_t>(E);
    D.mpc_interp = c.take<double>(E);
    D.sr_state = c.take<int32_t>(E * 4);
    D.hvel_next = c.take<float2>(E * Nh);
    D.goal_changed = c.take<uint8_t>(E * Nh);
    D.cache_valid = c.take<uint8_t>(E);
    D.lp_class = c.take<uint8_t>(E * Nh);
    D.stat_counters = c.take<unsigned long long>(16);
    D.stat_sums = c.take<double>(4);
    bytes = (c.off + 255) / 256 * 256;
}

void carve_slab(const EnvDev &D, int n_mpc_act, void *base, dre_step_out &o, EnvOutDev &od, size_t &bytes)
{
    Carver c(base);
    const size_t E = D.E, Nh = D.Nh, H = D.LB + 1, LB = D.LB;
    o.pt_state = od.pt_state = c.take<double>(E * 4 * (D.lookahead + D.lookbehind));
    o.actions_used = od.actions_used = c.take<double>(E * 2);
    o.mpc_actions = od.mpc_actions = c.take<double>(E * n_mpc_act);
    o.robot_node = od.robot_node = c.take<double>(E * 2 * LB);
    o.past_robot_vel = od.past_robot_vel = c.take<double>(E * 2 * LB);
    o.percent_episode = od.percent_episode = c.take<double>(E);
    o.spatial_edges = od.spatial_edges = c.take<double>(E * Nh * 2 * H);
    o.human_valid = od.human_valid = c.take<double>(E * Nh);
    o.detected_humans = od.detected_humans = c.take<double>(E);
    o.rewards = od.rewards = c.take<double>(E * 3);
    o.info = od.info = c.take<double>(E * 8);
    o.done_bits = od.done_bits = c.take<uint32_t>(E);
    o.mpc_status = od.mpc_status = c.take<dre_mpc_status>(E);
    o.done_flags = od.done_flags = c.take<uint8_t>(E * 16);
    bytes = (c.off + 255) / 256 * 256;
    o.slab = base;
    o.slab_bytes = bytes;
}

// make slab `b` the one the next launches write (and dre_env_outputs reports); the other one holds the previous step
void use_slab(dre_env *h, int b)
{
    h->cur = b;
    h->slab = h->slabs[b];
    h->out = h->outs[b];
    h->D.out = h->outdev[b];
    h->D.prev_done_bits = h->outdev[b ^ 1].done_bits;
}

// The copy stream also runs the float32 conversion kernels of host-output mode 1: highest priority, so that their few
// CTAs take the next CTA slot that frees up instead of queueing behind the crowd half's pending CTAs.
cudaError_t create_copy_stream(cudaStream_t *st)
{
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return e;
    return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, hi);
}

int n_mpc_actions(const dre_env_config &c)
{
    const int a = c.mpc_actions_in_input > 0 ? c.mpc_actions_in_input : 4;
    return 2 * (c.mpc.horizon < a ? c.mpc.horizon : a);
}
}  // namespace

extern "C" int dre_env_create(const dre_env_config *cfg, dre_env **out)
{
    if (!cfg || !out) return DRE_ERR_INVALID;
    if (cfg->num_envs <= 0 || cfg->num_humans <= 0 || cfg->num_humans > DRE_MAX_HUMANS || cfg->lookback < 1 || cfg->lookback > 32)
        return DRE_ERR_INVALID;
    if (cfg->lookahead + cfg->lookbehind <= 0 || !(cfg->time_step > 0)) return DRE_ERR_INVALID;
    if (cfg->soft_reset_on_device && !cfg->auto_path_switch) return DRE_ERR_INVALID;
    if (cfg->num_static_humans < 0 || cfg->num_static_humans > 4 || cfg->num_static_humans >= cfg->num_humans) return DRE_ERR_INVALID;
    dre_env *h = new (std::nothrow) dre_env();
    if (!h) return DRE_ERR_NOMEM;
    h->cfg = *cfg;
    EnvDev &D = h->D;
    memset(&D, 0, sizeof D);
    D.E = cfg->num_envs; D.Nh = cfg->num_humans; D.LB = cfg->lookback;
    D.orca = cfg->orca;
    D.dt = cfg->time_step; D.time_limit = cfg->time_limit; D.human_radius = cfg->human_radius; D.robot_radius = cfg->robot_radius;
    D.circle_radius = cfg->circle_radius; D.circle_radius_humans = cfg->circle_radius_humans;
    D.end_goal_changing = cfg->end_goal_changing;
    D.collision_penalty = cfg->collision_penalty; D.safety_dist = cfg->safety_dist; D.safety_penalty = cfg->safety_penalty;
    D.dist_radius = cfg->disturbance_radius; D.dist_factor_v = cfg->disturbance_factor_v; D.dist_factor_th = cfg->disturbance_factor_th;
    D.act_min_vel = cfg->actuation_min_vel; D.act_penalty = cfg->actuation_penalty;
    D.pt_overall_scaling = cfg->pt_overall_scaling; D.goal_radius = cfg->goal_radius; D.goal_angle = cfg->goal_angle;
    D.goal_reward = cfg->goal_reward; D.path_adv_scaling = cfg->path_advancement_scaling; D.dev_xy = cfg->deviation_xy;
    D.dev_th = cfg->deviation_th; D.dev_scale = cfg->deviation_scale; D.corridor_penalty = cfg->corridor_penalty;
    D.corridor_max_dev = cfg->corridor_max_dev; D.eop_penalty = cfg->end_of_path_penalty;
    D.safety_corridor_penalty = cfg->safety_corridor_penalty; D.safety_corridor_max_dev = cfg->safety_corridor_max_dev;
    D.lookahead = cfg->lookahead; D.lookbehind = cfg->lookbehind; D.auto_path_switch = cfg->auto_path_switch;
    D.soft_reset = cfg->soft_reset_on_device; D.orca_dedup = cfg->orca_dedup; D.lookahead_prune = cfg->orca_lookahead_prune;
    D.sensor_restriction = cfg->sensor_restriction; D.sensor_range = cfg->sensor_range;
    D.num_static = cfg->num_static_humans;
    for (int i = 0; i < 8; ++i) D.static_pos[i] = cfg->static_pos[i];
    D.seed = cfg->seed; D.env_id_offset = cfg->env_id_offset;

    int rc = dre_mpc_create(&cfg->mpc, cfg->num_envs, &h->mpc);
    if (rc) { delete h; return rc; }
    size_t sbytes = 0, obytes = 0;
    carve_state(D, nullptr, sbytes);
    EnvOutDev od;
    carve_slab(D, n_mpc_actions(*cfg), nullptr, h->out, od, obytes);
    cudaError_t e;
    if ((e = cudaMalloc(&h->state_block, sbytes)) != cudaSuccess || (e = cudaMemset(h->state_block, 0, sbytes)) != cudaSuccess ||
        (e = cudaMalloc(&h->slabs[0], obytes)) != cudaSuccess || (e = cudaMemset(h->slabs[0], 0, obytes)) != cudaSuccess ||
        (e = cudaMalloc(&h->slabs[1], obytes)) != cudaSuccess || (e = cudaMemset(h->slabs[1], 0, obytes)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_actions, (size_t)D.E * 2 * sizeof(double))) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = create_copy_stream(&h->copy_stream)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_pt, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->mpc_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&h->h_flags, dre_env::MAX_CHUNKS * sizeof(int), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer((void **)&h->d_flags, h->h_flags, 0)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_chunk_count, dre_env::MAX_CHUNKS * sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(h->d_chunk_count, 0, dre_env::MAX_CHUNKS * sizeof(int))) != cudaSuccess) {
        dre_set_cuda_error(e, __FILE__, __LINE__);
        dre_env_destroy(h);
        return DRE_ERR_CUDA;
    }
    memset(h->h_flags, 0, dre_env::MAX_CHUNKS * sizeof(int));
    carve_state(D, h->state_block, sbytes);
    for (int b = 0; b < 2; ++b) carve_slab(D, n_mpc_actions(*cfg), h->slabs[b], h->outs[b], h->outdev[b], obytes);
    h->slab_bytes = obytes;
    use_slab(h, 0);
    D.mpc_iteration0 = dre_mpc_state_of(h->mpc).iteration0;
    *out = h;
    return DRE_OK;
}

extern "C" int dre_env_destroy(dre_env *h)
{
    if (!h) return DRE_OK;
    cudaDeviceSynchronize();
    dre_mpc_destroy(h->mpc);
    if (h->profiling) dre_env_set_profiling(h, 0);
    cudaFree(h->state_block); cudaFree(h->slabs[0]); cudaFree(h->slabs[1]); cudaFree(h->d_actions); cudaFree(h->d_stage32);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_front) cudaEventDestroy(h->ev_front);
    if (h->ev_pt) cudaEventDestroy(h->ev_pt);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->ev_sync) cudaEventDestroy(h->ev_sync);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->mpc_stream) cudaStreamDestroy(h->mpc_stream);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    cudaFree(h->d_chunk_count);
    if (h->stats_comm && nccl_api().ok) nccl_api().CommDestroy(h->stats_comm);
    delete h;
    return DRE_OK;
}

extern "C" int dre_env_set_paths(dre_env *h, const double *paths_host, int32_t num_paths, int32_t path_stride, const int32_t *lens_host)
{
    if (!h) return DRE_ERR_INVALID;
    const int rc = dre_mpc_set_paths(h->mpc, paths_host, num_paths, path_stride, lens_host);
    if (rc) return rc;
    const MpcDeviceState &st = dre_mpc_state_of(h->mpc);
    h->D.paths = st.paths; h->D.lens = st.lens; h->D.num_paths = st.num_paths; h->D.path_stride = st.path_stride;
    return DRE_OK;
}

extern "C" int dre_env_state(dre_env *h, dre_env_state_view *v)
{
    if (!h || !v) return DRE_ERR_INVALID;
    const EnvDev &D = h->D;
    v->hpos = (double *)D.hpos; v->hvel = (float *)D.hvel; v->hgoal = (double *)D.hgoal; v->hstatic = D.hstatic;
    v->hhist = (double *)D.hhist; v->hvalid = D.hvalid; v->rpose = D.rpose; v->rprev = D.rprev; v->rhist = D.rhist;
    v->rpastv = D.rpastv; v->rvalid = D.rvalid; v->gtime = D.gtime; v->path_idx = D.path_idx; v->loc_to_start = D.loc_to_start;
    v->step_count = D.step_count; v->sr_state = D.sr_state;
    return DRE_OK;
}

extern "C" int dre_env_outputs(dre_env *h, dre_step_out *out)
{
    if (!h || !out) return DRE_ERR_INVALID;
    *out = h->out;
    return DRE_OK;
}

extern "C" int dre_env_outputs_at(dre_env *h, int32_t which, dre_step_out *out)
{
    if (!h || !out || (which != 0 && which != 1)) return DRE_ERR_INVALID;
    *out = h->outs[which];
    return DRE_OK;
}

extern "C" int dre_env_current_slab(dre_env *h) { return h ? h->cur : DRE_ERR_INVALID; }

extern "C" dre_mpc *dre_env_mpc(dre_env *h) { return h ? h->mpc : nullptr; }

const dre_env_config &dre_env_config_of(const dre_env *h) { return h->cfg; }
int dre_env_mpc_actions_width(const dre_env *h) { return n_mpc_actions(h->cfg); }
void dre_env_note_stream(dre_env *h, cudaStream_t s) { h->last_stream = s; h->last_stream_set = true; }

static int env_solve_mpc(dre_env *h, const EnvDev &D, cudaStream_t s)
{
    const int nact = n_mpc_actions(h->cfg);
    int rc = dre_mpc_launch(dre_mpc_params_of(h->mpc), dre_mpc_state_of(h->mpc), D.path_idx, D.mpc_interp, D.rpose, 4,
                            D.out.mpc_actions, nact, nact, D.out.mpc_status, D.env_mask, s);
    if (rc) return rc;
    return dre_env_launch_mpc_stats(D, s);
}

extern "C" int dre_env_observe_masked(dre_env *h, const uint8_t *env_mask, int32_t solve_mpc, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    if (!h->D.paths) return DRE_ERR_STATE;
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s; h->last_stream_set = true;
    EnvDev D = h->D;
    D.env_mask = env_mask;
    int rc = dre_env_launch_ha(D, 2, 1, s);
    if (rc) return rc;
    rc = dre_env_launch_pt(D, 1, s);
    if (rc) return rc;
    return solve_mpc ? env_solve_mpc(h, D, s) : DRE_OK;
}

extern "C" int dre_env_observe(dre_env *h, int32_t solve_mpc, void *stream) { return dre_env_observe_masked(h, nullptr, solve_mpc, stream); }

extern "C" int dre_env_reset_masked(dre_env *h, const uint8_t *env_mask, const double *hpos_init, const double *hgoal_init, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    if (!h->D.paths) return DRE_ERR_STATE;
    if (env_mask && !h->was_reset) return DRE_ERR_STATE;   // the other envs must hold a valid episode
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s; h->last_stream_set = true;
    EnvDev D = h->D;
    D.env_mask = env_mask;
    int rc = dre_env_launch_reset(D, hpos_init, hgoal_init, s);
    if (rc) return rc;
    // S_PT = PT_env.reset(...) incl. the first MPC solve (HA_and_PT env.py:268)
    rc = dre_env_launch_pt(D, 1, s);
    if (rc) return rc;
    rc = env_solve_mpc(h, D, s);
    if (rc) return rc;
    // HA_env.reset + warm_start(lookback) (HA_and_PT env.py:274-276)
    for (int w = 0; w < D.LB; ++w) {
        rc = dre_env_launch_ha(D, 1, w == D.LB - 1 ? 1 : 0, s);
        if (rc) return rc;
    }
    h->was_reset = true;
    return DRE_OK;
}

extern "C" int dre_env_reset(dre_env *h, const double *hpos_init, const double *hgoal_init, void *stream)
{
    return dre_env_reset_masked(h, nullptr, hpos_init, hgoal_init, stream);
}

extern "C" int dre_env_crowd_step(dre_env *h, int32_t write_obs, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;
    return dre_env_launch_ha(h->D, 1, write_obs ? 1 : 0, (cudaStream_t)stream);
}

extern "C" int dre_env_switch_path(dre_env *h, const uint8_t *env_mask, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;
    EnvDev D = h->D;
    D.env_mask = env_mask;
    return dre_env_launch_switch_path(D, (cudaStream_t)stream);
}

extern "C" int dre_env_soft_reset_decide(dre_env *h, int32_t commit, uint8_t *scripted, double *actions, uint8_t *stuck, void *stream)
{
    if (!h || !scripted || !actions) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;
    return dre_env_launch_sr_decide(h->D, commit, scripted, actions, stuck, (cudaStream_t)stream);
}

extern "C" int dre_env_step(dre_env *h, const double *actions, void *stream)
{
    if (!h || !actions) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s; h->last_stream_set = true;
    use_slab(h, h->cur ^ 1);       // this step's outputs go to the other slab; the previous step's stay readable
    EnvDev D = h->D;
    D.actions = actions;
    // robot/path half, crowd half (which merges rewards / dones), MPC solve: see env_kernels.cu. The crowd half and the
    // solve only depend on pt_step; the solve goes last so that, on the host path, the crowd's copies overlap it.
    if (!h->profiling) {
        // measured at 16384 x 20 (mean of 300 steps): serial 1.14 ms, solve launched first 1.06 ms, crowd half first 1.04 ms
        static const int overlap = getenv("DRE_STEP_OVERLAP") ? atoi(getenv("DRE_STEP_OVERLAP")) : 2;   // 0 serial, 1 MPC first, 2 crowd first
        int rc = dre_env_launch_pt(D, 0, s);
        if (rc) return rc;
        if (overlap == 0) {
            rc = dre_env_launch_ha(D, 0, 1, s);
            if (rc) return rc;
            return env_solve_mpc(h, D, s);
        }
        DRE_CUDA_TRY(cudaEventRecord(h->ev_fork, s));
        DRE_CUDA_TRY(cudaStreamWaitEvent(h->mpc_stream, h->ev_fork, 0));
        if (overlap == 1) {
            rc = env_solve_mpc(h, D, h->mpc_stream);
            if (rc) return rc;
            rc = dre_env_launch_ha(D, 0, 1, s);
        } else {
            rc = dre_env_launch_ha(D, 0, 1, s);
            if (rc) return rc;
            rc = env_solve_mpc(h, D, h->mpc_stream);
        }
        if (rc) return rc;
        DRE_CUDA_TRY(cudaEventRecord(h->ev_join, h->mpc_stream));
        DRE_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
        return DRE_OK;
    }
    if (h->ev_used == dre_env::PROF_RING) { const int rc = prof_drain(h); if (rc) return rc; }
    cudaEvent_t *ev = h->ev[h->ev_used];
    const int nact = n_mpc_actions(h->cfg);
    // event order = launch order: pt_step, ha_step, mpc_solve, mpc_stats (prof_drain maps them back to the
    // documented {ha_step, pt_step, mpc_solve, mpc_stats} order)
    DRE_CUDA_TRY(cudaEventRecord(ev[0], s));
    int rc = dre_env_launch_pt(D, 0, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaEventRecord(ev[1], s));
    rc = dre_env_launch_ha(D, 0, 1, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaEventRecord(ev[2], s));
    rc = dre_mpc_launch(dre_mpc_params_of(h->mpc), dre_mpc_state_of(h->mpc), D.path_idx, D.mpc_interp, D.rpose, 4,
                        D.out.mpc_actions, nact, nact, D.out.mpc_status, nullptr, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaEventRecord(ev[3], s));
    rc = dre_env_launch_mpc_stats(D, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaEventRecord(ev[4], s));
    h->ev_used++;
    return DRE_OK;
}

extern "C" int dre_env_set_profiling(dre_env *h, int32_t on)
{
    if (!h) return DRE_ERR_INVALID;
    if (on && !h->profiling) {
        for (int i = 0; i < dre_env::PROF_RING; ++i)
            for (int k = 0; k < 5; ++k) DRE_CUDA_TRY(cudaEventCreate(&h->ev[i][k]));
    } else if (!on && h->profiling) {
        const int rc = prof_drain(h);
        if (rc) return rc;
        for (int i = 0; i < dre_env::PROF_RING; ++i)
            for (int k = 0; k < 5; ++k) cudaEventDestroy(h->ev[i][k]);
    }
    h->profiling = on != 0;
    return DRE_OK;
}

extern "C" int dre_env_kernel_ms(dre_env *h, double *ms4, int64_t *steps, int32_t reset)
{
    if (!h) return DRE_ERR_INVALID;
    if (h->profiling) { const int rc = prof_drain(h); if (rc) return rc; }
    if (ms4) for (int k = 0; k < 4; ++k) ms4[k] = h->prof_ms[k];
    if (steps) *steps = h->prof_steps;
    if (reset) { for (int k = 0; k < 4; ++k) h->prof_ms[k] = 0.0; h->prof_steps = 0; }
    return DRE_OK;
}

// step() on host buffers, pipelined: H2D actions -> robot/path half (its PT state travels during the MPC solve) -> MPC
// -> ONE crowd-half launch whose CTAs finish roughly in env order; the last CTA of each env chunk raises a flag in
// mapped host memory, this thread polls the flags and queues the chunk's observation copy, so of the whole D2H only the
// last chunk's share and the small per-env fields are exposed.
extern "C" int dre_env_step_host(dre_env *h, const double *actions_host, void *slab_host)
{
    if (!h || !actions_host || !slab_host) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    cudaStream_t s = h->own_stream, c = h->copy_stream;
    const int E = h->D.E, Nh = h->D.Nh, H = h->D.LB + 1;
    static const bool trace = getenv("DRE_HOST_TRACE") != nullptr;   // measurement aid: host-side timeline of one call in 64
    const bool tr = trace && (h->step_seq & 63) == 63;
    auto now_us = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t_start = tr ? now_us() : 0.0;
    double t_flag[dre_env::MAX_CHUNKS] = {}, t_enq = 0.0;
    if (h->last_stream_set) {   // earlier reset() / step() / observe() calls ran on the caller's stream: run after them
        DRE_CUDA_TRY(cudaEventRecord(h->ev_sync, h->last_stream));
        DRE_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_sync, 0));
        h->last_stream_set = false;
    }
    DRE_CUDA_TRY(cudaMemcpyAsync(h->d_actions, actions_host, (size_t)E * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    use_slab(h, h->cur ^ 1);
    EnvDev D = h->D;
    D.actions = h->d_actions;
    char *dev = (char *)h->slab, *host = (char *)slab_host;
    const int mode = h->host_mode;
    if (mode == 1 && !h->d_stage32) DRE_CUDA_TRY(cudaMalloc(&h->d_stage32, h->slab_bytes));
    // elements [first, first + count) of one float64 field of the slab to the host: as they are, or (mode 1, observation
    // fields) rounded once to float32 by a small kernel on the copy stream; the float32 image of a field starts at the
    // field's own byte offset of the host slab. `raw` runs (several control fields + padding) are copied byte for byte.
    auto send_field = [&](const double *field, size_t first, size_t count, bool is_obs) -> int {
        if (count == 0) return DRE_OK;
        const size_t foff = (const char *)field - dev;
        if (mode == 1 && is_obs) {
            float *st = reinterpret_cast<float *>(reinterpret_cast<char *>(h->d_stage32) + foff) + first;
            const int r = dre_env_launch_cvt32(field + first, st, count, c);
            if (r) return r;
            DRE_CUDA_TRY(cudaMemcpyAsync(host + foff + first * sizeof(float), st, count * sizeof(float), cudaMemcpyDeviceToHost, c));
        } else {
            DRE_CUDA_TRY(cudaMemcpyAsync(host + foff + first * sizeof(double), field + first, count * sizeof(double), cudaMemcpyDeviceToHost, c));
        }
        return DRE_OK;
    };
    auto send_raw = [&](const void *p0, const void *p1) -> int {
        const size_t off = (const char *)p0 - dev, nb = (const char *)p1 - (const char *)p0;
        if (nb) DRE_CUDA_TRY(cudaMemcpyAsync(host + off, dev + off, nb, cudaMemcpyDeviceToHost, c));
        return DRE_OK;
    };
    const size_t LB = h->D.LB, NPT = 4 * (size_t)(h->D.lookahead + h->D.lookbehind);
    int rc = dre_env_launch_pt(D, 0, s);
    if (rc) return rc;
    // the PT state (the second largest output) and the executed actions are final after pt_step
    DRE_CUDA_TRY(cudaEventRecord(h->ev_pt, s));
    DRE_CUDA_TRY(cudaStreamWaitEvent(c, h->ev_pt, 0));
    if (mode != 2) { rc = send_field(h->out.pt_state, 0, (size_t)E * NPT, true); if (rc) return rc; }
    rc = send_raw(h->out.actions_used, h->out.mpc_actions);
    if (rc) return rc;
    // crowd half BEFORE the MPC solve on this path: it produces 3/4 of the bytes, so its copies get the MPC solve's
    // duration as additional cover (neither half reads the other's outputs; the merge only needs pt_step's)
    const int epb = dre_env_ha_envs_per_cta(D);
    int chunks = h->host_chunks < dre_env::MAX_CHUNKS ? h->host_chunks : dre_env::MAX_CHUNKS;
    if (chunks < 1) chunks = 1;
    int per = ((E + chunks - 1) / chunks + epb - 1) / epb * epb;
    chunks = (E + per - 1) / per;
    h->step_seq = h->step_seq == 0x7fffffff ? 1 : h->step_seq + 1;
    if (mode != 2) { D.chunk_flags = h->d_flags; D.chunk_count = h->d_chunk_count; D.chunk_envs = per; D.step_seq = h->step_seq; }
    rc = dre_env_launch_ha(D, 0, 1, s);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaEventRecord(h->ev_front, s));     // crowd half done
    // The solve stays BEHIND the crowd half here, not next to it as in dre_env_step: this path is bound by the copies, and
    // a solve that shares the SMs delays the chunk flags (measured p50 1.19 ms serial vs 1.27 ms concurrent, also with the
    // solve on a lowest-priority stream: priorities do not stop its persistent CTAs from taking freed slots early).
    if (mode == 2) {   // nothing big travels: the solve runs NEXT to the crowd half, as in dre_env_step
        DRE_CUDA_TRY(cudaStreamWaitEvent(h->mpc_stream, h->ev_pt, 0));
        rc = env_solve_mpc(h, D, h->mpc_stream);
        if (rc) return rc;
        DRE_CUDA_TRY(cudaEventRecord(h->ev_join, h->mpc_stream));
        DRE_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
    } else {
        rc = env_solve_mpc(h, D, s);
        if (rc) return rc;
    }
    DRE_CUDA_TRY(cudaEventRecord(h->ev_done, s));      // step done
    if (tr) t_enq = now_us();
    if (mode != 2) {
        volatile int *flags = h->h_flags;
        const size_t per_env = (size_t)Nh * 2 * H;
        for (int i = 0; i < chunks; ++i) {
            const int e0 = i * per, e1 = e0 + per < E ? e0 + per : E;
            for (unsigned spin = 0; flags[i] != h->step_seq; ++spin) {
                if ((spin & 0x3ff) == 0x3ff) {   // the launch may have failed or finished without us seeing the flag yet
                    const cudaError_t q = cudaEventQuery(h->ev_front);
                    if (q == cudaSuccess) { if (flags[i] != h->step_seq) { h->h_flags[i] = h->step_seq; } break; }
                    if (q != cudaErrorNotReady) { dre_set_cuda_error(q, __FILE__, __LINE__); return DRE_ERR_CUDA; }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            if (tr) t_flag[i] = now_us();
            // the pair history of the crowd is 3/4 of the crowd half's bytes: one copy per chunk
            rc = send_field(h->out.spatial_edges, (size_t)e0 * per_env, (size_t)(e1 - e0) * per_env, true);
            if (rc) return rc;
        }
        // every CTA has written its observation: the small observation arrays go as two contiguous runs of the slab
        // (robot_node .. percent_episode, human_valid .. detected_humans) while the last CTAs finish their look-ahead pass
        if (mode == 0) {
            rc = send_raw(h->out.robot_node, h->out.spatial_edges);
            if (rc) return rc;
            rc = send_raw(h->out.human_valid, h->out.rewards);
            if (rc) return rc;
        } else {
            if ((rc = send_field(h->out.robot_node, 0, (size_t)E * 2 * LB, true)) || (rc = send_field(h->out.past_robot_vel, 0, (size_t)E * 2 * LB, true)) ||
                (rc = send_field(h->out.percent_episode, 0, (size_t)E, true)) || (rc = send_field(h->out.human_valid, 0, (size_t)E * Nh, true)) ||
                (rc = send_field(h->out.detected_humans, 0, (size_t)E, true))) return rc;
        }
    }
    // rewards, info, done bits once the crowd half is complete; MPC actions, MPC status and the done flags (which sit
    // behind the status in the slab) once the solve is
    DRE_CUDA_TRY(cudaStreamWaitEvent(c, h->ev_front, 0));
    rc = send_raw(h->out.rewards, h->out.mpc_status);
    if (rc) return rc;
    DRE_CUDA_TRY(cudaStreamWaitEvent(c, h->ev_done, 0));
    rc = send_raw(h->out.mpc_actions, h->out.robot_node);
    if (rc) return rc;
    rc = send_raw(h->out.mpc_status, dev + h->slab_bytes);
    if (rc) return rc;
    if (tr) {
        cudaEventSynchronize(h->ev_done);
        const double t_done = now_us();
        cudaStreamSynchronize(c);
        const double t_end = now_us();
        fprintf(stderr, "[host trace] enqueued %.0f us, chunk flags at", t_enq - t_start);
        for (int i = 0; i < chunks; ++i) fprintf(stderr, " %.0f", t_flag[i] - t_start);
        fprintf(stderr, ", step (crowd half, then MPC) done %.0f, copies done %.0f us\n", t_done - t_start, t_end - t_start);
    }
    DRE_CUDA_TRY(cudaStreamSynchronize(c));
    return DRE_OK;
}

extern "C" int dre_env_set_host_outputs(dre_env *h, int32_t mode)
{
    if (!h || mode < 0 || mode > 2) return DRE_ERR_INVALID;
    h->host_mode = mode;
    return DRE_OK;
}

// the D2H floor of this box for one step's slab: plain pinned copies, nothing else on the GPU
extern "C" int dre_env_d2h_probe(dre_env *h, void *slab_host, int32_t reps, double *ms_per_slab)
{
    if (!h || !slab_host || reps < 1 || !ms_per_slab) return DRE_ERR_INVALID;
    cudaEvent_t e0, e1;
    DRE_CUDA_TRY(cudaEventCreate(&e0));
    DRE_CUDA_TRY(cudaEventCreate(&e1));
    DRE_CUDA_TRY(cudaDeviceSynchronize());
    DRE_CUDA_TRY(cudaMemcpyAsync(slab_host, h->slab, h->slab_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    DRE_CUDA_TRY(cudaEventRecord(e0, h->copy_stream));
    for (int i = 0; i < reps; ++i) DRE_CUDA_TRY(cudaMemcpyAsync(slab_host, h->slab, h->slab_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    DRE_CUDA_TRY(cudaEventRecord(e1, h->copy_stream));
    DRE_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    DRE_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_slab = (double)ms / reps;
    return DRE_OK;
}

extern "C" int dre_env_set_host_chunks(dre_env *h, int32_t chunks)
{
    if (!h || chunks < 1 || chunks > dre_env::MAX_CHUNKS) return DRE_ERR_INVALID;
    h->host_chunks = chunks;
    return DRE_OK;
}

extern "C" int dre_env_invalidate_orca_cache(dre_env *h, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;
    DRE_CUDA_TRY(cudaMemsetAsync(h->D.cache_valid, 0, (size_t)h->D.E, (cudaStream_t)stream));
    return DRE_OK;
}

extern "C" int dre_env_cvmm_filter(dre_env *h, const double *actions, int32_t lookahead_steps, const double *backup_actions,
                                   int32_t num_backup, double *safe_actions, int32_t *info4, uint8_t *candidate_safe, void *stream)
{
    if (!h || !actions || !backup_actions || !safe_actions) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;   // reads the state: a later step_host must run after it
    return dre_env_launch_cvmm(h->D, actions, lookahead_steps, backup_actions, num_backup, safe_actions, info4, candidate_safe,
                               (cudaStream_t)stream);
}

extern "C" int dre_env_cvmm_check(dre_env *h, const double *actions, int32_t lookahead_steps, uint8_t *safe, double *dist_to_path, void *stream)
{
    if (!h || !actions || !safe) return DRE_ERR_INVALID;
    if (!h->was_reset || !h->D.paths) return DRE_ERR_STATE;
    h->last_stream = (cudaStream_t)stream; h->last_stream_set = true;
    return dre_env_launch_cvmm_check(h->D, actions, lookahead_steps, safe, dist_to_path, (cudaStream_t)stream);
}

extern "C" int dre_env_set_seed(dre_env *h, uint64_t seed)
{
    if (!h) return DRE_ERR_INVALID;
    h->cfg.seed = seed;
    h->D.seed = seed;
    return DRE_OK;
}

extern "C" int dre_env_stats(dre_env *h, int64_t *counters_dev16, double *sums_dev4, int32_t reset, void *stream)
{
    if (!h) return DRE_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s; h->last_stream_set = true;
    if (counters_dev16) DRE_CUDA_TRY(cudaMemcpyAsync(counters_dev16, h->D.stat_counters, 16 * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    if (sums_dev4) DRE_CUDA_TRY(cudaMemcpyAsync(sums_dev4, h->D.stat_sums, 4 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (reset) {
        DRE_CUDA_TRY(cudaMemsetAsync(h->D.stat_counters, 0, 16 * sizeof(int64_t), s));
        DRE_CUDA_TRY(cudaMemsetAsync(h->D.stat_sums, 0, 4 * sizeof(double), s));
    }
    return DRE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Episode statistics across GPUs: the ONE collective of the engine: an all-reduce(sum) of 16 counters + 4
// sums over NCCL / NVLink per reporting interval. libnccl is not linked: the entry points are resolved at run time from
// the libnccl.so.2 the process has loaded (torch's) or can load; a process that never reduces never needs it.
// ---------------------------------------------------------------------------------------------------------------

extern "C" int dre_stats_comm_id(char *id128)
{
    if (!id128) return DRE_ERR_INVALID;
    const NcclApi &A = nccl_api();
    if (!A.ok) { dre_set_error_text("libnccl.so.2 could not be loaded"); return DRE_ERR_CUDA; }
    NcclApi::UniqueId id;
    const int rc = A.GetUniqueId(&id);
    if (rc) return nccl_fail(rc, __LINE__);
    memcpy(id128, id.internal, 128);
    return DRE_OK;
}

extern "C" int dre_env_stats_comm_init(dre_env *h, const char *id128, int32_t rank, int32_t world)
{
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return DRE_ERR_INVALID;
    const NcclApi &A = nccl_api();
    if (!A.ok) { dre_set_error_text("libnccl.so.2 could not be loaded"); return DRE_ERR_CUDA; }
    if (h->stats_comm) { A.CommDestroy(h->stats_comm); h->stats_comm = nullptr; }
    NcclApi::UniqueId id;
    memcpy(id.internal, id128, 128);
    const int rc = A.CommInitRank(&h->stats_comm, world, id, rank);
    if (rc) return nccl_fail(rc, __LINE__);
    return DRE_OK;
}

extern "C" int dre_stats_reduce(dre_env *h, void *nccl_comm, int64_t *counters_dev16, double *sums_dev4, int32_t reset, void *stream)
{
    if (!h || !counters_dev16 || !sums_dev4) return DRE_ERR_INVALID;
    void *comm = nccl_comm ? nccl_comm : h->stats_comm;
    if (!comm) return DRE_ERR_STATE;
    const NcclApi &A = nccl_api();
    if (!A.ok) { dre_set_error_text("libnccl.so.2 could not be loaded"); return DRE_ERR_CUDA; }
    int rc = dre_env_stats(h, counters_dev16, sums_dev4, reset, stream);
    if (rc) return rc;
    rc = A.AllReduce(counters_dev16, counters_dev16, 16, kNcclInt64, kNcclSum, comm, (cudaStream_t)stream);
    if (rc) return nccl_fail(rc, __LINE__);
    rc = A.AllReduce(sums_dev4, sums_dev4, 4, kNcclFloat64, kNcclSum, comm, (cudaStream_t)stream);
    if (rc) return nccl_fail(rc, __LINE__);
    return DRE_OK;
}
