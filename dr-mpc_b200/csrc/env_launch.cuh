// env_launch.cuh -- the env's device-side view (all pointers into HBM) shared by env_kernels.cu and env_capi.cu.
#pragma once
#include "launch.cuh"

struct EnvOutDev {
    double *pt_state, *actions_used, *mpc_actions, *robot_node, *past_robot_vel, *percent_episode, *spatial_edges, *human_valid,
        *detected_humans, *rewards, *info;
    uint32_t *done_bits;
    dre_mpc_status *mpc_status;
    uint8_t *done_flags;
};

struct EnvDev {
    int E, Nh, LB;
    // pipelined host path: the last CTA of every chunk of `chunk_envs` envs stores `step_seq` to chunk_flags[chunk]
    // (mapped host memory); chunk_count = per-chunk CTA counters in device memory. nullptr = no signalling.
    int *chunk_flags, *chunk_count;
    int chunk_envs, step_seq;
    dre_orca_params orca;
    double dt, time_limit, human_radius, robot_radius, circle_radius, circle_radius_humans;
    int end_goal_changing;
    double collision_penalty, safety_dist, safety_penalty, dist_radius, dist_factor_v, dist_factor_th, act_min_vel, act_penalty;
    double pt_overall_scaling, goal_radius, goal_angle, goal_reward, path_adv_scaling, dev_xy, dev_th, dev_scale;
    double corridor_penalty, corridor_max_dev, eop_penalty, safety_corridor_penalty, safety_corridor_max_dev;
    int lookahead, lookbehind, auto_path_switch, soft_reset, orca_dedup, lookahead_prune;
    int sensor_restriction;
    double sensor_range;
    int num_static;
    double static_pos[8];
    uint64_t seed;
    int64_t env_id_offset;
    // ---- state (structure of arrays, env-major) ----
    double2 *hpos;      // [E][Nh]
    float2 *hvel;       // [E][Nh]
    double2 *hgoal;     // [E][Nh]
    uint8_t *hstatic;   // [E][Nh]
    double2 *hhist;     // [E][Nh][LB+1]
    int32_t *hvalid;    // [E][Nh]
    double *rpose;      // [E][4]
    double *rprev;      // [E][4]
    double *rhist;      // [E][LB+1][4]
    double *rpastv;     // [E][LB][2]
    int32_t *rvalid;    // [E]
    double *gtime;      // [E]
    int32_t *path_idx;  // [E]
    uint8_t *loc_to_start;  // [E]
    int64_t *step_count;    // [E]
    int32_t *reset_epoch;   // [E]
    double *mpc_interp;     // [E]
    uint8_t *mpc_iteration0; // [E] (aliases the MPC object's flags)
    int32_t *sr_state;      // [E][4] soft-reset state machine
    float2 *hvel_next;      // [E][Nh] orca_dedup: look-ahead velocities of the last step (= next step's movement pass)
    uint8_t *goal_changed;  // [E][Nh] orca_dedup: goal resampled since hvel_next was computed
    uint8_t *cache_valid;   // [E]     orca_dedup: hvel_next belongs to the current state
    uint8_t *lp_class;      // [E][Nh] 1 = the human's latest ORCA solve needed linearProgram3 (queue order of the next pass)
    // ---- paths ----
    const double *paths;
    const int32_t *lens;
    int num_paths, path_stride;
    // ---- per-step input / output ----
    const double *actions;  // [E][2]
    const uint8_t *env_mask;            // [E] or nullptr: masked reset / observe touch only the envs whose byte is non-zero
    const uint32_t *prev_done_bits;     // [E] done word of the previous step (the output slab is double-buffered)
    EnvOutDev out;
    unsigned long long *stat_counters;  // [16]
    double *stat_sums;                  // [4]
};

int dre_env_launch_ha(const EnvDev &D, int mode, int write_obs, cudaStream_t s);
int dre_env_ha_envs_per_cta(const EnvDev &D);
int dre_env_launch_pt(const EnvDev &D, int mode, cudaStream_t s);
int dre_env_launch_mpc_stats(const EnvDev &D, cudaStream_t s);
int dre_env_launch_cvmm(const EnvDev &D, const double *actions, int num_steps, const double *backup, int A, double *safe_out,
                        int32_t *info, uint8_t *cand_safe, cudaStream_t s);
int dre_env_launch_cvt32(const double *src, float *dst, size_t n, cudaStream_t s);
int dre_env_launch_cvmm_check(const EnvDev &D, const double *actions, int num_steps, uint8_t *safe_out, double *dist_out, cudaStream_t s);
int dre_env_launch_reset(const EnvDev &D, const double *hpos_init, const double *hgoal_init, cudaStream_t s);
int dre_env_launch_switch_path(const EnvDev &D, cudaStream_t s);
int dre_env_launch_sr_decide(const EnvDev &D, int commit, uint8_t *scripted, double *actions, uint8_t *stuck, cudaStream_t s);

// env_capi.cu accessors (replay_capi.cu)
struct dre_env;
const dre_env_config &dre_env_config_of(const dre_env *h);
int dre_env_mpc_actions_width(const dre_env *h);   // doubles per env in dre_step_out.mpc_actions
void dre_env_note_stream(dre_env *h, cudaStream_t s); // work reading the env's state / slabs was queued on s: the next step_host runs after it

// capi.cu accessors
struct dre_mpc;
const dre_mpc_params &dre_mpc_params_of(const dre_mpc *h);
const MpcDeviceState &dre_mpc_state_of(const dre_mpc *h);
