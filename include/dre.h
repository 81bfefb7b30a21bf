/*
 * include/dre.h -- C ABI of libdre.so, the B200-native batched step engine for DR-MPC's hot path
 * (ORCA human step + path-tracking MPC + the env step that wraps them).
 *
 * The reference has no FFI or plugin registry for this path (it is duck-typed Python); each entry
 * point below cites the reference interface it replaces. Conventions: plain pointers and sizes, no
 * exceptions across the ABI (0 = DRE_OK, negative = error, dre_strerror() for text); "dev" pointers
 * are CUDA device pointers owned by the caller (e.g. torch tensors), "host" pointers are host
 * memory (pinned for best speed); `stream` is a cudaStream_t passed as void* (NULL = default
 * stream). A handle is bound to one GPU and is not thread-safe: one handle per GPU / per process.
 * Solver failures are reported through per-env status words, never by aborting
 * (reference mpc.py:188-190 swallows solver exceptions the same way).
 */
#ifndef DRE_H
#define DRE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRE_OK 0
#define DRE_ERR_INVALID (-1)   /* bad argument / unsupported size                     */
#define DRE_ERR_CUDA (-2)      /* a CUDA runtime call failed (dre_last_cuda_error())  */
#define DRE_ERR_NOMEM (-3)
#define DRE_ERR_STATE (-4)     /* call order violated (e.g. step before reset)        */

#define DRE_MAX_HUMANS 64      /* per env, excluding the robot                         */
#define DRE_MAX_HORIZON 40     /* MPC horizon K                                        */
#define DRE_PT_STATE_DIM 100   /* 25 nodes x (x,y,cos,sin), reference config_PT.py:40-45 */

/* ---------------------------------------------------------------------------------------------
 * ORCA parameters. Mirrors what scripts/policy/orca.py:79-89 passes to rvo2.PyRVOSimulator/addAgent
 * (neighbor_dist, max_neighbors, time_horizon, radius = r + 0.01 + safety_space, max_speed) with the
 * defaults of scripts/configs/config_HA.py:100-106.
 * ------------------------------------------------------------------------------------------- */
typedef struct dre_orca_params {
    float neighbor_dist;   /* config_HA.orca.neighbor_dist (10)                          */
    float time_horizon;    /* config_HA.orca.time_horizon (5)                            */
    float time_step;       /* config_general.env.time_step (0.25)                        */
    float radius;          /* ORCA radius of every agent: 0.3 + 0.01 + 0.175             */
    float max_speed_self;  /* human v_max (orca.py:86)                                   */
    int32_t max_neighbors; /* <=0: all others (orca.py:77: len(state.human_states))      */
    int32_t robot_visible; /* config_HA.robot.visible                                    */
    int32_t group_lanes;   /* lanes cooperating on one agent: 2..32 (0 = auto: 8 from 9 agents up). 1 (| 0x400: projected
                            * lines in global scratch) = thread-per-agent variant of the stand-alone pass, a measurement aid */
} dre_orca_params;

/* per-solve diagnostics, used to count "LP-branch divergences" against the oracle */
typedef struct dre_orca_stats {
    int32_t n_lines, lp2_fail, n_lp1, lp1_work;
    uint32_t branch_mask;  /* bit0 cut-off circle, bit1 left leg, bit2 right leg, bit3 collision */
    int32_t n_lp3_proj;
} dre_orca_stats;

/*
 * Batched replacement of CrowdSim.get_human_actions (crowd_sim.py:264-284) ->
 * Human.act -> ORCA.predict (orca.py:64-117): one ORCA solve per human per env.
 *   hpos  dev double[E][Nh][2]  human positions        hvel dev float[E][Nh][2] current velocities
 *   goal  dev double[E][Nh][2]  goals                  rpos dev double[E][2]   robot position
 *   is_static dev uint8[E][Nh] or NULL (static humans return (0,0), crowd_sim.py:270-272)
 *   vnew  dev float[E][Nh][2]   OUT new velocities     stats dev dre_orca_stats[E][Nh] or NULL
 */
int dre_orca_predict(const dre_orca_params *p, int32_t num_envs, int32_t num_humans,
                     const double *hpos, const float *hvel, const double *goal, const double *rpos,
                     const uint8_t *is_static, float *vnew, dre_orca_stats *stats, void *stream);
/* same call on HOST buffers: copies in, launches, copies out, synchronises */
int dre_orca_predict_host(const dre_orca_params *p, int32_t num_envs, int32_t num_humans,
                          const double *hpos, const float *hvel, const double *goal, const double *rpos,
                          const uint8_t *is_static, float *vnew, dre_orca_stats *stats);

/*
 * The literal boundary-B call: ORCA(config).predict(JointState) -> ActionXY (scripts/policy/orca.py:64-117) for N
 * independent joint states of n_others neighbours each, on HOST buffers (synchronous). What crosses is what the
 * reference hands to its private rvo2 simulator, already cast to float like the Cython boundary does:
 *   self8   host float[N][8]           px, py, vx, vy, pref_vx, pref_vy (orca.py:97-100), radius + 0.01 + safety_space
 *                                      (orca.py:85), max speed = v_pref (orca.py:86)
 *   others5 host float[N][n_others][5] px, py, vx, vy, radius + 0.01 + safety_space (orca.py:88), in human_states order
 *   vnew    host float[N][2]  OUT      sim.getAgentVelocity(0) after doStep      stats host dre_orca_stats[N] or NULL
 * p->neighbor_dist / time_horizon / time_step / max_neighbors (<= 0: n_others, orca.py:77) apply; radius, max_speed_self
 * and robot_visible of *p are not used (they come per agent).
 */
int dre_orca_solve_host(const dre_orca_params *p, int32_t num_states, int32_t n_others, const float *self8,
                        const float *others5, float *vnew, dre_orca_stats *stats);

/* ---------------------------------------------------------------------------------------------
 * Path-tracking MPC. Replaces MPC(config_master, K, dt, continuous_task, max_v, max_w) /
 * .reset() / .act({'path','interp_index','pose'}) (environment/path_tracking/utils/mpc.py:21,55,58)
 * and the solver in pysteam_augmented/lev_marq_gauss_newton_custom_solver.py for N problems at once.
 * ------------------------------------------------------------------------------------------- */
typedef struct dre_mpc_params {
    int32_t horizon;          /* K, config_PT.MPC.planning_horizon (4); <= DRE_MAX_HORIZON      */
    double time_step;         /* 0.25                                                           */
    double v_max, w_max;      /* config_general.robot.v_max / w_max (1, 1); v_min = 0, w_min=-w_max */
    double node_separation;   /* config_PT.env.node_separation (0.2)                            */
    int32_t max_iterations;   /* mpc.py:144 (1500)                                              */
    double abs_change_tol;    /* pysteam Solver default absolute_cost_change_threshold (1e-4)   */
    double rel_change_tol;    /* mpc.py:144 relative_cost_change_threshold (1e-4)               */
    int32_t pose_jac_exact;   /* 0: STEAM's -J_l^-1(e) pose-error Jacobian, 1: exact -J_r^-1(e) */
    int32_t max_retries;      /* cap on mpc.py:188-190's pose_error_scaling*=5 loop (default 20) */
} dre_mpc_params;

typedef struct dre_mpc_status {
    int32_t iterations;  /* outer LM iterations of the final attempt            */
    int32_t retries;     /* pose_error_scaling *= 5 retries                     */
    int32_t lm_solves;   /* damped linear solves, all attempts                  */
    int32_t cost_evals;  /* trial cost evaluations, all attempts                */
    int32_t term;        /* 1 zero-gradient, 3 max iterations, 4 zero cost, 5 abs change, 6 rel change, -1 gave up */
    int32_t zero_steps;  /* damped solves whose box-feasibility scale was exactly 0 (the step is void, both trials are rejected) */
    double final_cost;
} dre_mpc_status;

typedef struct dre_mpc dre_mpc;

int dre_mpc_create(const dre_mpc_params *p, int32_t num_problems, dre_mpc **out);
int dre_mpc_destroy(dre_mpc *h);
/* path table: host double[num_paths][3][path_stride] (rows x, y, theta like the reference's 3xL arrays), lens[num_paths] */
int dre_mpc_set_paths(dre_mpc *h, const double *paths_host, int32_t num_paths, int32_t path_stride, const int32_t *lens_host);
/* MPC.reset(): mask dev uint8[N] or NULL (= all) */
int dre_mpc_reset(dre_mpc *h, const uint8_t *mask, void *stream);
/*
 * MPC.act(): path_id dev int32[N] (NULL = path 0), interp_index dev double[N], pose dev double[N][3]
 * -> controls dev double[N][2K] = [v0,w0,...]; status dev dre_mpc_status[N] or NULL.
 */
int dre_mpc_act(dre_mpc *h, const int32_t *path_id, const double *interp_index, const double *pose,
                double *controls, dre_mpc_status *status, void *stream);
int dre_mpc_act_host(dre_mpc *h, const int32_t *path_id, const double *interp_index, const double *pose,
                     double *controls, dre_mpc_status *status);
/* warm-start state for parity tests / checkpoints: host double[N][K][4] (c,s,x,y of the inverse poses),
 * host double[N][K][2], host uint8[N] */
int dre_mpc_get_warm(dre_mpc *h, double *warm_T, double *warm_u, uint8_t *iteration0);
int dre_mpc_set_warm(dre_mpc *h, const double *warm_T, const double *warm_u, const uint8_t *iteration0);

/* ---------------------------------------------------------------------------------------------
 * Full environment. Replaces HAAndPTEnv(config_master, seed_addition) / .reset() / .step()
 * (environment/HA_and_PT/human_avoidance_and_path_tracking_env.py:32,243,305) for E envs at once.
 * ------------------------------------------------------------------------------------------- */
typedef struct dre_env_config {
    int32_t num_envs, num_humans, lookback;      /* lookback: config_general.env.lookback (6)           */
    dre_orca_params orca;
    dre_mpc_params mpc;
    double time_step, time_limit;                /* config_general.env.time_step / time_limit            */
    double human_radius, robot_radius;           /* config_HA.humans.radius / robot.radius (0.3)         */
    double circle_radius, circle_radius_humans;  /* config_HA.sim.circle_radius (4) / _humans (5)        */
    int32_t end_goal_changing;                   /* config_HA.humans.end_goal_changing                   */
    /* HA rewards, config_HA.py:27-36 */
    double collision_penalty, safety_dist, safety_penalty;
    double disturbance_radius, disturbance_factor_v, disturbance_factor_th;
    double actuation_min_vel, actuation_penalty;
    /* PT rewards, config_PT.py:11-22 */
    double pt_overall_scaling, goal_radius, goal_angle, goal_reward;
    double path_advancement_scaling, deviation_xy, deviation_th, deviation_scale;
    double corridor_penalty, corridor_max_dev, end_of_path_penalty, safety_corridor_penalty, safety_corridor_max_dev;
    int32_t lookahead, lookbehind;               /* config_PT.py:38-39 (20, 5): lookahead+lookbehind = 25 */
    int32_t mpc_actions_in_input;                /* config_PT.MPC.actions_in_input (4)                   */
    int32_t auto_path_switch;                    /* 1: advance to the next path on goal / end-of-path inside step()
                                                    (what soft_reset does at HA_and_PT env.py:130-141)   */
    int32_t soft_reset_on_device;                /* 1: HAAndPTEnv.soft_reset (env.py:120-240) runs as a per-env state machine
                                                    inside step(): while an env recovers from a termination its action is
                                                    the scripted one (wait / back up / turn to the path) instead of the
                                                    caller's, and the transition is flagged (done_flags[12]) so the caller
                                                    keeps it out of the replay buffer. Needs auto_path_switch = 1.   */
    int32_t orca_dedup;                          /* 1: reuse the look-ahead ORCA pass of step t (human_avoidance_env.py:245) as
                                                    the movement pass of step t+1 (:129). The two are the same computation
                                                    on the same state unless a human's goal was resampled in between
                                                    (:181-184); those humans are solved again. Bit-identical results, one
                                                    ORCA pass per step instead of two. Writers of the state views must call
                                                    dre_env_invalidate_orca_cache(). Default 0 = two passes like the reference */
    int32_t orca_lookahead_prune;                /* 1: the look-ahead ORCA pass (human_avoidance_env.py:245) solves only the humans whose
                                                    future velocity the disturbance reward reads (:247-252: within its radius of
                                                    the robot, not colliding); the other look-ahead solves are dead work. Bit-
                                                    identical results. Ignored with orca_dedup (which needs the whole pass).
                                                    Default 0 = the whole second pass like the reference */
    int32_t sensor_restriction;                  /* config_HA.robot.sensor_restriction (False): humans farther than sensor_range
                                                    from the robot lose their history (human_avoidance_env.py:102-105,161-164,
                                                    331-334) and the observation lists the visible ones first (:53-72)     */
    double sensor_range;                         /* config_HA.robot.sensor_range (4)                                        */
    int32_t num_static_humans;                   /* the LAST num_static_humans of num_humans are static obstacles (isObstacle,
                                                    crowd_sim.py:171-183, 270-272): placed at static_pos by reset(), zero
                                                    action, no goal resampling, no disturbance term. 0..4                */
    double static_pos[8];                        /* (x, y) per static human; the reference's continuous task hard-codes
                                                    (-4, 0) and (0, 0) when include_static_humans is on (crowd_sim.py:173-183) */
    uint64_t seed;                               /* Philox key for device-side spawning / goal resampling */
    int64_t env_id_offset;                       /* global id of env 0 of this shard (multi-GPU)         */
} dre_env_config;

/* Device pointers into the engine's structure-of-arrays env state (for parity tests, checkpoints and
 * zero-copy consumers). Layouts in DESIGN.md "Data layout in HBM". */
typedef struct dre_env_state_view {
    double *hpos;      /* [E][Nh][2]            */
    float *hvel;       /* [E][Nh][2]            */
    double *hgoal;     /* [E][Nh][2]            */
    uint8_t *hstatic;  /* [E][Nh]               */
    double *hhist;     /* [E][Nh][LB+1][2] global position history, oldest first      */
    int32_t *hvalid;   /* [E][Nh]               */
    double *rpose;     /* [E][4] x,y,theta,pad  */
    double *rprev;     /* [E][4]                */
    double *rhist;     /* [E][LB+1][4]          */
    double *rpastv;    /* [E][LB][2]            */
    int32_t *rvalid;   /* [E]                   */
    double *gtime;     /* [E]                   */
    int32_t *path_idx; /* [E]                   */
    uint8_t *loc_to_start; /* [E]               */
    int64_t *step_count;   /* [E]               */
    int32_t *sr_state;     /* [E][4] soft-reset machine: {mode | out_of_trigger<<8 | extra_backup_left<<16,
                              consecutive_no_v, steps_in_soft_reset, num_backup_steps} (env.py:121-127)   */
} dre_env_state_view;

/* One step's outputs: device pointers into ONE contiguous slab (so that a single D2H copy moves it).
 * Observation layout = the reference's S dict (HA_and_PT env.py:325; human_avoidance_env.py:24-77;
 * path_tracking_env.py:190-204) with the leading batch dim E instead of 1. */
typedef struct dre_step_out {
    double *pt_state;        /* [E][100]            S['PT']['PT_state']            */
    double *actions_used;    /* [E][2]  the (v, w) the step executed: the caller's, or the scripted soft-reset action */
    double *mpc_actions;     /* [E][2*min(K,A)]     S['PT']['MPC_actions']         */
    double *robot_node;      /* [E][2*LB]           S['HA']['robot_node']          */
    double *past_robot_vel;  /* [E][2*LB]           S['HA']['past_robot_velocities'] */
    double *percent_episode; /* [E]                 S['HA']['percent_episode']     */
    double *spatial_edges;   /* [E][Nh][2*(LB+1)]   S['HA']['spatial_edges']       */
    double *human_valid;     /* [E][Nh]             S['HA']['human_valid_history'] */
    double *detected_humans; /* [E]                 S['HA']['detected_human_num']  */
    double *rewards;         /* [E][3]  R_PT, R_DO(HA), R      (env.py:413)        */
    double *info;            /* [E][8]  disturbance_v, disturbance_th, path_advancement, deviation, dmin, interp_index, xy_dev, spare */
    uint32_t *done_bits;     /* [E]     DRE_DONE_* bits        (env.py:391-414)    */
    dre_mpc_status *mpc_status; /* [E]                                              */
    uint8_t *done_flags;     /* [E][16] 0/1: done_PT, done_HA, done, done_for_replay (env.py:414), then eps_done_info
                                (env.py:391-404): success, collision, safety_human_raise, timeout, deviation_end_of_path,
                                corridor_hit, safety_corridor_raise, actuation_termination; [12] this step ran a scripted
                                soft-reset action, [13] soft reset gave up (> 250 steps, env.py:235-238), [14..15] spare */
    void *slab;              /* start of the contiguous slab                        */
    size_t slab_bytes;
} dre_step_out;

typedef struct dre_env dre_env;

int dre_env_create(const dre_env_config *cfg, dre_env **out);
int dre_env_destroy(dre_env *h);
int dre_env_set_paths(dre_env *h, const double *paths_host, int32_t num_paths, int32_t path_stride, const int32_t *lens_host);
int dre_env_state(dre_env *h, dre_env_state_view *view);
/* Device pointers of the output slab holding the LATEST outputs. The slab is double-buffered: step() number t writes slab
 * t & 1, so the observation S a caller got from step t-1 is still intact while it holds S' of step t (the reference's
 * `replay_buffer.insert(S, a, S', ...); S = S'`, online_continuous_task.py:189, works on zero-copy views without a clone).
 * dre_env_outputs_at(which = 0 | 1) returns either slab, dre_env_current_slab() which one is the latest. */
int dre_env_outputs(dre_env *h, dre_step_out *out);
int dre_env_outputs_at(dre_env *h, int32_t which, dre_step_out *out);
int dre_env_current_slab(dre_env *h);
dre_mpc *dre_env_mpc(dre_env *h);                     /* the env's MPC instance (warm-start access) */
/*
 * reset(): robot at (0,-circle_radius, pi) (env.py:286-303), humans spawned by device-side rejection
 * sampling (crowd_sim.py:232-262 with Philox instead of np.random), first MPC solve, LB warm-start
 * steps (human_avoidance_env.py:79-125). If `hpos_init`/`hgoal_init` (dev double[E][Nh][2]) are given
 * they replace the spawning (parity runs upload the oracle's humans).
 */
int dre_env_reset(dre_env *h, const double *hpos_init, const double *hgoal_init, void *stream);
/* reset() of the envs whose byte in env_mask (dev uint8[E]) is non-zero; the others keep their episode (SURVEY 8(b):
 * a reset taking `seeds, env_mask` -- the per-env seed is (config seed, global env id, that env's reset epoch), so
 * no seed array is needed). NULL mask = all envs = dre_env_reset. */
int dre_env_reset_masked(dre_env *h, const uint8_t *env_mask, const double *hpos_init, const double *hgoal_init, void *stream);
/* step(): actions dev double[E][2] = ActionRot(v,w) per env; results in the output slab. */
int dre_env_step(dre_env *h, const double *actions, void *stream);
/* "ORCA-only" step (BASELINE.json configs[3]): CrowdSim.get_human_actions + the humans' Euler step, histories and goal
 * resampling with the robot frozen = one step of HAEnv.warm_start (human_avoidance_env.py:79-125). write_obs = 1 also
 * regenerates the HA observation. */
int dre_env_crowd_step(dre_env *h, int32_t write_obs, void *stream);
/* step() on HOST buffers: actions host double[E][2] -> slab_host (slab_bytes, same layout as the device slab; pinned
 * memory lets the copies overlap the kernels). Synchronous: returns when the whole slab has arrived. Runs on the handle's own
 * streams, ordered after earlier reset() / step() / observe() calls whatever stream those were given. */
int dre_env_step_host(dre_env *h, const double *actions_host, void *slab_host);
/* What dre_env_step_host returns to the host each step (the device slab always holds everything in float64):
 *   0  the whole float64 slab (default; what the reference's loop receives: float64 numpy arrays)
 *   1  observation fields (pt_state, robot_node, past_robot_vel, percent_episode, spatial_edges, human_valid,
 *      detected_humans) as float32 -- each float64 result rounded ONCE on the device -- stored at the field's own byte
 *      offset of slab_host (first half of its region); the control fields below stay float64
 *   2  control fields only: actions_used, mpc_actions, rewards, info, done_bits, mpc_status, done_flags -- for a caller whose
 *      networks consume the observations on the GPU (the reference moves S to cuda:0 anyway, scripts/utils/misc.py:7-21) */
int dre_env_set_host_outputs(dre_env *h, int32_t mode);
/* measurement aid: milliseconds of one plain pinned D2H copy of slab_bytes on this handle's copy stream (mean of reps) */
int dre_env_d2h_probe(dre_env *h, void *slab_host, int32_t reps, double *ms_per_slab);
/* number of env chunks whose observation copies dre_env_step_host overlaps with the crowd half (1..32, default 8) */
int dre_env_set_host_chunks(dre_env *h, int32_t chunks);
/* orca_dedup: forget the cached look-ahead velocities (call after writing human / robot state through dre_env_state) */
int dre_env_invalidate_orca_cache(dre_env *h, void *stream);
/* regenerate observations only (HAEnv.soft_reset / PTEnv.soft_reset, env.py:141-148) */
int dre_env_observe(dre_env *h, int32_t solve_mpc, void *stream);
int dre_env_observe_masked(dre_env *h, const uint8_t *env_mask, int32_t solve_mpc, void *stream);   /* only the masked envs */
/* The path switch of HAAndPTEnv.soft_reset (env.py:130-141) for the masked envs (NULL = all): path_idx <- next path,
 * localize_to_start <- True, MPC.reset(). Follow with dre_env_observe_masked(mask, 1) = PT_env.soft_reset + HA_env.soft_reset
 * (env.py:141-148). For callers that keep auto_path_switch = 0 and drive the reference's soft_reset loop themselves. */
int dre_env_switch_path(dre_env *h, const uint8_t *env_mask, void *stream);
/* One decision of HAAndPTEnv.soft_reset's loop (env.py:128-238) per env, from the latest step's done word and the loop
 * variables in sr_state: scripted dev uint8[E] = 1 where the reference would now issue self.step(action, soft_reset=True)
 * with actions dev double[E][2]; 0 where soft_reset() returns (or was never entered). stuck dev uint8[E] or NULL = the
 * "Stuck in soft reset" exit (:235-238). commit = 1 advances the loop variables (call once per step); 0 only peeks. */
int dre_env_soft_reset_decide(dre_env *h, int32_t commit, uint8_t *scripted, double *actions, uint8_t *stuck, void *stream);
/*
 * CVMM safety filter: HAAndPTEnv.cvmm_diverse_safety_pipeline(action_execute, num_steps, backup_actions)
 * (HA_and_PT env.py:422-513) for every env. actions dev double[E][2] = the policy's (v, w); backup_actions dev
 * double[A][2] (A <= 31; config_general.safety_module.params['action_set']); lookahead_steps = params['lookahead_steps'].
 * -> safe_actions dev double[E][2]; info4 dev int32[E][4] or NULL = {original action safe, number of safe back-up actions,
 * chosen back-up index or -1 for the original, 1 if it is the random fall-back taken when nothing is safe};
 * candidate_safe dev uint8[E][A+1] or NULL = cvmm_safety_check of the original (slot 0) and of every back-up action.
 * The env state is not modified.
 */
int dre_env_cvmm_filter(dre_env *h, const double *actions, int32_t lookahead_steps, const double *backup_actions,
                        int32_t num_backup, double *safe_actions, int32_t *info4, uint8_t *candidate_safe, void *stream);
/* HAAndPTEnv.cvmm_safety_check(action, num_steps) (env.py:475-513) for one action per env: actions dev double[E][2] ->
 * safe dev uint8[E], dist_to_path dev double[E] or NULL (FDE_to_path: 1e10 when unsafe). */
int dre_env_cvmm_check(dre_env *h, const double *actions, int32_t lookahead_steps, uint8_t *safe, double *dist_to_path, void *stream);
/* Philox key of the device-side spawning / goal resampling from the next reset() on (the reference re-seeds np.random
 * with the case number at every hard reset, env.py:246-252). */
int dre_env_set_seed(dre_env *h, uint64_t seed);
/* episode statistics accumulated on device since the last call: int64[16]
 * {steps, success, collision, safety_human_raise, timeout, deviation_end_of_path, corridor_hit,
 *  safety_corridor_raise, actuation_termination, mpc_retries, mpc_gave_up, orca_lp3, 4 spare} and
 * double[4] {sum R, sum R_PT, sum R_HA, spare}; `reset` clears them. Feed to NCCL all-reduce. */
int dre_env_stats(dre_env *h, int64_t *counters_dev16, double *sums_dev4, int32_t reset, void *stream);
/* The ONE collective of the engine: dre_env_stats followed by an in-place ncclAllReduce(sum) of both buffers
 * over `nccl_comm` (an ncclComm_t; NULL = the communicator made by dre_env_stats_comm_init) on `stream`. libnccl.so.2 is
 * resolved at run time from the process (torch's copy), it is not a link dependency. Communicator set-up: rank 0 calls
 * dre_stats_comm_id(id), the host side broadcasts the 128 bytes, every rank calls dre_env_stats_comm_init. */
int dre_stats_comm_id(char *id128);
int dre_env_stats_comm_init(dre_env *h, const char *id128, int32_t rank, int32_t world_size);
int dre_stats_reduce(dre_env *h, void *nccl_comm, int64_t *counters_dev16, double *sums_dev4, int32_t reset, void *stream);

/*
 * Replay ring in device memory, fed straight from the env's double-buffered output slab.
 * Replaces ReplayBuffer (reference scripts/utils/storage.py:6-133) together with the numpy -> torch -> cuda conversion in
 * front of it (scripts/online_continuous_task.py:136,148,187; scripts/utils/misc.py:7-21) for E envs at once.
 * Storage = the reference's (storage.py:21-35): one array [capacity][...] per observation key for 'S' and for 'S_prime',
 * 'rewards' [N][1], 'actions' [N][2], 'done' [N][1]; element type float64 (the reference's) or float32 (each float64 result
 * rounded once). It is a ring: a full buffer overwrites its oldest rows where the reference rolls every tensor by one
 * (storage.py:56-75), so the reference's row j (chronological, 0 = oldest) is slot (head + j) % capacity once full.
 */
#define DRE_REPLAY_F64 0
#define DRE_REPLAY_F32 1
typedef struct dre_replay dre_replay;
typedef struct dre_replay_fields {   /* dev pointers, [rows][...] each; same keys as dre_step_out's observation part */
    void *pt_state;        /* S['PT']['PT_state']   [rows][4*(lookahead+lookbehind)] */
    void *mpc_actions;     /* S['PT']['MPC_actions'] [rows][2*min(K,A)]              */
    void *robot_node;      /* [rows][2*LB]          */
    void *past_robot_vel;  /* [rows][2*LB]          */
    void *percent_episode; /* [rows]                */
    void *spatial_edges;   /* [rows][Nh][2*(LB+1)]  */
    void *human_valid;     /* [rows][Nh]            */
    void *detected_humans; /* [rows]                */
} dre_replay_fields;
typedef struct dre_replay_view {
    dre_replay_fields S, S_prime;
    void *rewards;         /* [rows][1] */
    void *actions;         /* [rows][2] */
    void *done;            /* [rows][1] */
    int64_t rows;
    int32_t dtype;         /* DRE_REPLAY_F64 | DRE_REPLAY_F32 */
} dre_replay_view;
/* ReplayBuffer.__init__ (storage.py:8-35): capacity = buffer_size rows of the env's observation shapes. */
int dre_replay_create(dre_env *env, int64_t capacity, int32_t dtype, dre_replay **out);
int dre_replay_destroy(dre_replay *rb);
/* the ring's arrays (zero-copy access for the learner), the 8 row widths in dre_replay_fields order, the bytes allocated */
int dre_replay_storage(dre_replay *rb, dre_replay_view *view, int32_t *widths8, size_t *bytes);
/* ReplayBuffer.insert(S, A, S_prime, R, done) (storage.py:54-98; call site online_continuous_task.py:189) for every env of
 * the latest step at once: S = the observation of the PREVIOUS slab (what the action was chosen on), S_prime = the latest
 * slab, R = rewards[e][2] ('R'), done = done_for_replay; A = actions dev double[E][2] (the policy's action_model), NULL =
 * the executed actions. keep_mask dev uint8[E] or NULL selects envs; skip_scripted = 1 also leaves out the envs whose step
 * ran a scripted soft-reset action (the reference's soft_reset=True steps never reach insert, :143,189). Asynchronous: the
 * row count stays on the device (dre_replay_count synchronises). */
int dre_replay_insert(dre_replay *rb, const double *actions, const uint8_t *keep_mask, int32_t skip_scripted, void *stream);
/* valid_entries (storage.py:13), the ring head, and the rows stored by the latest insert; any pointer may be NULL */
int dre_replay_count(dre_replay *rb, int64_t *valid, int64_t *head, int64_t *last_inserted, void *stream);
int dre_replay_clear(dre_replay *rb, void *stream);
/* ReplayBuffer.sample_from_buffer (storage.py:101-133): gather the rows with CHRONOLOGICAL indices (dev int64[n], 0 = oldest
 * valid row, like the reference's indices_to_sample) into the caller's arrays `batch` ([n][...] each, element type
 * out_dtype; a float64 ring can hand out float32 batches -- the cast model.py applies -- in the same pass). */
int dre_replay_sample(dre_replay *rb, const int64_t *indices, int64_t n, int32_t out_dtype, const dre_replay_view *batch, void *stream);

/* Measurement support (bench.py): when profiling is on, step() brackets each of its kernels with CUDA events on the
 * caller's stream; dre_env_kernel_ms() synchronises and returns the summed milliseconds per kernel since the last
 * reset, order {ha_step, pt_step, mpc_solve, mpc_stats}, and the number of profiled steps. */
int dre_env_set_profiling(dre_env *h, int32_t on);
int dre_env_kernel_ms(dre_env *h, double *ms4, int64_t *steps, int32_t reset);
/* Accuracy self-test of the MPC kernel's own 1/x (hardware seed + two Newton steps): evaluates it for n device doubles */
int dre_mpc_math_selftest(int32_t n, const double *x_dev, double *rcp_out_dev, void *stream);
/* Same for the kernel's other in-house elementary functions: fn 0 = 1/x, 1 = barrier value (ln x on (0,1], -500 for x <= 0,
 * 500 for x > 1: mpc.py's barrier), 2 / 3 = sin / cos of the small-angle routine (falls back to the library beyond 0.5 rad) */
int dre_mpc_math_selftest_fn(int32_t fn, int32_t n, const double *x_dev, double *out_dev, void *stream);
/* FMA-chain microbenchmark of the CUDA-core peaks on the current device: kind 0 = fp32, 1 = fp64 -> TFLOP/s */
int dre_measure_peak(int32_t kind, double *tflops);

const char *dre_strerror(int code);
const char *dre_last_cuda_error(void);
int dre_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DRE_H */
