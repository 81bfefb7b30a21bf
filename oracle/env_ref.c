/*
 * oracle/env_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat C restatement of one DR-MPC environment step for E independent environments (OpenMP over envs),
 * following the reference's Python line by line:
 *     environment/HA_and_PT/human_avoidance_and_path_tracking_env.py:311-416   (step_continuous_task: merge)
 *     environment/human_avoidance/human_avoidance_env.py:128-186,189-291,24-77,79-125 (HA step / rewards / obs / warm start)
 *     environment/human_avoidance/subENVs/crowd_sim.py:114-144,232-262,264-284 (goal resampling, spawning, ORCA fan-out)
 *     environment/human_avoidance/utils/agent.py:164-176                       (Euler / unicycle integration)
 *     environment/path_tracking/path_tracking_env.py:80-101,190-204             (PT step / state)
 *     environment/path_tracking/utils/path_tracking_core.py:21-81,94-145        (localisation, 100-d state)
 *     environment/path_tracking/utils/reward_comp.py:20-116                    (PT rewards)
 *     environment/path_tracking/utils/unicycle.py:89-126                       (exact-arc unicycle)
 *     environment/HA_and_PT/human_avoidance_and_path_tracking_env.py:120-240   (soft_reset: the scripted recovery loop)
 *     environment/human_avoidance/human_avoidance_env.py:53-72,102-105,161-164,331-334 (robot sensor restriction)
 *     environment/human_avoidance/subENVs/crowd_sim.py:171-183,270-272         (static humans)
 * ORCA and MPC come from rvo2_ref.c / mpc_ref.c. np.random cannot be matched draw for draw, so goal resampling and
 * spawning use Philox4x32-10 keyed by (seed, env id, step, human, attempt) (SURVEY 8(d) "Seeds / RNG").
 * Pinned against tests/golden/env_rollouts.npz (the reference's own HAAndPTEnv.step run on the shims).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- from rvo2_ref.c / mpc_ref.c ---- */
typedef struct { int32_t n_lines, lp2_fail, n_lp1, lp1_work; uint32_t branch_mask; int32_t n_lp3_proj; } orca_stats;
void orca_ref_solve(int n, const float *self8, const float *others5, float neighborDist, int maxNeighbors, float timeHorizon,
                    float timeStep, float *out2, orca_stats *st_out);
typedef struct {
    int32_t K; double dt, vmax, wmax, node_inc; int32_t max_iterations; double abs_change_tol, rel_change_tol;
    int32_t pose_jac_exact, max_retries;
} mpc_params;
typedef struct { int32_t iterations, retries, lm_solves, cost_evals, term, pad; double final_cost; } mpc_stats;
void mpc_ref_act(const mpc_params *P, const double *path, int Lstride, int L, double interp_idx, const double pose[3],
                 double *warm_T, double *warm_u, int32_t *iteration0, double *out_u, mpc_stats *st_out);

#define PI 3.14159265358979323846

typedef struct {
    int32_t E, Nh, LB;
    /* ORCA */
    float neighbor_dist, time_horizon, orca_radius, max_speed_self; int32_t max_neighbors, robot_visible;
    /* general */
    double dt, time_limit, human_radius, robot_radius, circle_radius, circle_radius_humans; int32_t end_goal_changing;
    /* HA rewards */
    double collision_penalty, safety_dist, safety_penalty, dist_radius, dist_factor_v, dist_factor_th, act_min_vel, act_penalty;
    /* PT rewards */
    double pt_scaling, goal_radius, goal_angle, goal_reward, path_adv_scaling, dev_xy, dev_th, dev_scale, corridor_penalty,
        corridor_max_dev, eop_penalty, safety_corridor_penalty, safety_corridor_max_dev;
    int32_t lookahead, lookbehind, mpc_actions_in_input, auto_path_switch;
    uint64_t seed; int64_t env_id_offset;
    mpc_params mpc;
    /* paths */
    const double *paths; const int32_t *lens; int32_t num_paths, path_stride;
    /* options off by default in the reference */
    int32_t soft_reset;              /* run HAAndPTEnv.soft_reset's scripted loop across step() calls (needs auto_path_switch) */
    int32_t sensor_restriction;      /* config_HA.robot.sensor_restriction */
    int32_t num_static;              /* the last num_static humans are obstacles (crowd_sim.py:171-183) */
    double sensor_range;             /* config_HA.robot.sensor_range */
    double static_pos[8];
} env_cfg;

typedef struct {
    double *hpos, *hgoal, *hhist; float *hvel; uint8_t *hstatic; int32_t *hvalid;
    double *rpose, *rprev, *rhist, *rpastv; int32_t *rvalid; double *gtime; int32_t *path_idx; uint8_t *loc_to_start;
    int64_t *step_count; int32_t *reset_epoch;
    double *warm_T, *warm_u; int32_t *iteration0;
    /* outputs */
    double *pt_state, *mpc_actions, *robot_node, *past_robot_vel, *percent_episode, *spatial_edges, *human_valid,
        *detected_humans, *rewards, *info; uint32_t *done_bits; mpc_stats *mpc_status;
    /* soft reset: loop variables of HAAndPTEnv.soft_reset carried from one step() call to the next, in the engine's
     * packing {phase | out_of_trigger_because_of_backup << 8 | extra back-ups left << 16, consecutive_no_v,
     * steps_in_soft_reset, num_backup_steps}; the action each step executed; {scripted, stuck} flags */
    int32_t *sr_state; double *actions_used; uint8_t *soft_flags;
} env_arrays;

enum { B_COLLISION = 1, B_SAFETY = 2, B_ACT = 4, B_TIMEOUT = 8, B_GOAL = 16, B_EOP = 32, B_CORRIDOR = 64, B_SAFETY_CORRIDOR = 128,
       B_HA = 1 << 16, B_PT = 1 << 17, B_ANY = 1 << 18, B_REPLAY = 1 << 19 };

/* ---- Philox4x32-10, identical stream definition to the engine's (dre_common.cuh) ---- */
static void philox(uint64_t seed, uint64_t lo, uint64_t hi, uint32_t out[4])
{
    uint32_t c[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}
static double u01(uint32_t hi, uint32_t lo) { return (double)(((((uint64_t)hi) << 32) | lo) >> 11) * (1.0 / 9007199254740992.0); }
/* key = seed with the reset epoch added to its high word; counter = {env id (64 bit), step count (32 bit),
 * draw | stream << 16 | kind << 24 | sub-step << 26}: the engine's stream definition (env_device.cuh: env_uniform2) */
enum { RNG_STEP = 0, RNG_RESET = 1, RNG_WARM = 2 };
static void uniform2(uint64_t seed, uint32_t epoch, uint64_t env, uint32_t kind, uint32_t sub, uint64_t step, uint32_t stream,
                     uint32_t draw, double *a, double *b)
{
    uint32_t r[4];
    const uint32_t c3 = (draw & 0xFFFFu) | ((stream & 0xFFu) << 16) | ((kind & 3u) << 24) | ((sub & 0x3Fu) << 26);
    philox(seed + ((uint64_t)epoch << 32), env, (uint64_t)(uint32_t)step | ((uint64_t)c3 << 32), r);
    *a = u01(r[0], r[1]); *b = u01(r[2], r[3]);
}

static double np_mod(double a, double b) { double r = fmod(a, b); if (r != 0.0 && r < 0.0) r += b; return r; }
static double wrap_pi(double t) { return np_mod(t + PI, 2.0 * PI) - PI; }

/* unicycle.py:89-126 */
static void unicycle(const double x[3], double v, double w, double dt, double y[3])
{
    double dxl, dyl, dth;
    if (w == 0.0) { dxl = v * dt; dyl = 0.0; dth = 0.0; }
    else { dxl = (v / w) * sin(w * dt); dyl = (v / w) * (1.0 - cos(w * dt)); dth = w * dt; }
    const double c = cos(x[2]), s = sin(x[2]);
    y[0] = x[0] + (c * dxl - s * dyl);
    y[1] = x[1] + (s * dxl + c * dyl);
    y[2] = wrap_pi(x[2] + dth);
}

typedef struct { double pose[3]; double index; } loc_t;

/* path_tracking_core.py:21-81, strict=False */
static loc_t localize(const double *px, const double *py, const double *pth, int L, double qx, double qy, int to_start)
{
    int i0 = 0, i1 = 1; double d0 = INFINITY, d1 = INFINITY;
    for (int i = 0; i < L; ++i) {
        const double dx = px[i] - qx, dy = py[i] - qy, d = dx * dx + dy * dy;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; } else if (d < d1) { d1 = d; i1 = i; }
    }
    loc_t R;
#define RET_NODE(n) do { R.pose[0] = px[n]; R.pose[1] = py[n]; R.pose[2] = pth[n]; R.index = (double)(n); return R; } while (0)
    if (to_start) { if (i0 > 5 * L / 8 || i1 > 5 * L / 8) RET_NODE(0); }
    else { if (i0 < L / 6 || i1 < L / 6) RET_NODE(L - 1); }
    const int ad = abs(i0 - i1);
    if (ad != 1 && ad != L - 1) RET_NODE(i0);
    if (i0 == 0) { if ((qx - px[0]) * (px[1] - px[0]) + (qy - py[0]) * (py[1] - py[0]) < 0.0) RET_NODE(0); }
    else if (i0 == L - 1) { if ((qx - px[L - 1]) * (px[L - 2] - px[L - 1]) + (qy - py[L - 1]) * (py[L - 2] - py[L - 1]) < 0.0) RET_NODE(L - 1); }
    double bx = px[i1] - px[i0], by = py[i1] - py[i0];
    const double dx = qx - px[i0], dy = qy - py[i0];
    double proj = (dx * bx + dy * by) / (bx * bx + by * by);
    int other = i1;
    if (proj < 0.0) {
        other = ((i0 - (i1 - i0)) % L + L) % L;
        bx = px[other] - px[i0]; by = py[other] - py[i0];
        proj = (dx * bx + dy * by) / (bx * bx + by * by);
    }
    R.index = (double)i0 + proj * (double)(other - i0);
    R.pose[0] = px[i0] + proj * (px[other] - px[i0]);
    R.pose[1] = py[i0] + proj * (py[other] - py[i0]);
    R.pose[2] = pth[i0] + proj * (pth[other] - pth[i0]);
    return R;
#undef RET_NODE
}

/* crowd_sim.py:264-284 + orca.py:64-117 for one env; pos/vel/goal arrays of Nh humans */
static void human_actions(const env_cfg *c, const double *P, const float *V, const double *G, const uint8_t *stat,
                          const double rp[2], float *vout, float *others, int *lp3)
{
    const int Nh = c->Nh;
    for (int i = 0; i < Nh; ++i) {
        if (stat && stat[i]) { vout[2 * i] = 0.f; vout[2 * i + 1] = 0.f; continue; }
        int n = 0;
        for (int j = 0; j < Nh; ++j) {
            if (j == i) continue;
            others[5 * n] = (float)P[2 * j]; others[5 * n + 1] = (float)P[2 * j + 1];
            others[5 * n + 2] = V[2 * j]; others[5 * n + 3] = V[2 * j + 1]; others[5 * n + 4] = c->orca_radius; ++n;
        }
        if (c->robot_visible) {
            others[5 * n] = (float)rp[0]; others[5 * n + 1] = (float)rp[1]; others[5 * n + 2] = 0.f; others[5 * n + 3] = 0.f;
            others[5 * n + 4] = c->orca_radius; ++n;
        }
        const double dx = G[2 * i] - P[2 * i], dy = G[2 * i + 1] - P[2 * i + 1];
        const double speed = sqrt(dx * dx + dy * dy);
        double pvx = dx, pvy = dy;
        if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
        const float self8[8] = {(float)P[2 * i], (float)P[2 * i + 1], V[2 * i], V[2 * i + 1], (float)pvx, (float)pvy, c->orca_radius, c->max_speed_self};
        orca_stats st;
        orca_ref_solve(n, self8, others, c->neighbor_dist, c->max_neighbors > 0 ? c->max_neighbors : n, c->time_horizon,
                       (float)c->dt, vout + 2 * i, &st);
        if (lp3 && st.lp2_fail < st.n_lines) (*lp3)++;
    }
}

/* crowd_sim.py:114-144 with Philox */
static void resample_goals(const env_cfg *c, const env_arrays *A, int e, uint32_t kind, uint32_t sub)
{
    const uint64_t stp = (uint64_t)A->step_count[e];
    const uint32_t epoch = (uint32_t)A->reset_epoch[e];
    const int Nh = c->Nh;
    double *P = A->hpos + (size_t)e * Nh * 2, *G = A->hgoal + (size_t)e * Nh * 2;
    const uint8_t *stat = A->hstatic + (size_t)e * Nh;
    const double rgx = 0.0, rgy = c->circle_radius, rpx = A->rpose[4 * (size_t)e], rpy = A->rpose[4 * (size_t)e + 1];
    const uint64_t gid = (uint64_t)(c->env_id_offset + e);
    for (int i = 0; i < Nh; ++i) {
        if (stat[i]) continue;
        const double ddx = G[2 * i] - P[2 * i], ddy = G[2 * i + 1] - P[2 * i + 1];
        if (!(sqrt(ddx * ddx + ddy * ddy) < c->human_radius)) continue;
        const double md_r = c->human_radius + c->robot_radius + 0.2, md_h = c->human_radius + c->human_radius + 0.2;
        double gx = G[2 * i], gy = G[2 * i + 1];
        for (uint32_t attempt = 0; attempt < 4096; ++attempt) {
            double u0, u1, u2, u3;
            uniform2(c->seed, epoch, gid, kind, sub, stp, (uint32_t)i, 2 * attempt, &u0, &u1);
            uniform2(c->seed, epoch, gid, kind, sub, stp, (uint32_t)i, 2 * attempt + 1, &u2, &u3);
            const double angle = u0 * PI * 2.0;
            gx = c->circle_radius_humans * cos(angle) + (u1 - 0.5);
            gy = c->circle_radius_humans * sin(angle) + (u2 - 0.5);
            int collide = 0;
            { const double a = gx - rgx, b = gy - rgy, cc = gx - rpx, d = gy - rpy;
              if (sqrt(a * a + b * b) < md_r || sqrt(cc * cc + d * d) < md_r) collide = 1; }
            for (int j = 0; j < Nh && !collide; ++j) {
                if (j == i) continue;
                const double a = gx - G[2 * j], b = gy - G[2 * j + 1], cc = gx - P[2 * j], d = gy - P[2 * j + 1];
                if (sqrt(a * a + b * b) < md_h || sqrt(cc * cc + d * d) < md_h) collide = 1;
            }
            if (!collide) break;
        }
        G[2 * i] = gx; G[2 * i + 1] = gy;
    }
}

/* human_avoidance_env.py:24-77 */
static void write_ha_obs(const env_cfg *c, const env_arrays *A, int e)
{
    const int Nh = c->Nh, H = c->LB + 1, LB = c->LB;
    const double *rp = A->rpose + 4 * (size_t)e;
    const double ct = cos(rp[2]), st = sin(rp[2]);
    for (int i = 0; i < Nh; ++i)
        for (int s = 0; s < H; ++s) {
            const double *q = A->hhist + (((size_t)e * Nh + i) * H + s) * 2;
            const double dx = q[0] - rp[0], dy = q[1] - rp[1];
            double *o = A->spatial_edges + ((size_t)e * Nh + i) * 2 * H + 2 * s;
            o[0] = ct * dx + st * dy; o[1] = -st * dx + ct * dy;
        }
    for (int i = 0; i < Nh; ++i) A->human_valid[(size_t)e * Nh + i] = (double)A->hvalid[(size_t)e * Nh + i];
    int detected = Nh;
    if (c->sensor_restriction) {                                                   /* :53-72 */
        /* ob['spatial_edges'][:nv] = ob['spatial_edges'][visible]: the visible humans' rows move to the front, the
         * rows behind them keep what they held; valid counts likewise, zero beyond nv */
        const int *hv = A->hvalid + (size_t)e * Nh;
        double *se = A->spatial_edges + (size_t)e * Nh * 2 * H, *hvo = A->human_valid + (size_t)e * Nh;
        int nv = 0;
        for (int i = 0; i < Nh; ++i)
            if (hv[i] != 0) {
                if (nv != i) memcpy(se + (size_t)nv * 2 * H, se + (size_t)i * 2 * H, sizeof(double) * 2 * (size_t)H);   /* nv <= i: sources are never overwritten early */
                hvo[nv] = (double)hv[i];
                ++nv;
            }
        if (nv > 0) { for (int i = nv; i < Nh; ++i) hvo[i] = 0.0; detected = nv; }
        else {                                                                     /* a fake human far away */
            for (int k = 0; k < 2 * H; ++k) se[k] = -10.0;
            for (int i = 0; i < Nh; ++i) hvo[i] = 1.0;
            detected = 1;
        }
    }
    for (int s = 0; s < LB; ++s) {
        const double *rh = A->rhist + ((size_t)e * H + s) * 4;
        const double dx = rh[0] - rp[0], dy = rh[1] - rp[1];
        A->robot_node[(size_t)e * 2 * LB + 2 * s] = ct * dx + st * dy;
        A->robot_node[(size_t)e * 2 * LB + 2 * s + 1] = -st * dx + ct * dy;
        A->past_robot_vel[(size_t)e * 2 * LB + 2 * s] = A->rpastv[((size_t)e * LB + s) * 2];
        A->past_robot_vel[(size_t)e * 2 * LB + 2 * s + 1] = A->rpastv[((size_t)e * LB + s) * 2 + 1];
    }
    A->percent_episode[e] = A->gtime[e] / c->time_limit * 10.0 - 5.0;
    A->detected_humans[e] = (double)detected;
}

/* HA step for env e. mode 0 step, 1 warm-start step */
static void ha_step(const env_cfg *c, const env_arrays *A, int e, const double *action, int mode, int write_obs, float *scratch, int *lp3)
{
    const int Nh = c->Nh, H = c->LB + 1, LB = c->LB;
    double *P = A->hpos + (size_t)e * Nh * 2, *G = A->hgoal + (size_t)e * Nh * 2;
    float *V = A->hvel + (size_t)e * Nh * 2;
    const uint8_t *stat = A->hstatic + (size_t)e * Nh;
    double *rp = A->rpose + 4 * (size_t)e;
    float *vnew = scratch, *vnext = scratch + 2 * Nh, *others = scratch + 4 * Nh;
    human_actions(c, P, V, G, stat, rp, vnew, others, lp3);                         /* human_avoidance_env.py:129 */
    double av = 0.0, aw = 0.0;
    if (mode == 0) {                                                               /* :132 robot.step */
        av = action[0]; aw = action[1];
        double y[3];
        memcpy(A->rprev + 4 * (size_t)e, rp, 3 * sizeof(double));
        unicycle(rp, av, aw, c->dt, y);
        rp[0] = y[0]; rp[1] = y[1]; rp[2] = y[2];
        A->gtime[e] += c->dt;
        A->step_count[e] += 1;
    }
    {                                                                              /* :133-141 / :114-122 */
        double *rh = A->rhist + (size_t)e * H * 4, *pv = A->rpastv + (size_t)e * LB * 2;
        const int valid = A->rvalid[e];
        if (valid == H) {
            memmove(rh, rh + 4, sizeof(double) * 4 * (size_t)(H - 1));
            rh[(H - 1) * 4] = rp[0]; rh[(H - 1) * 4 + 1] = rp[1]; rh[(H - 1) * 4 + 2] = rp[2];
            memmove(pv, pv + 2, sizeof(double) * 2 * (size_t)(LB - 1));
            pv[(LB - 1) * 2] = av; pv[(LB - 1) * 2 + 1] = aw;
        } else {
            rh[valid * 4] = rp[0]; rh[valid * 4 + 1] = rp[1]; rh[valid * 4 + 2] = rp[2];
            if (valid - 1 >= 0 && valid - 1 < LB) { pv[(valid - 1) * 2] = av; pv[(valid - 1) * 2 + 1] = aw; }
            A->rvalid[e] = valid + 1 < H ? valid + 1 : H;
        }
    }
    for (int i = 0; i < Nh; ++i) {                                                 /* :144-167 */
        P[2 * i] = P[2 * i] + (double)vnew[2 * i] * c->dt;
        P[2 * i + 1] = P[2 * i + 1] + (double)vnew[2 * i + 1] * c->dt;
        V[2 * i] = vnew[2 * i]; V[2 * i + 1] = vnew[2 * i + 1];
        double *hh = A->hhist + ((size_t)e * Nh + i) * H * 2;
        int *hv = A->hvalid + (size_t)e * Nh + i;
        if (*hv == H) { memmove(hh, hh + 2, sizeof(double) * 2 * (size_t)(H - 1)); hh[(H - 1) * 2] = P[2 * i]; hh[(H - 1) * 2 + 1] = P[2 * i + 1]; }
        else { hh[*hv * 2] = P[2 * i]; hh[*hv * 2 + 1] = P[2 * i + 1]; *hv = *hv + 1 < H ? *hv + 1 : H; }
        if (c->sensor_restriction) {                                               /* :161-164 / :102-105 */
            const double ct = cos(rp[2]), st = sin(rp[2]), dx = P[2 * i] - rp[0], dy = P[2 * i + 1] - rp[1];
            const double lx = ct * dx + st * dy, ly = -st * dx + ct * dy;
            if (sqrt(lx * lx + ly * ly) > c->sensor_range) *hv = 0;
        }
    }
    if (mode == 0) {                                                               /* calc_reward :189-291 */
        double reward = 0.0, dmin = INFINITY; uint32_t bits = 0; int collision = 0;
        for (int i = 0; i < Nh; ++i) {
            const double dx = P[2 * i] - rp[0], dy = P[2 * i + 1] - rp[1];
            const double closest = sqrt(dx * dx + dy * dy) - c->human_radius - c->robot_radius;
            if (closest <= 0.0) { collision = 1; dmin = 0.0; break; }
            else if (closest < dmin) dmin = closest;
        }
        if (collision) { reward += c->collision_penalty; bits |= B_COLLISION | B_HA; }
        if (dmin < c->safety_dist) { reward += c->safety_penalty; bits |= B_SAFETY | B_HA; }
        human_actions(c, P, V, G, stat, rp, vnext, others, NULL);                   /* :245 */
        double pen_v = 0.0, pen_th = 0.0;
        for (int i = 0; i < Nh; ++i) {
            if (stat[i]) continue;
            const double dx = rp[0] - P[2 * i], dy = rp[1] - P[2 * i + 1];
            const double dist = sqrt(dx * dx + dy * dy) - c->human_radius - c->robot_radius;
            if (dist > c->dist_radius || dist <= 0.0) continue;
            const double cvx = V[2 * i], cvy = V[2 * i + 1], fvx = vnext[2 * i], fvy = vnext[2 * i + 1];
            const double curr_v = sqrt(cvx * cvx + cvy * cvy), fut_v = sqrt(fvx * fvx + fvy * fvy);
            const double curr_th = atan2(cvy, cvx), fut_th = atan2(fvy, fvx);
            const double dv = fut_v - curr_v;
            const double pv_i = c->dist_factor_v * (dv < 0.0 ? dv : 0.0);
            const double pth_i = -c->dist_factor_th * fabs(wrap_pi(fut_th - curr_th));
            const double scaling = 1.0 - dist / c->dist_radius;
            pen_v += pv_i * scaling; pen_th += pth_i * scaling;
        }
        reward += (pen_v + pen_th);
        double total_v = 0.0;
        for (int s = 0; s < LB; ++s) total_v += fabs(A->rpastv[((size_t)e * LB + s) * 2]);
        if (total_v < c->act_min_vel) { reward += c->act_penalty; bits |= B_ACT | B_HA; }
        A->rewards[3 * (size_t)e + 1] = reward;
        A->info[8 * (size_t)e] = pen_v; A->info[8 * (size_t)e + 1] = pen_th; A->info[8 * (size_t)e + 4] = dmin;
        A->done_bits[e] = bits;
    }
    if (write_obs) write_ha_obs(c, A, e);
    if (c->end_goal_changing) resample_goals(c, A, e, mode == 1 ? RNG_WARM : RNG_STEP, mode == 1 ? (uint32_t)A->rvalid[e] : 0u);
}

/* path_tracking_core.py:94-145 (continuous task) */
static void pt_state(const env_cfg *c, const double *px, const double *py, const double *pth, int L, const double pose[3],
                     double interp_index, double *out)
{
    const double ct = cos(pose[2]), st = sin(pose[2]);
    const int n = c->lookahead + c->lookbehind, next = (int)ceil(interp_index), start = next - c->lookbehind;
    for (int j = 0; j < n; ++j) {
        const int idx = ((start + j) % L + L) % L;
        const double dx = px[idx] - pose[0], dy = py[idx] - pose[1], th = pth[idx] - pose[2];
        out[j] = ct * dx + st * dy; out[n + j] = -st * dx + ct * dy; out[2 * n + j] = cos(th); out[3 * n + j] = sin(th);
    }
}

/* PT step for env e. mode 0 step, 1 observe */
static void pt_step(const env_cfg *c, const env_arrays *A, int e, const double *action, int mode)
{
    int pid = A->path_idx[e];
    const double *px = c->paths + (size_t)pid * 3 * c->path_stride, *py = px + c->path_stride, *pth = py + c->path_stride;
    int L = c->lens[pid];
    const double *cur_pose = A->rpose + 4 * (size_t)e;
    int lts = A->loc_to_start[e] != 0;
    loc_t cur;
    if (mode == 0) {
        const double *prev_pose = A->rprev + 4 * (size_t)e;
        const loc_t prv = localize(px, py, pth, L, prev_pose[0], prev_pose[1], lts);
        cur = localize(px, py, pth, L, cur_pose[0], cur_pose[1], lts);
        if (cur.index >= (double)(L / 3) && cur.index < (double)(2 * L / 3)) lts = 0;
        double reward = 0.0; uint32_t bits = 0; const double sc = c->pt_scaling;
        {   /* goal */
            int hit = 0;
            for (int i = 1; i <= 5; ++i) {
                double q[3];
                unicycle(prev_pose, action[0], action[1], c->dt * (double)i / 5.0, q);
                const double dx = q[0] - px[L - 1], dy = q[1] - py[L - 1];
                const double ad = fabs(np_mod(q[2] - pth[L - 1] + PI, 2.0 * PI) - PI);
                if (sqrt(dx * dx + dy * dy) < c->goal_radius && ad < c->goal_angle) hit = 1;
            }
            if (hit) { reward += sc * c->goal_reward; bits |= B_GOAL | B_PT; } else reward += sc * 0.0;
        }
        double adv, devr;
        {   /* path_advancement */
            double pi_ = prv.index, ci_ = cur.index;
            if (fabs(ci_ - pi_) > 10.0) { const double half = (double)(L / 2); pi_ = np_mod(pi_ + half, (double)L); ci_ = np_mod(ci_ + half, (double)L); }
            adv = sc * (c->path_adv_scaling * (ci_ - pi_));
            reward += adv;
        }
        const double ddx = cur_pose[0] - cur.pose[0], ddy = cur_pose[1] - cur.pose[1];
        const double xy_dev = sqrt(ddx * ddx + ddy * ddy);
        {   /* deviation */
            const double th_dev = np_mod(cur_pose[2] - cur.pose[2] + PI, 2.0 * PI) - PI;
            devr = sc * ((-xy_dev * c->dev_xy + -c->dev_th * fabs(th_dev)) * c->dev_scale);
            reward += devr;
        }
        if (L - 1 == (int)prv.index) { reward += sc * c->eop_penalty; bits |= B_EOP | B_PT; } else reward += sc * 0.0;
        if (c->corridor_max_dev < xy_dev) { reward += sc * c->corridor_penalty; bits |= B_CORRIDOR | B_PT; } else reward += sc * 0.0;
        if (c->safety_corridor_max_dev < xy_dev) { reward += sc * c->safety_corridor_penalty; bits |= B_SAFETY_CORRIDOR | B_PT; } else reward += sc * 0.0;
        const double r_ha = A->rewards[3 * (size_t)e + 1];
        uint32_t all = A->done_bits[e] | bits;
        if (all & (B_HA | B_PT)) all |= B_ANY | B_REPLAY;
        A->rewards[3 * (size_t)e] = reward; A->rewards[3 * (size_t)e + 2] = reward + r_ha;
        A->done_bits[e] = all;
        A->info[8 * (size_t)e + 2] = adv; A->info[8 * (size_t)e + 3] = devr; A->info[8 * (size_t)e + 5] = cur.index;
        A->info[8 * (size_t)e + 6] = xy_dev; A->info[8 * (size_t)e + 7] = prv.index;
        if (c->auto_path_switch && (bits & (B_GOAL | B_EOP))) {
            pid = (pid + 1) % c->num_paths;
            A->path_idx[e] = pid; lts = 1; A->iteration0[e] = 1;
            px = c->paths + (size_t)pid * 3 * c->path_stride; py = px + c->path_stride; pth = py + c->path_stride; L = c->lens[pid];
            cur = localize(px, py, pth, L, cur_pose[0], cur_pose[1], lts);
        }
        A->loc_to_start[e] = (uint8_t)lts;
    } else {
        cur = localize(px, py, pth, L, cur_pose[0], cur_pose[1], lts);
    }
    const int npt = 4 * (c->lookahead + c->lookbehind);
    pt_state(c, px, py, pth, L, cur_pose, cur.index, A->pt_state + (size_t)e * npt);
    /* MPC.act (path_tracking_env.py:200-202) */
    const int K = c->mpc.K;
    int nact = 2 * (K < c->mpc_actions_in_input ? K : c->mpc_actions_in_input);
    double u[2 * 64];
    mpc_ref_act(&c->mpc, c->paths + (size_t)pid * 3 * c->path_stride, c->path_stride, L, cur.index, cur_pose,
                A->warm_T + (size_t)e * K * 4, A->warm_u + (size_t)e * K * 2, A->iteration0 + e, u, A->mpc_status + e);
    memcpy(A->mpc_actions + (size_t)e * nact, u, sizeof(double) * (size_t)nact);
}


/* HAAndPTEnv.soft_reset (HA_and_PT env.py:120-240) for one env of a batch that steps in lock-step: the reference's
 * `while done:` loop body is entered once per step() call instead of looping, with its local variables kept in
 * sr_state between calls. Called BEFORE a step with the previous step's done word; returns 1 and the scripted action
 * when the reference would be inside soft_reset() issuing self.step(action, soft_reset=True), 0 when control is with
 * the caller's policy. *stuck: the "Stuck in soft reset" exit (:235-238).
 * phase 0: outside soft_reset(); 1: a `while done` iteration has issued its step; 2: inside the 7-step extra back-up
 * loop (:227-233). The path switch on goal / end of path (:130-181) has already happened inside the step that raised it
 * (auto_path_switch), so those reasons just hand the new state back. */
static int soft_reset_next(const env_cfg *c, const env_arrays *A, int e, double act[2], int *stuck)
{
    int32_t *st = A->sr_state + 4 * (size_t)e;
    int phase = st[0] & 0xff, oot = (st[0] >> 8) & 0xff, extra = (st[0] >> 16) & 0xff, consec = st[1], steps = st[2], nb = st[3];
    const uint32_t w = A->done_bits[e];
    const int done = (w & B_ANY) != 0, done_HA = (w & B_HA) != 0;
    const uint32_t why = w & 0xFFu;
    int scripted = 0;
    *stuck = 0;
    for (;;) {
        if (phase == 2) {                                       /* for i in range(7): step(back-up); if done: break */
            if (!done && extra > 0) { extra -= 1; act[0] = -0.3; act[1] = 0.0; scripted = 1; break; }
            phase = 1;                                          /* the for loop is over: fall to the checks that follow it */
            if (steps > 250) { *stuck = 1; phase = 0; break; }
            if (!done) { phase = 0; break; }
            /* done: next `while` iteration */
        } else if (phase == 1) {                                /* a while-iteration's step has just run */
            if (!done && oot) { phase = 2; extra = 7; continue; }
            if (steps > 250) { *stuck = 1; phase = 0; break; }
            if (!done) { phase = 0; break; }
        } else {                                                /* caller's loop: `if done: S = env.soft_reset(...)` */
            if (!done) break;
            consec = 0; steps = 0; oot = 0; nb = 0; extra = 0;
        }
        /* ---- top of a `while done:` iteration ---- */
        if ((why & (B_GOAL | B_EOP)) != 0) { phase = 0; break; }                       /* :130-181 */
        if (why == B_ACT) { phase = 0; break; }                                         /* :184-185 */
        phase = 1;
        if (done_HA && consec > 10 && nb < 5) { act[0] = -0.3; act[1] = 0.0; nb += 1; oot = 1; }   /* :187-190 */
        else {
            nb = 0; oot = 0;
            const int pid = A->path_idx[e], L = c->lens[pid];
            const double *px = c->paths + (size_t)pid * 3 * c->path_stride, *py = px + c->path_stride, *pth = py + c->path_stride;
            const double *rp = A->rpose + 4 * (size_t)e;
            const loc_t loc = localize(px, py, pth, L, rp[0], rp[1], A->loc_to_start[e] != 0);   /* query_closest_path_node_plus_one */
            const int node = (int)(loc.index + 1.0) % L;
            const double desired = atan2(py[node] - rp[1], px[node] - rp[0]);
            const double th_diff = np_mod(desired - rp[2] + PI, 2.0 * PI) - PI;
            const double turn = th_diff < 0.0 ? -0.1 : 0.1;
            if ((why & (B_CORRIDOR | B_SAFETY_CORRIDOR)) != 0 && (why & B_SAFETY) == 0) {   /* :208-213 */
                if (fabs(th_diff) > 0.3) { act[0] = 0.0; act[1] = turn * 7.0; } else { act[0] = 0.5; act[1] = turn; }
                consec = 0;
            } else { act[0] = 0.0; act[1] = 0.0; consec += 1; }                             /* :215-216 */
        }
        steps += 1;
        scripted = 1;
        break;
    }
    st[0] = phase | (oot << 8) | (extra << 16); st[1] = consec; st[2] = steps; st[3] = nb;
    return scripted;
}

void env_ref_step(const env_cfg *c, const env_arrays *A, const double *actions, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        float *scratch = (float *)malloc(sizeof(float) * (size_t)(4 * c->Nh + 5 * (c->Nh + 1)));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int e = 0; e < c->E; ++e) {
            double act[2] = {actions[2 * (size_t)e], actions[2 * (size_t)e + 1]};
            if (c->soft_reset && A->sr_state) {
                int stuck = 0;
                const int scripted = soft_reset_next(c, A, e, act, &stuck);
                if (A->soft_flags) { A->soft_flags[2 * (size_t)e] = (uint8_t)scripted; A->soft_flags[2 * (size_t)e + 1] = (uint8_t)stuck; }
            }
            if (A->actions_used) { A->actions_used[2 * (size_t)e] = act[0]; A->actions_used[2 * (size_t)e + 1] = act[1]; }
            ha_step(c, A, e, act, 0, 1, scratch, NULL);
            pt_step(c, A, e, act, 0);
        }
        free(scratch);
    }
}

/* "ORCA-only" step (BASELINE config 4): one ORCA pass + integration + histories + goal resampling, robot frozen =
 * one warm-start step of HAEnv.warm_start (human_avoidance_env.py:79-125) */
void env_ref_crowd_step(const env_cfg *c, const env_arrays *A, int nthreads, int64_t *lp3_count)
{
    int64_t total = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel reduction(+ : total)
#endif
    {
        float *scratch = (float *)malloc(sizeof(float) * (size_t)(4 * c->Nh + 5 * (c->Nh + 1)));
        int lp3 = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int e = 0; e < c->E; ++e) ha_step(c, A, e, NULL, 1, 0, scratch, &lp3);
        total += lp3;
        free(scratch);
    }
    if (lp3_count) *lp3_count = total;
}

void env_ref_observe(const env_cfg *c, const env_arrays *A, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int e = 0; e < c->E; ++e) { write_ha_obs(c, A, e); pt_step(c, A, e, NULL, 1); }
}

/* reset_continuous_task (HA_and_PT env.py:263-280) */
void env_ref_reset(const env_cfg *c, const env_arrays *A, const double *hpos_init, const double *hgoal_init, int nthreads)
{
    const int Nh = c->Nh, H = c->LB + 1, LB = c->LB;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        float *scratch = (float *)malloc(sizeof(float) * (size_t)(4 * Nh + 5 * (Nh + 1)));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int e = 0; e < c->E; ++e) {
            const double rx = 0.0, ry = -c->circle_radius, rth = PI;
            double *rp = A->rpose + 4 * (size_t)e, *rv = A->rprev + 4 * (size_t)e;
            rp[0] = rv[0] = rx; rp[1] = rv[1] = ry; rp[2] = rv[2] = rth; rp[3] = rv[3] = 0.0;
            for (int s = 0; s < H; ++s) { double *rh = A->rhist + ((size_t)e * H + s) * 4; rh[0] = rx; rh[1] = ry; rh[2] = rth; rh[3] = 0.0; }
            memset(A->rpastv + (size_t)e * LB * 2, 0, sizeof(double) * 2 * (size_t)LB);
            A->rvalid[e] = 1; A->gtime[e] = 0.0; A->step_count[e] = 0; A->reset_epoch[e] += 1;
            A->path_idx[e] = 0; A->loc_to_start[e] = 1; A->iteration0[e] = 1;
            A->done_bits[e] = 0u;
            if (A->sr_state) memset(A->sr_state + 4 * (size_t)e, 0, 4 * sizeof(int32_t));
            if (A->soft_flags) { A->soft_flags[2 * (size_t)e] = 0; A->soft_flags[2 * (size_t)e + 1] = 0; }
            const uint64_t gid = (uint64_t)(c->env_id_offset + e);
            const uint32_t epoch = (uint32_t)A->reset_epoch[e];
            for (int i = 0; i < Nh; ++i) {
                const size_t g = (size_t)e * Nh + i;
                double px, py, gx, gy;
                const int si = i - (Nh - c->num_static);                           /* >= 0: static human (crowd_sim.py:171-183) */
                if (si >= 0 && !hpos_init) { px = c->static_pos[2 * si]; py = c->static_pos[2 * si + 1]; gx = 0.0; gy = 0.0; }
                else if (hpos_init) { px = hpos_init[2 * g]; py = hpos_init[2 * g + 1]; gx = hgoal_init ? hgoal_init[2 * g] : -px; gy = hgoal_init ? hgoal_init[2 * g + 1] : -py; }
                else {
                    px = py = 0.0;
                    for (uint32_t attempt = 0; attempt < 4096; ++attempt) {     /* crowd_sim.py:232-262 */
                        double u0, u1, u2, u3;
                        uniform2(c->seed, epoch, gid, RNG_RESET, 0u, 0ull, (uint32_t)i, 2 * attempt, &u0, &u1);
                        uniform2(c->seed, epoch, gid, RNG_RESET, 0u, 0ull, (uint32_t)i, 2 * attempt + 1, &u2, &u3);
                        const double ang = u0 * PI * 2.0;
                        px = c->circle_radius_humans * cos(ang) + u1 * 1.0;
                        py = c->circle_radius_humans * sin(ang) + u2 * 1.0;
                        int collide = 0;
                        { const double a = px - rx, b = py - ry; if (sqrt(a * a + b * b) < c->human_radius + c->robot_radius + 0.5) collide = 1; }
                        for (int j = 0; j < i && !collide; ++j) {
                            const double a = px - A->hpos[2 * ((size_t)e * Nh + j)], b = py - A->hpos[2 * ((size_t)e * Nh + j) + 1];
                            if (sqrt(a * a + b * b) < c->human_radius + c->human_radius + 0.2) collide = 1;
                        }
                        if (!collide) break;
                    }
                    gx = -px; gy = -py;
                }
                A->hpos[2 * g] = px; A->hpos[2 * g + 1] = py; A->hgoal[2 * g] = gx; A->hgoal[2 * g + 1] = gy;
                A->hvel[2 * g] = 0.f; A->hvel[2 * g + 1] = 0.f; A->hvalid[g] = 1;
                A->hstatic[g] = si >= 0 ? 1 : 0;
                if (c->sensor_restriction) {                                       /* human_avoidance_env.py:331-334 */
                    const double ct = cos(rth), st = sin(rth), dx = px - rx, dy = py - ry;
                    const double lx = ct * dx + st * dy, ly = -st * dx + ct * dy;
                    if (sqrt(lx * lx + ly * ly) > c->sensor_range) A->hvalid[g] = 0;
                }
                for (int s = 0; s < H; ++s) { A->hhist[(g * H + s) * 2] = px; A->hhist[(g * H + s) * 2 + 1] = py; }
            }
            pt_step(c, A, e, NULL, 1);
            for (int w = 0; w < LB; ++w) ha_step(c, A, e, NULL, 1, w == LB - 1, scratch, NULL);
        }
        free(scratch);
    }
}
