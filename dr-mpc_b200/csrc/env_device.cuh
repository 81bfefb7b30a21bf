// env_device.cuh -- scalar fp64 building blocks of the env step (host+device so the CPU test-suite can
// exercise them): exact-arc unicycle, path localisation, PT state, PT rewards, Philox goal sampling.
#pragma once
#include "dre_common.cuh"

// numpy.mod semantics for a positive divisor
DRE_HD double np_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}
DRE_HD double wrap_pi(double th) { return np_mod(th + DRE_PI, 2.0 * DRE_PI) - DRE_PI; }

// Unicycle.step_external (reference environment/path_tracking/utils/unicycle.py:89-126): exact arc,
// straight line iff w == 0 exactly, theta wrapped to [-pi, pi).
DRE_HD void unicycle_step(double x, double y, double th, double v, double w, double dt, double &ox, double &oy, double &oth)
{
    double dxl, dyl, dth;
    if (w == 0.0) { dxl = v * dt; dyl = 0.0; dth = 0.0; }
    else {
        const double r = v / w;
        double sn, cs;
        sincos(w * dt, &sn, &cs);
        dxl = r * sn; dyl = r * (1.0 - cs); dth = w * dt;
    }
    double st, ct;
    sincos(th, &st, &ct);
    ox = x + (ct * dxl - st * dyl);
    oy = y + (st * dxl + ct * dyl);
    oth = wrap_pi(th + dth);
}

struct PathView {
    const double *x, *y, *th;   // rows of the reference's 3xL array
    int L;
};
DRE_HD PathView path_view(const double *paths, int stride, int pid, const int32_t *lens)
{
    PathView p;
    p.x = paths + (size_t)pid * 3 * stride; p.y = p.x + stride; p.th = p.y + stride; p.L = lens[pid];
    return p;
}

struct Localization { double ix, iy, ith, index; };

// The k=2 KD-tree query of localize_on_path as a brute-force scan (paths have <= 125 nodes): the two nearest nodes,
// ties resolved towards the lower index (lexicographic order on (distance^2, index)).
DRE_HD void path_nearest_two(const PathView &P, double qx, double qy, int &i0, int &i1)
{
    i0 = 0; i1 = 1;
    double d0 = INFINITY, d1 = INFINITY;
    for (int i = 0; i < P.L; ++i) {
        const double dx = P.x[i] - qx, dy = P.y[i] - qy;
        const double d = dx * dx + dy * dy;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
        else if (d < d1) { d1 = d; i1 = i; }
    }
}

// PTEnvCore.localize_on_path (reference path_tracking_core.py:21-81) with strict=False, given the two nearest nodes.
DRE_HD Localization localize_from_nearest(const PathView &P, double qx, double qy, bool localize_to_start, int i0, int i1)
{
    const int L = P.L;
    Localization R;
    if (localize_to_start) {
        if (i0 > 5 * L / 8 || i1 > 5 * L / 8) { R.ix = P.x[0]; R.iy = P.y[0]; R.ith = P.th[0]; R.index = 0.0; return R; }
    } else {
        if (i0 < L / 6 || i1 < L / 6) { R.ix = P.x[L - 1]; R.iy = P.y[L - 1]; R.ith = P.th[L - 1]; R.index = (double)(L - 1); return R; }
    }
    const int ad = i0 > i1 ? i0 - i1 : i1 - i0;
    if (ad != 1 && ad != L - 1) { R.ix = P.x[i0]; R.iy = P.y[i0]; R.ith = P.th[i0]; R.index = (double)i0; return R; }
    if (i0 == 0) {
        if ((qx - P.x[0]) * (P.x[1] - P.x[0]) + (qy - P.y[0]) * (P.y[1] - P.y[0]) < 0.0) {
            R.ix = P.x[0]; R.iy = P.y[0]; R.ith = P.th[0]; R.index = 0.0; return R;
        }
    } else if (i0 == L - 1) {
        if ((qx - P.x[L - 1]) * (P.x[L - 2] - P.x[L - 1]) + (qy - P.y[L - 1]) * (P.y[L - 2] - P.y[L - 1]) < 0.0) {
            R.ix = P.x[L - 1]; R.iy = P.y[L - 1]; R.ith = P.th[L - 1]; R.index = (double)(L - 1); return R;
        }
    }
    double bx = P.x[i1] - P.x[i0], by = P.y[i1] - P.y[i0];
    const double dx = qx - P.x[i0], dy = qy - P.y[i0];
    double proj = (dx * bx + dy * by) / (bx * bx + by * by);
    int other = i1;
    if (proj < 0.0) {
        other = ((i0 - (i1 - i0)) % L + L) % L;
        bx = P.x[other] - P.x[i0]; by = P.y[other] - P.y[i0];
        proj = (dx * bx + dy * by) / (bx * bx + by * by);
    }
    R.index = (double)i0 + proj * (double)(other - i0);
    R.ix = P.x[i0] + proj * (P.x[other] - P.x[i0]);
    R.iy = P.y[i0] + proj * (P.y[other] - P.y[i0]);
    R.ith = P.th[i0] + proj * (P.th[other] - P.th[i0]);
    return R;
}

DRE_HD Localization localize_on_path(const PathView &P, double qx, double qy, bool localize_to_start)
{
    int i0, i1;
    path_nearest_two(P, qx, qy, i0, i1);
    return localize_from_nearest(P, qx, qy, localize_to_start, i0, i1);
}

// PTEnvCore.state_gen_MLP_local_format, continuous-task branch (path_tracking_core.py:94-145):
// (lookbehind+lookahead) nodes around ceil(interp_index), wrapped modulo L, as [x.., y.., cos.., sin..] in the robot frame.
// (j0, jstep): the nodes this caller fills -- (0, 1) for all of them, (lane, lanes) when a lane-group shares the work
DRE_HD void pt_state_gen(const PathView &P, double px, double py, double pth, double interp_index, int lookahead,
                         int lookbehind, double *out, int out_stride, int j0 = 0, int jstep = 1)
{
    double st, ct;
    sincos(pth, &st, &ct);
    const int n = lookahead + lookbehind;
    const int next = (int)ceil(interp_index);
    const int start = next - lookbehind;
    for (int j = j0; j < n; j += jstep) {
        const int idx = ((start + j) % P.L + P.L) % P.L;
        const double dx = P.x[idx] - px, dy = P.y[idx] - py;
        double sr, cr;
        sincos(P.th[idx] - pth, &sr, &cr);
        out[(size_t)j * out_stride] = ct * dx + st * dy;
        out[(size_t)(n + j) * out_stride] = -st * dx + ct * dy;
        out[(size_t)(2 * n + j) * out_stride] = cr;
        out[(size_t)(3 * n + j) * out_stride] = sr;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// HAAndPTEnv.soft_reset (reference HA_and_PT env.py:120-240) as a per-env state machine. The reference loops
// `while done: pick a scripted action; self.step(action, soft_reset=True)` for ONE env; a batch steps in lock-step,
// so the loop is unrolled over step() calls: before every step the machine looks at the previous step's done bits
// and either hands control back to the caller's action (return 0) or supplies the scripted action (return 1).
// ---------------------------------------------------------------------------------------------------------------
enum { DRE_SR_NORMAL = 0, DRE_SR_MAIN = 1, DRE_SR_EXTRA = 2 };
struct SrState {
    int mode;     // NORMAL: not in soft_reset(); MAIN: inside the `while done` loop; EXTRA: inside the 7-step back-up loop (:227-233)
    int oot;      // out_of_trigger_because_of_backup (:126)
    int extra;    // back-up steps left in the EXTRA loop
    int consec;   // consecutive_no_v (:122)
    int steps;    // steps_in_soft_reset (:123)
    int nb;       // num_backup_steps (:127)
};
DRE_HD SrState sr_unpack(const int32_t *p)
{
    SrState S;
    S.mode = p[0] & 0xff; S.oot = (p[0] >> 8) & 0xff; S.extra = (p[0] >> 16) & 0xff;
    S.consec = p[1]; S.steps = p[2]; S.nb = p[3];
    return S;
}
DRE_HD void sr_pack(const SrState &S, int32_t *p)
{
    p[0] = S.mode | (S.oot << 8) | (S.extra << 16); p[1] = S.consec; p[2] = S.steps; p[3] = S.nb;
}

// prev_bits: DRE_DONE_* word of the previous step of this env (0 right after reset). (x, y, th): current robot pose.
// P / lts: current path and localize_to_start flag. Returns 1 and sets (av, aw) when the step must run a scripted
// action; `stuck` reports the reference's "Stuck in soft reset" exit (> 250 steps, :235-238).
DRE_HD int soft_reset_decide(SrState &S, uint32_t prev_bits, const PathView &P, double x, double y, double th, bool lts,
                             double &av, double &aw, bool &stuck)
{
    stuck = false;
    const bool prev_done = (prev_bits & DRE_DONE_ANY) != 0;
    const uint32_t reasons = prev_bits & 0xFFu;
    if (S.mode == DRE_SR_MAIN) {
        // end of a `while` iteration: the extra back-up loop starts if the main step was a back-up that cleared the triggers
        if (!prev_done && S.oot) { S.mode = DRE_SR_EXTRA; S.extra = 7; }
        else {
            if (S.steps > 250) { stuck = true; S.mode = DRE_SR_NORMAL; return 0; }
            if (!prev_done) { S.mode = DRE_SR_NORMAL; return 0; }          // `while done` ends: S' goes back to the caller
        }
    }
    if (S.mode == DRE_SR_EXTRA) {
        if (!prev_done && S.extra > 0) { S.extra -= 1; av = -0.3; aw = 0.0; return 1; }   // :229-233
        if (S.steps > 250) { stuck = true; S.mode = DRE_SR_NORMAL; return 0; }
        if (!prev_done) { S.mode = DRE_SR_NORMAL; return 0; }
        S.mode = DRE_SR_MAIN;                                                // a trigger fired during the back-up: loop on
    }
    if (S.mode == DRE_SR_NORMAL) {
        if (!prev_done) return 0;
        S.consec = 0; S.steps = 0; S.nb = 0; S.oot = 0; S.extra = 0;        // soft_reset() entered (:121-127)
    }
    // ---- top of a `while done` iteration (:128-216) ----
    // goal / end of path: the path was already switched inside the step that raised it (auto_path_switch, :130-141)
    if ((reasons & (DRE_DONE_GOAL | DRE_DONE_END_OF_PATH)) != 0u || reasons == DRE_DONE_ACTUATION) {   // :184-185
        S.mode = DRE_SR_NORMAL;
        return 0;
    }
    S.mode = DRE_SR_MAIN;
    if ((prev_bits & DRE_DONE_HA) != 0u && S.consec > 10 && S.nb < 5) {     // :187-190
        av = -0.3; aw = 0.0;
        S.nb += 1; S.oot = 1;
    } else {
        S.nb = 0; S.oot = 0;
        // query_closest_path_node_plus_one (path_tracking_env.py:221-225), turn towards it (:194-206)
        const Localization loc = localize_on_path(P, x, y, lts);
        const int idx = (int)(loc.index + 1.0) % P.L;
        const double desired = atan2(P.y[idx] - y, P.x[idx] - x);
        const double th_diff = np_mod(desired - th + DRE_PI, 2.0 * DRE_PI) - DRE_PI;
        const double w = th_diff < 0.0 ? -0.1 : 0.1;
        if ((reasons & (DRE_DONE_CORRIDOR_HIT | DRE_DONE_SAFETY_CORRIDOR)) != 0u && (reasons & DRE_DONE_SAFETY_HUMAN) == 0u) {
            if (fabs(th_diff) > 0.3) { av = 0.0; aw = w * 7.0; }
            else { av = 0.5; aw = w; }
            S.consec = 0;
        } else {
            av = 0.0; aw = 0.0;
            S.consec += 1;
        }
    }
    S.steps += 1;
    return 1;
}

// Uniform doubles for (env, reset epoch, kind of draw, step, stream id, draw): two per Philox call. Every field has its
// own bits of the 64-bit key / 128-bit counter, so no two draws of a run share a counter:
//   key     = seed, with the reset epoch added to its high word
//   c0, c1  = global env id
//   c2      = step count of the episode (low 32 bits)
//   c3      = draw (bits 0..15) | stream (16..23: the human index, or 0xC7 for the safety filter) | kind (24..25) |
//             sub-step (26..31: the warm-start step of a reset)
enum { DRE_RNG_STEP = 0, DRE_RNG_RESET = 1, DRE_RNG_WARM = 2, DRE_RNG_CVMM = 3 };
DRE_HD void env_uniform2(uint64_t seed, uint32_t epoch, uint64_t env_id, uint32_t kind, uint32_t sub, uint64_t step, uint32_t stream,
                         uint32_t draw, double &u0, double &u1)
{
    uint32_t r[4];
    const uint64_t key = seed + ((uint64_t)epoch << 32);
    const uint32_t c3 = (draw & 0xFFFFu) | ((stream & 0xFFu) << 16) | ((kind & 3u) << 24) | ((sub & 0x3Fu) << 26);
    philox4x32_10(key, env_id, (uint64_t)(uint32_t)step | ((uint64_t)c3 << 32), r);
    u0 = u01_from_bits(r[0], r[1]);
    u1 = u01_from_bits(r[2], r[3]);
}
