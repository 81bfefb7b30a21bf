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

// PTEnvCore.localize_on_path (reference path_tracking_core.py:21-81) with strict=False; the k=2 KD-tree query is a
// brute-force scan (paths have <= 125 nodes).
DRE_HD Localization localize_on_path(const PathView &P, double qx, double qy, bool localize_to_start)
{
    const int L = P.L;
    int i0 = 0, i1 = 1;
    double d0 = INFINITY, d1 = INFINITY;
    for (int i = 0; i < L; ++i) {
        const double dx = P.x[i] - qx, dy = P.y[i] - qy;
        const double d = dx * dx + dy * dy;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
        else if (d < d1) { d1 = d; i1 = i; }
    }
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

// PTEnvCore.state_gen_MLP_local_format, continuous-task branch (path_tracking_core.py:94-145):
// (lookbehind+lookahead) nodes around ceil(interp_index), wrapped modulo L, as [x.., y.., cos.., sin..] in the robot frame.
DRE_HD void pt_state_gen(const PathView &P, double px, double py, double pth, double interp_index, int lookahead,
                         int lookbehind, double *out, int out_stride)
{
    double st, ct;
    sincos(pth, &st, &ct);
    const int n = lookahead + lookbehind;
    const int next = (int)ceil(interp_index);
    const int start = next - lookbehind;
    for (int j = 0; j < n; ++j) {
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

// uniform doubles for (env, step, stream id, draw): two per Philox call
DRE_HD void env_uniform2(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t stream, uint32_t draw, double &u0, double &u1)
{
    uint32_t r[4];
    philox4x32_10(seed, env_id, (step << 24) ^ ((uint64_t)stream << 48) ^ (uint64_t)draw, r);
    u0 = u01_from_bits(r[0], r[1]);
    u1 = u01_from_bits(r[2], r[3]);
}
