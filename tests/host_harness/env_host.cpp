// tests/host_harness/env_host.cpp -- TEST INFRASTRUCTURE. Compiles the PRODUCT's host+device env header
// (dr-mpc_b200/csrc/env_device.cuh: localisation, soft-reset state machine) for the host so the CPU-only test suite can
// replay the reference's recorded soft_reset() runs through the kernel's own decision code. Never shipped.
#include <cuda_runtime.h>
#include "../../dr-mpc_b200/csrc/env_device.cuh"

// Replays T consecutive env.step calls of ONE env: prev_bits[t] are the done bits of call t-1 (0 for t = 0), pose[t] the
// robot pose before call t. Writes scripted[t] (0/1) and action[t] (only meaningful where scripted). The machine's
// state persists over the calls, exactly as on the device.
extern "C" int host_soft_reset_replay(int T, const uint32_t *prev_bits, const double *pose /*[T][3]*/, const int32_t *path_idx,
                                      const uint8_t *lts, const double *paths, int stride, const int32_t *lens,
                                      uint8_t *scripted, double *action /*[T][2]*/, uint8_t *stuck_any)
{
    int32_t packed[4] = {0, 0, 0, 0};
    *stuck_any = 0;
    for (int t = 0; t < T; ++t) {
        SrState S = sr_unpack(packed);
        const PathView P = path_view(paths, stride, path_idx[t], lens);
        double av = 0.0, aw = 0.0;
        bool stuck = false;
        scripted[t] = (uint8_t)soft_reset_decide(S, prev_bits[t], P, pose[3 * t], pose[3 * t + 1], pose[3 * t + 2], lts[t] != 0, av, aw, stuck);
        action[2 * t] = av; action[2 * t + 1] = aw;
        if (stuck) *stuck_any = 1;
        sr_pack(S, packed);
    }
    return 0;
}

// Robot half building blocks on recorded states (tests/test_host.py::test_robot_half_blocks_on_host_match_reference_rollout):
// out_pose[t] = unicycle_step(pose[t], action[t]) (agent.py:169-176 -> unicycle.py:89-126).
extern "C" int host_unicycle(int T, const double *pose /*[T][3]*/, const double *action /*[T][2]*/, double dt, double *out_pose /*[T][3]*/)
{
    for (int t = 0; t < T; ++t)
        unicycle_step(pose[3 * t], pose[3 * t + 1], pose[3 * t + 2], action[2 * t], action[2 * t + 1], dt, out_pose[3 * t], out_pose[3 * t + 1],
                      out_pose[3 * t + 2]);
    return 0;
}

// index[t] = localize_on_path(pose[t]).index (path_tracking_core.py:21-81), pt_state[t] = state_gen_MLP_local_format at that
// index (path_tracking_core.py:94-145) laid out [x.., y.., cos.., sin..] like the reference's 100-vector.
extern "C" int host_pt_state(int T, const double *pose /*[T][3]*/, const int32_t *path_idx, const uint8_t *lts, const double *paths, int stride,
                             const int32_t *lens, int lookahead, int lookbehind, double *index /*[T]*/, double *pt_state /*[T][4 (la + lb)]*/)
{
    const int n = 4 * (lookahead + lookbehind);
    for (int t = 0; t < T; ++t) {
        const PathView P = path_view(paths, stride, path_idx[t], lens);
        const Localization Lc = localize_on_path(P, pose[3 * t], pose[3 * t + 1], lts[t] != 0);
        index[t] = Lc.index;
        pt_state_gen(P, pose[3 * t], pose[3 * t + 1], pose[3 * t + 2], Lc.index, lookahead, lookbehind, pt_state + (size_t)t * n, 1);
    }
    return 0;
}
