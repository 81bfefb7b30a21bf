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
