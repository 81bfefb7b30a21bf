"""tools/probe_replay.py: the replay insert / sample kernels at the C3 shape (16384 envs x 20 humans) against their HBM
roofline and against the torch-indexing BatchedReplayBuffer doing the same insert. Writes one JSON object.
    python tools/probe_replay.py [--envs 16384] [--humans 20] [--capacity 1000000] > gpurun_out/replay.json"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import dr_mpc_b200 as d
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--humans", type=int, default=20)
    ap.add_argument("--capacity", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=4096)
    a = ap.parse_args()
    E = a.envs
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=a.humans), num_envs=E)
    S = env.reset()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else {}
    out = {"envs": E, "humans": a.humans, "capacity": a.capacity, "reps": a.reps, "measured_peaks": peaks}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)

    def timed(fn, reps):
        ms = []
        for _ in range(reps):
            flush.fill_(1)                                   # 256 MB > L2: every rep starts with a cold L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        return {"median_ms": ms[len(ms) // 2], "min_ms": ms[0], "max_ms": ms[-1]}

    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        rb = d.DeviceReplayBuffer(env, a.capacity, dtype=dt)
        W = sum(rb._widths)
        es = 8 if dt == torch.float64 else 4
        S_prime, R, D, info, _ = env.step(S['PT']['MPC_actions'][:, :2].clone())
        for _ in range(3):
            rb.insert()
        torch.cuda.synchronize()
        r = timed(lambda: rb.insert(), a.reps)
        # algorithmic bytes of one insert: every kept row reads S and S' (2 W float64) + (A, R, done flag) and writes them once
        bytes_ins = E * (2 * W * 8 + 2 * 8 + 8 + 1 + 4) + E * (2 * W + 4) * es
        r.update({"rows": E, "row_doubles": 2 * W + 4, "algorithmic_bytes": bytes_ins, "GBps": bytes_ins / r["median_ms"] / 1e6,
                  "storage_bytes": rb.storage_bytes})
        out[f"insert_{name}"] = r
        # as in the data-collection loop: straight after the step that wrote S' (its slab is partly L2-resident), no flush
        ms = []
        Sx = S
        for _ in range(a.reps):
            Sx, *_ = env.step(Sx['PT']['MPC_actions'][:, :2].clone())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rb.insert()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        out[f"insert_{name}_after_step"] = {"median_ms": ms[len(ms) // 2], "min_ms": ms[0], "max_ms": ms[-1], "rows": E}
        m = (torch.arange(E, device=env.device) % 3 != 0)
        r = timed(lambda: rb.insert(mask=m), a.reps)
        kept = int(m.sum())
        b = kept * (2 * W * 8 + 29) + kept * (2 * W + 4) * es + E * 2
        r.update({"rows": kept, "algorithmic_bytes": b, "GBps": b / r["median_ms"] / 1e6})
        out[f"insert_masked_{name}"] = r
        for oname, odt in (("f64", torch.float64), ("f32", torch.float32)):
            if dt == torch.float32 and odt == torch.float64:
                continue
            idx = torch.randint(0, rb.valid_entries, (a.batch,), device=env.device)
            for _ in range(3):
                rb.sample_from_buffer(a.batch, idx, dtype=odt)
            r = timed(lambda: rb.sample_from_buffer(a.batch, idx, dtype=odt), a.reps)
            ob = 8 if odt == torch.float64 else 4
            b = a.batch * (2 * W + 4) * (es + ob) + a.batch * 8
            r.update({"rows": a.batch, "algorithmic_bytes": b, "GBps": b / r["median_ms"] / 1e6,
                      "note": "includes the 19 torch.empty allocations of the batch"})
            out[f"sample_{name}_to_{oname}"] = r
        del rb
    # the torch-indexing ring doing the same insert (round 1's path: ~40 index_select / index_copy_ launches)
    tb = d.BatchedReplayBuffer(S, buffer_size=min(a.capacity, 200000))
    S_prime, R, D, info, _ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    prev = env._outs[env._L.dre_env_current_slab(env._h) ^ 1]
    S_prev = {'PT': {'PT_state': prev["pt_state"], 'MPC_actions': prev["mpc_actions"]},
              'HA': {'robot_node': prev["robot_node"], 'past_robot_velocities': prev["past_robot_vel"],
                     'percent_episode': prev["percent_episode"], 'spatial_edges': prev["spatial_edges"],
                     'detected_human_num': prev["detected_humans"], 'human_valid_history': prev["human_valid"]}}
    for _ in range(3):
        tb.insert(S_prev, info['actions_used'], S_prime, R['R'], D['done_for_replay'])
    out["insert_torch_indexing_f64"] = timed(lambda: tb.insert(S_prev, info['actions_used'], S_prime, R['R'], D['done_for_replay']), a.reps)
    m = (torch.arange(E, device=env.device) % 3 != 0)
    out["insert_masked_torch_indexing_f64"] = timed(lambda: tb.insert(S_prev, info['actions_used'], S_prime, R['R'], D['done_for_replay'], mask=m), a.reps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
