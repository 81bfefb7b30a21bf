"""GPU: the device-resident replay ring (include/dre.h dre_replay_*, dr_mpc_b200.DeviceReplayBuffer) against the torch-indexing
BatchedReplayBuffer -- which tests/test_replay.py pins to the reference's own ReplayBuffer (scripts/utils/storage.py:6-133) --
on the same free-running rollout: same rows in the same slots, bit for bit; float32 = one rounding of the float64 rows."""
import pytest

pytestmark = pytest.mark.gpu


def _flat(S):
    return {(k, kk): vv for k, v in S.items() for kk, vv in v.items()}


def _rollout(env, steps, on_step):
    import torch
    S = env.reset()
    E = env.E
    for t in range(steps):
        a = S['PT']['MPC_actions'][:, :2].clone()
        a[:, 1] += 0.4 * torch.sin(torch.arange(E, device=env.device) * 0.37 + t)      # a sloppy policy: terminations happen
        a[:, 1].clamp_(-1, 1)
        S_prime, R, D, info, eps = env.step(a)
        on_step(t, S, a, S_prime, R, D, info)
        S = S_prime


@pytest.mark.parametrize("E,Nh,capacity", [(512, 6, 1500), (512, 6, 20000), (128, 50, 700), (200, 7, 333), (1100, 3, 5000)])
def test_device_ring_equals_torch_ring(E, Nh, capacity):
    """Masked inserts over 30 steps with the soft-reset machine on (scripted steps stay out), through several wrap-arounds
    of the small ring: every stored tensor, the counters and chronological sampling agree with BatchedReplayBuffer. 50 humans =
    a row longer than one tile of copy units; 7 and 3 humans = odd-width fields (single-element units); 1100 envs = two
    selection CTAs, the second one partly empty."""
    import torch
    import dr_mpc_b200 as d
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, soft_reset_on_device=True), num_envs=E)
    S0 = env.reset()
    ref = d.BatchedReplayBuffer(S0, buffer_size=capacity)
    ours = d.DeviceReplayBuffer(env, capacity)
    stored = [0, 0]

    def on_step(t, S, a, S_prime, R, D, info):
        keep = (torch.arange(E, device=env.device) + t) % 5 != 0
        a_model = a + 0.01                                        # the stored action may differ from the executed one (:178-181)
        stored[0] += ref.insert(S, a_model, S_prime, R['R'], D['done_for_replay'], mask=keep & ~info['soft_reset'])
        ours.insert(a_model, mask=keep)
        stored[1] += ours.last_inserted
    _rollout(env, 30, on_step)
    torch.cuda.synchronize()
    assert stored[0] == stored[1] > min(capacity // 2, 20 * E)
    assert ours.valid_entries == ref.valid_entries == min(stored[0], capacity)
    assert ours.physical_index(0) == ref.physical_index(0)
    fr, fo = ref.replay_buffer, ours.replay_buffer
    for grp in ('S', 'S_prime'):
        for key, t_ref in _flat(fr[grp]).items():
            t_our = _flat(fo[grp])[key]
            assert t_our.shape == t_ref.shape and t_our.dtype == t_ref.dtype, (grp, key)
            assert torch.equal(t_our, t_ref), (grp, key)
    for key in ('rewards', 'actions', 'done'):
        assert torch.equal(fo[key], fr[key]), key
    assert (fo['S']['HA']['spatial_edges'] != fo['S_prime']['HA']['spatial_edges']).any()
    if (E, Nh) == (512, 6):
        assert fo['done'].sum() > 0
    # sampling at chronological indices (storage.py:101-133), float64 and float32 hand-out
    idx = torch.randint(0, ours.valid_entries, (777,), device=env.device)
    a64 = ours.sample_from_buffer(777, idx)
    r64 = ref.sample_from_buffer(777, idx)
    a32 = ours.sample_from_buffer(777, idx, dtype=torch.float32)
    for got, want, got32 in zip(a64, r64, a32):
        if isinstance(want, dict):
            for key, w in _flat(want).items():
                assert torch.equal(_flat(got)[key], w), key
                assert _flat(got32)[key].dtype == torch.float32 and torch.equal(_flat(got32)[key], w.float()), key
        else:
            assert torch.equal(got, want) and torch.equal(got32, want.float())
    ga, gr = ours.grab_all_data(1000), ref.grab_all_data(1000)                # storage.py:135-145 (OOD.py:120)
    assert len(ga) == len(gr) == -(-ours.valid_entries // 1000)
    assert all(torch.equal(x[1], y[1]) and torch.equal(x[0]['HA']['spatial_edges'], y[0]['HA']['spatial_edges']) for x, y in zip(ga, gr))
    Sb = ours.sample_from_buffer(64)[0]
    assert Sb['HA']['spatial_edges'].shape == (64, Nh, 14) and Sb['PT']['PT_state'].shape == (64, 100)


def test_float32_ring_and_insert_larger_than_the_ring():
    """capacity < E: only the newest `capacity` rows of an insert survive, in order (the reference would have rolled the
    others out, storage.py:56-75); a float32 ring holds each float64 value rounded once; no mask = every env."""
    import torch
    import dr_mpc_b200 as d
    E, cap = 300, 100
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=5), num_envs=E)
    S0 = env.reset()
    ref = d.BatchedReplayBuffer(S0, buffer_size=cap)
    ours = d.DeviceReplayBuffer(env, cap, dtype=torch.float32)
    bytes64 = sum(v.numel() * 8 for g in ('S', 'S_prime') for v in _flat(ref.replay_buffer[g]).values())
    assert 0.5 * bytes64 <= ours.storage_bytes < 0.6 * bytes64

    def on_step(t, S, a, S_prime, R, D, info):
        ref.insert(S, info['actions_used'], S_prime, R['R'], D['done_for_replay'])
        ours.insert()
        assert ours.last_inserted == cap
    _rollout(env, 4, on_step)
    torch.cuda.synchronize()
    assert ours.full() and ref.full() and ours.physical_index(0) == ref.physical_index(0)
    for grp in ('S', 'S_prime'):
        for key, t_ref in _flat(ref.replay_buffer[grp]).items():
            assert torch.equal(_flat(ours.replay_buffer[grp])[key], t_ref.float()), (grp, key)
    for key in ('rewards', 'actions', 'done'):
        assert torch.equal(ours.replay_buffer[key], ref.replay_buffer[key].float()), key
    with pytest.raises(Exception):
        ours.sample_from_buffer(8, dtype=torch.float64)          # a float32 ring cannot hand float64 back
    ours.clear()
    assert ours.valid_entries == 0


def test_ring_outlives_the_python_objects():
    import gc
    import torch
    import dr_mpc_b200 as d
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=4), num_envs=64)
    S = env.reset()
    rb = d.DeviceReplayBuffer(env, 256)
    S2, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    rb.insert()
    torch.cuda.synchronize()
    t = rb.replay_buffer['S_prime']['HA']['spatial_edges']
    want = t.clone()
    assert torch.equal(want[:64], S2['HA']['spatial_edges'])
    del rb, env, S, S2
    gc.collect()
    torch.zeros(1 << 20, device="cuda")
    assert torch.equal(t, want)


def test_sample_rejects_indices_outside_the_ring():
    """storage.py:108-133 indexes its tensors with the caller's indices, so a bad index raises; the gather kernel would read
    past the ring instead, hence the host-side range check (IndexError) in front of it."""
    import torch
    import dr_mpc_b200 as d
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=4), num_envs=32)
    S = env.reset()
    rb = d.DeviceReplayBuffer(env, 96)
    env.step(S['PT']['MPC_actions'][:, :2].clone())
    rb.insert()
    for bad in ([0, 96], [-1, 3], torch.tensor([5, 1 << 40])):
        with pytest.raises(IndexError):
            rb.sample_from_buffer(2, indices_to_sample=bad)
    assert rb.sample_from_buffer(3, indices_to_sample=[0, 31, 95])[1].shape == (3, 2)     # inside the ring: fine
    assert len(rb.grab_all_data(10)) == 4 and rb.grab_all_data(10)[-1][1].shape == (2, 2)


def test_insert_then_host_step_reads_the_slab_before_it_is_rewritten():
    """dre_replay_insert (caller's stream) reads BOTH output slabs; the dre_env_step_host that follows runs on the engine's
    own streams and writes the older slab: the engine orders it after the insert. Checked against the same loop with a
    device-wide synchronise between the two calls."""
    import numpy as np
    import torch
    import dr_mpc_b200 as d
    rings = []
    for sync in (False, True):
        env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20, seed=11), num_envs=2048)
        S = env.reset()
        rb = d.DeviceReplayBuffer(env, 4 * 2048)
        a = S['PT']['MPC_actions'][:, :2].cpu().numpy().copy()
        for t in range(4):
            out = env.step_host(a)
            rb.insert()
            if sync:
                torch.cuda.synchronize()
            a = np.ascontiguousarray(out['mpc_actions'][:, :2]).copy()
        torch.cuda.synchronize()
        rings.append({k: v.clone() for k, v in _flat(rb.replay_buffer['S']).items()} | {('A', 'A'): rb.replay_buffer['actions'].clone()})
        assert rb.valid_entries == 4 * 2048
    for k in rings[0]:
        assert torch.equal(rings[0][k], rings[1][k]), k
