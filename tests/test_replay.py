"""Host logic of the batched replay storage (dr_mpc_b200.replay) against the semantics of the reference's
ReplayBuffer (scripts/utils/storage.py:6-133): same tensors per key, same sample() return, masked batch inserts into a
ring instead of one-at-a-time inserts with a roll. Runs on CPU tensors (pure torch indexing)."""
import numpy as np
import torch

import dr_mpc_b200 as d


def _obs(E, fill):
    return {'PT': {'PT_state': torch.full((E, 100), float(fill)), 'MPC_actions': torch.full((E, 8), float(fill))},
            'HA': {'spatial_edges': torch.full((E, 6, 14), float(fill)), 'robot_node': torch.full((E, 12), float(fill))}}


def test_layout_matches_reference_buffer():
    rb = d.BatchedReplayBuffer(_obs(4, 0), buffer_size=10)
    buf = rb.replay_buffer
    assert set(buf) == {'S', 'S_prime', 'rewards', 'actions', 'done'}
    assert buf['S']['PT']['PT_state'].shape == (10, 100) and buf['S_prime']['HA']['spatial_edges'].shape == (10, 6, 14)
    assert buf['rewards'].shape == (10, 1) and buf['actions'].shape == (10, 2) and buf['done'].shape == (10, 1)
    assert all(t.dtype == torch.float64 for t in (buf['rewards'], buf['S']['PT']['PT_state']))


def test_masked_batch_insert_ring_and_sample():
    E, N = 4, 10
    rb = d.BatchedReplayBuffer(_obs(E, 0), buffer_size=N)
    stored = []
    for t in range(7):
        S, Sp = _obs(E, t), _obs(E, t + 0.5)
        A = torch.arange(E * 2, dtype=torch.float64).reshape(E, 2) + 100 * t
        R = torch.arange(E, dtype=torch.float64) + 10 * t
        done = torch.tensor([False, True, False, False])
        mask = torch.tensor([True, t % 2 == 0, False, True])          # env 2 always in soft reset, env 1 every other step
        n = rb.insert(S, A, Sp, R, done, mask=mask)
        assert n == int(mask.sum())
        stored += [(t, e) for e in range(E) if mask[e]]
    assert rb.valid_entries == N and rb.full()
    newest = stored[-N:]                                              # the ring keeps the newest N transitions
    got = sorted((int(r // 10), int(r % 10)) for r in rb.replay_buffer['rewards'][:, 0].tolist())
    assert got == sorted(newest)
    idx = torch.arange(N)
    S, A, Sp, R, D = rb.sample_from_buffer(N, indices_to_sample=idx)
    t_of = (R[:, 0] // 10)
    assert torch.equal(S['PT']['PT_state'][:, 0], t_of) and torch.equal(Sp['HA']['spatial_edges'][:, 0, 0], t_of + 0.5)
    assert torch.equal(A[:, 0] // 100, t_of) and D.shape == (N, 1) and set(D[:, 0].tolist()) <= {0.0, 1.0}
    S, A, Sp, R, D = rb.sample_from_buffer(32)
    assert S['HA']['robot_node'].shape == (32, 12) and A.shape == (32, 2)


def test_oversized_batch_keeps_the_newest_rows():
    rb = d.BatchedReplayBuffer(_obs(8, 0), buffer_size=5)
    R = torch.arange(8, dtype=torch.float64)
    rb.insert(_obs(8, 1), torch.zeros(8, 2), _obs(8, 2), R, torch.zeros(8))
    assert sorted(rb.replay_buffer['rewards'][:, 0].tolist()) == [3.0, 4.0, 5.0, 6.0, 7.0]


def test_matches_reference_replay_buffer_on_the_same_trace():
    """The reference's own ReplayBuffer (scripts/utils/storage.py:6-133, imported unmodified; torch only) and
    BatchedReplayBuffer fed the same transitions -- one by one there, in masked batches here, through two overflows of the
    buffer -- hold the same entries in the same chronological order and return the same samples for the same indices."""
    import os
    import sys
    import types
    import torch
    ref_root = os.environ.get("DRE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "scripts")):
        pytest.skip("the reference tree is only in the build container")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_storage", os.path.join(ref_root, "scripts", "utils", "storage.py"))
    ref_storage = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_storage)
    import dr_mpc_b200 as d
    E, Nh, size = 7, 3, 40
    gen = torch.Generator().manual_seed(3)

    def obs(n):
        return {'PT': {'PT_state': torch.randn(n, 100, generator=gen, dtype=torch.float64), 'MPC_actions': torch.randn(n, 8, generator=gen, dtype=torch.float64)},
                'HA': {'spatial_edges': torch.randn(n, Nh, 14, generator=gen, dtype=torch.float64), 'robot_node': torch.randn(n, 12, generator=gen, dtype=torch.float64),
                       'detected_human_num': torch.full((n,), float(Nh), dtype=torch.float64)}}
    cm = types.SimpleNamespace(config_general=types.SimpleNamespace(device='cpu'))
    take = lambda S, i: {k: {kk: vv[i:i + 1] for kk, vv in v.items()} for k, v in S.items()}
    ref = ref_storage.ReplayBuffer(take(obs(1), 0), cm, size)
    ours = d.BatchedReplayBuffer(obs(E), buffer_size=size, device='cpu')
    for t in range(16):                                           # 16 x ~5 kept rows: the buffer overflows twice
        S, Sp = obs(E), obs(E)
        A = torch.randn(E, 2, generator=gen, dtype=torch.float64); R = torch.randn(E, generator=gen, dtype=torch.float64)
        done = (torch.rand(E, generator=gen) < 0.2)
        mask = torch.rand(E, generator=gen) < 0.75
        ours.insert(S, A, Sp, R, done, mask=mask)
        for i in torch.nonzero(mask).flatten().tolist():          # the reference inserts one transition per call
            ref.insert(take(S, i), A[i:i + 1], take(Sp, i), R[i].item(), float(done[i]))
    assert ref.valid_entries == ours.valid_entries == size and ours.full() and ref.full()
    order = ours.physical_index(torch.arange(size))
    for grp in ('S', 'S_prime'):
        for k, v in ref.replay_buffer[grp].items():
            for kk, vv in v.items():
                assert torch.equal(vv, ours.replay_buffer[grp][k][kk][order]), (grp, k, kk)
    for k in ('rewards', 'actions', 'done'):
        assert torch.equal(ref.replay_buffer[k], ours.replay_buffer[k][order]), k
    idx = torch.randint(0, size, (25,), generator=gen)
    a, b = ref.sample_from_buffer(25, idx), ours.sample_from_buffer(25, idx)
    assert torch.equal(a[1], b[1]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    for k in a[0]:
        for kk in a[0][k]:
            assert torch.equal(a[0][k][kk], b[0][k][kk]) and torch.equal(a[2][k][kk], b[2][k][kk])


def test_compat_replay_buffer_is_the_reference_class_call_for_call():
    """dr_mpc_b200.compat.ReplayBuffer takes the reference's constructor and call signatures (storage.py:8,54,101,135): the
    same single-transition trace through both, past an overflow, gives the same `replay_buffer` tensors in the same order,
    the same samples for the same indices and the same `grab_all_data` batches."""
    import os
    import types
    import importlib.util
    import torch
    ref_root = os.environ.get("DRE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "scripts")):
        pytest.skip("the reference tree is only in the build container")
    spec = importlib.util.spec_from_file_location("ref_storage2", os.path.join(ref_root, "scripts", "utils", "storage.py"))
    ref_storage = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_storage)
    from dr_mpc_b200 import compat
    gen = torch.Generator().manual_seed(11)
    Nh, size = 4, 9

    def obs():
        return {'PT': {'PT_state': torch.randn(1, 100, generator=gen, dtype=torch.float64), 'MPC_actions': torch.randn(1, 8, generator=gen, dtype=torch.float64)},
                'HA': {'spatial_edges': torch.randn(1, Nh, 14, generator=gen, dtype=torch.float64), 'percent_episode': torch.rand(1, 1, generator=gen, dtype=torch.float64),
                       'detected_human_num': torch.full((1,), float(Nh), dtype=torch.float64)}}
    cm = types.SimpleNamespace(config_general=types.SimpleNamespace(device='cpu'))
    ex = obs()
    ref, ours = ref_storage.ReplayBuffer(ex, cm, size), compat.ReplayBuffer(ex, cm, size)

    def same():
        assert ref.valid_entries == ours.valid_entries and ref.full() == ours.full()
        a, b = ref.replay_buffer, ours.replay_buffer
        for grp in ('S', 'S_prime'):
            for k, v in a[grp].items():
                for kk, vv in v.items():
                    assert torch.equal(vv, b[grp][k][kk]), (grp, k, kk)
        for k in ('rewards', 'actions', 'done'):
            assert torch.equal(a[k], b[k]), k

    for t in range(2 * size + 3):
        S, Sp = obs(), obs()
        A = torch.randn(1, 2, generator=gen, dtype=torch.float64)
        R, done = float(torch.randn((), generator=gen)), bool(torch.rand((), generator=gen) < 0.3)
        ref.insert(S, A, Sp, R, done)
        ours.insert(S, A, Sp, R, done)
        if t in (3, size - 1, size, 2 * size + 2):
            same()
    idx = torch.randint(0, size, (6,), generator=gen)
    for a, b in zip(ref.sample_from_buffer(6, idx), ours.sample_from_buffer(6, idx)):
        if isinstance(a, dict):
            for k in a:
                for kk in a[k]:
                    assert torch.equal(a[k][kk], b[k][kk])
        else:
            assert torch.equal(a, b)
    ga, gb = ref.grab_all_data(4), ours.grab_all_data(4)
    assert len(ga) == len(gb) == 3
    for x, y in zip(ga, gb):
        assert torch.equal(x[1], y[1]) and torch.equal(x[3], y[3]) and torch.equal(x[0]['PT']['PT_state'], y[0]['PT']['PT_state'])
    assert ours.sample_from_buffer(5)[1].shape == (5, 2)
