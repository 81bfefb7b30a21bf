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
