"""Batched replay storage with the reference ReplayBuffer's layout and sampling interface
(scripts/utils/storage.py:6-133), fed straight from the engine's device-resident step outputs.

The reference keeps one float64 tensor [buffer_size, *shape] per observation key under 'S' and 'S_prime', plus
'rewards' [N,1], 'actions' [N,2], 'done' [N,1], inserts ONE transition per call (storage.py:54-98; when full it rolls
every tensor by one) and samples with integer indices (:101-133). Here a whole batch of E transitions is inserted per
call, optionally masked (transitions that ran a scripted soft-reset action stay out, like the reference's
`soft_reset=True` steps that never reach `replay_buffer.insert`, online_continuous_task.py:143,189), into a ring: the
oldest entries are overwritten instead of rolling gigabytes. Sampling is uniform over the valid entries either way, so
the ring is indistinguishable from the roll for the learner. Pure torch indexing: plumbing, not a kernel."""


def _flatten(S):
    out = {}
    for k, v in S.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[(k, kk)] = vv
        else:
            out[(k,)] = v
    return out


def _nest(flat):
    out = {}
    for key, v in flat.items():
        if len(key) == 2:
            out.setdefault(key[0], {})[key[1]] = v
        else:
            out[key[0]] = v
    return out


class BatchedReplayBuffer:
    def __init__(self, obs_example, buffer_size, device=None, dtype=None):
        import torch
        flat = _flatten(obs_example)
        any_t = next(iter(flat.values()))
        self.device = torch.device(device) if device is not None else any_t.device
        self.dtype = dtype or torch.float64                       # storage.py:27: everything is float64
        self.buffer_size = int(buffer_size)
        self.valid_entries = 0
        self._head = 0
        mk = lambda shape: torch.zeros((self.buffer_size,) + tuple(shape), dtype=self.dtype, device=self.device)
        self._S = {k: mk(v.shape[1:]) for k, v in flat.items()}
        self._Sp = {k: mk(v.shape[1:]) for k, v in flat.items()}
        self.replay_buffer = {'S': _nest(self._S), 'S_prime': _nest(self._Sp), 'rewards': mk((1,)), 'actions': mk((2,)), 'done': mk((1,))}

    def full(self):
        return self.valid_entries == self.buffer_size

    def insert(self, S, A, S_prime, R, done, mask=None):
        """S / S_prime: observation dicts with a leading batch dim; A [E,2]; R [E] or [E,1]; done [E] or [E,1] (bool ok);
        mask [E] bool: rows to keep (e.g. ~info['soft_reset']). Returns the number of transitions stored."""
        import torch
        fS, fSp = _flatten(S), _flatten(S_prime)
        for k in fS:   # zero-copy views of the same slab: S was overwritten by the step that produced S_prime
            if fS[k].data_ptr() == fSp[k].data_ptr() and fS[k].numel() > 0:
                raise ValueError(f"S and S_prime alias the same memory ({'/'.join(k)}): pass the observation returned by the "
                                 "previous step (the engine double-buffers its output slab), or clone S")
        n_all = A.shape[0]
        if mask is not None:
            rows = torch.nonzero(mask.to(self.device), as_tuple=False).squeeze(1)
        else:
            rows = torch.arange(n_all, device=self.device)
        n = int(rows.numel())
        if n == 0:
            return 0
        if n > self.buffer_size:                                  # only the newest buffer_size rows can survive
            rows = rows[-self.buffer_size:]
            n = self.buffer_size
        slots = (self._head + torch.arange(n, device=self.device)) % self.buffer_size
        for k in self._S:
            self._S[k].index_copy_(0, slots, fS[k].to(self.device, self.dtype).index_select(0, rows))
            self._Sp[k].index_copy_(0, slots, fSp[k].to(self.device, self.dtype).index_select(0, rows))
        col = lambda t, w: t.to(self.device, self.dtype).reshape(n_all, w).index_select(0, rows)
        self.replay_buffer['rewards'].index_copy_(0, slots, col(R, 1))
        self.replay_buffer['actions'].index_copy_(0, slots, col(A, 2))
        self.replay_buffer['done'].index_copy_(0, slots, col(done, 1))
        self._head = (self._head + n) % self.buffer_size
        self.valid_entries = min(self.buffer_size, self.valid_entries + n)
        return n

    def physical_index(self, chronological):
        """Ring slot of the `chronological`-th oldest valid entry. The reference keeps its tensors in insertion order (it rolls
        everything by one when full, storage.py:56-75): its index j is this buffer's slot (head + j) % size once full."""
        if self.valid_entries < self.buffer_size:
            return chronological
        return (chronological + self._head) % self.buffer_size

    def sample_from_buffer(self, batch_size, indices_to_sample=None):
        """Same return as storage.py:101-133: (S, actions, S_prime, rewards, done) gathered at random valid indices;
        `indices_to_sample` are chronological (0 = oldest), like the reference's."""
        import torch
        if indices_to_sample is None:
            indices_to_sample = torch.randint(0, max(self.valid_entries, 1), (batch_size,), device=self.device)
        idx = self.physical_index(indices_to_sample.to(self.device))
        rb = self.replay_buffer
        return (_nest({k: v[idx] for k, v in self._S.items()}), rb['actions'][idx], _nest({k: v[idx] for k, v in self._Sp.items()}),
                rb['rewards'][idx], rb['done'][idx])
