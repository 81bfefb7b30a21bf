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


    def grab_all_data(self, batch_size):
        """storage.py:135-145 (the OOD model's pass over the whole buffer, scripts/utils/OOD.py:120): every valid row once, in
        chronological order, in batches of `batch_size`."""
        import torch
        out, cur, n = [], 0, self.valid_entries
        while cur < n:
            end = min(cur + batch_size, n)
            out.append(self.sample_from_buffer(0, indices_to_sample=torch.arange(cur, end, device=self.device)))
            cur = end
        return out

class _ReplayHandle:
    """Owns the dre_replay handle (and keeps the env handle it reads from alive)."""

    def __init__(self, L, h, env_owner):
        self.L, self.h, self.env_owner = L, h, env_owner

    def __del__(self):
        if self.h is not None and self.L is not None:
            self.L.dre_replay_destroy(self.h)
            self.h = None


class DeviceReplayBuffer:
    """The reference ReplayBuffer (scripts/utils/storage.py:6-133) as a ring in HBM owned by libdre and fed by one gather
    kernel per step from the env's double-buffered output slab (include/dre.h: dre_replay_*): no torch indexing, no host
    round trip, no clone of S. `replay_buffer` has the reference's nesting and keys (zero-copy views of the ring).

        rb = DeviceReplayBuffer(env, buffer_size)
        S = env.reset()
        loop:  a = policy(S); S, R, done, info, _ = env.step(a); rb.insert(a)        # S / S_prime come from the two slabs
               Sb, Ab, Spb, Rb, Db = rb.sample_from_buffer(256)

    dtype torch.float64 keeps the reference's storage type; torch.float32 halves the ring (one rounding of each float64)."""

    _OBS_KEYS = (('PT', 'PT_state', 'pt_state'), ('PT', 'MPC_actions', 'mpc_actions'), ('HA', 'robot_node', 'robot_node'),
                 ('HA', 'past_robot_velocities', 'past_robot_vel'), ('HA', 'percent_episode', 'percent_episode'),
                 ('HA', 'spatial_edges', 'spatial_edges'), ('HA', 'human_valid_history', 'human_valid'),
                 ('HA', 'detected_human_num', 'detected_humans'))

    def __init__(self, env, buffer_size, dtype=None):
        import ctypes
        import torch
        from . import _lib
        from .env import _view
        self.env = env
        self.device = env.device
        self.dtype = dtype or torch.float64
        if self.dtype not in (torch.float64, torch.float32):
            raise ValueError("DeviceReplayBuffer stores float64 or float32")
        self._code = 0 if self.dtype == torch.float64 else 1
        self.buffer_size = int(buffer_size)
        self._L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._L.dre_replay_create(env._h, self.buffer_size, self._code, ctypes.byref(h)))
        self._h = h
        self._owner = _ReplayHandle(self._L, h, env._owner)
        view = _lib.ReplayView()
        widths = (ctypes.c_int32 * 8)()
        nbytes = ctypes.c_size_t()
        _lib.check(self._L.dre_replay_storage(self._h, ctypes.byref(view), widths, ctypes.byref(nbytes)))
        self.storage_bytes = int(nbytes.value)
        self._widths = list(widths)
        self._shapes = self._row_shapes(env)
        kind = "f8" if self._code == 0 else "f4"
        mk = lambda ptr, shape: _view(ptr, (self.buffer_size,) + shape, kind, self.device, self._owner)
        self.replay_buffer = {'S': {'PT': {}, 'HA': {}}, 'S_prime': {'PT': {}, 'HA': {}},
                              'rewards': mk(view.rewards, (1,)), 'actions': mk(view.actions, (2,)), 'done': mk(view.done, (1,))}
        for grp, fields in (('S', view.S), ('S_prime', view.S_prime)):
            for top, key, name in self._OBS_KEYS:
                self.replay_buffer[grp][top][key] = mk(getattr(fields, name), self._shapes[name])

    def _row_shapes(self, env):
        shapes = {name: tuple(env.out[name].shape[1:]) for _, _, name in self._OBS_KEYS}
        for (_, _, name), w in zip(self._OBS_KEYS, self._widths):
            n = 1
            for s in shapes[name]:
                n *= s
            assert n == w, (name, shapes[name], w)
        return shapes

    def _stream(self):
        import ctypes
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _count(self):
        import ctypes
        from . import _lib
        v, hd, last = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._L.dre_replay_count(self._h, ctypes.byref(v), ctypes.byref(hd), ctypes.byref(last), self._stream()))
        return int(v.value), int(hd.value), int(last.value)

    @property
    def valid_entries(self):
        """storage.py:13. Reads the device-side counter (synchronises the current stream)."""
        return self._count()[0]

    @property
    def last_inserted(self):
        return self._count()[2]

    def full(self):
        return self.valid_entries == self.buffer_size

    def clear(self):
        from . import _lib
        _lib.check(self._L.dre_replay_clear(self._h, self._stream()))

    def insert(self, actions=None, mask=None, skip_scripted=True):
        """storage.py:54-98 / online_continuous_task.py:189 for all envs of the latest step: (S = previous slab, A,
        S_prime = latest slab, R['R'], done_for_replay). `actions` [E,2] float64 = the policy's action_model (None: the
        executed actions); `mask` [E] bool keeps a subset; skip_scripted leaves the soft-reset steps out. Asynchronous."""
        import ctypes
        import torch
        from . import _lib
        a = None
        if actions is not None:
            a = torch.as_tensor(actions).to(device=self.device, dtype=torch.float64).contiguous()
            if tuple(a.shape) != (self.env.E, 2):
                raise ValueError(f"actions must be [{self.env.E}, 2]")
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if m.numel() != self.env.E:
                raise ValueError(f"mask must have {self.env.E} entries")
        _lib.check(self._L.dre_replay_insert(self._h, ctypes.c_void_p(a.data_ptr()) if a is not None else None,
                                             ctypes.c_void_p(m.data_ptr()) if m is not None else None, int(bool(skip_scripted)),
                                             self._stream()))

    def physical_index(self, chronological):
        valid, head, _ = self._count()
        return chronological if valid < self.buffer_size else (chronological + head) % self.buffer_size

    def sample_from_buffer(self, batch_size, indices_to_sample=None, dtype=None, check_indices=True):
        """storage.py:101-133: (S, actions, S_prime, rewards, done) at random valid rows, or at the given chronological
        indices (0 = oldest). dtype torch.float32 hands the batch out in the networks' precision in the same pass.
        Caller-supplied indices are range-checked against the ring's capacity (IndexError, like the reference's tensor
        indexing; reading the two extrema back is one synchronisation -- check_indices=False skips it)."""
        import ctypes
        import torch
        from . import _lib
        out_dtype = dtype or self.dtype
        code = 0 if out_dtype == torch.float64 else 1
        if indices_to_sample is None:
            valid = self.valid_entries
            if valid < batch_size:      # storage.py:103-106 only warns
                print(f"Not enough entries in buffer to sample from. Only {valid} entries in buffer")
            indices_to_sample = torch.randint(0, max(valid, 1), (batch_size,), device=self.device)
            check_indices = False
        idx = torch.as_tensor(indices_to_sample).to(device=self.device, dtype=torch.int64).contiguous()
        n = int(idx.numel())
        if check_indices and n:     # the reference's tensor indexing raises on a bad index; the gather kernel would read past the ring
            lo, hi = (int(v) for v in torch.stack((idx.min(), idx.max())).tolist())
            if lo < 0 or hi >= self.buffer_size:
                raise IndexError(f"replay index out of range: [{lo}, {hi}] not inside [0, {self.buffer_size})")
        batch = _lib.ReplayView()
        mk = lambda shape: torch.empty((n,) + shape, dtype=out_dtype, device=self.device)
        S, Sp = {'PT': {}, 'HA': {}}, {'PT': {}, 'HA': {}}
        for grp, fields in ((S, batch.S), (Sp, batch.S_prime)):
            for top, key, name in self._OBS_KEYS:
                t = mk(self._shapes[name])
                grp[top][key] = t
                setattr(fields, name, t.data_ptr())
        A, R, D = mk((2,)), mk((1,)), mk((1,))
        batch.actions, batch.rewards, batch.done = A.data_ptr(), R.data_ptr(), D.data_ptr()
        batch.rows, batch.dtype = n, code
        _lib.check(self._L.dre_replay_sample(self._h, ctypes.c_void_p(idx.data_ptr()), n, code, ctypes.byref(batch), self._stream()))
        return S, A, Sp, R, D

    def grab_all_data(self, batch_size):
        """storage.py:135-145 (the OOD model's pass over the whole buffer, scripts/utils/OOD.py:120): every valid row once, in
        chronological order, in batches of `batch_size`."""
        import torch
        out, cur, n = [], 0, self.valid_entries
        while cur < n:
            end = min(cur + batch_size, n)
            out.append(self.sample_from_buffer(0, indices_to_sample=torch.arange(cur, end, device=self.device), check_indices=False))
            cur = end
        return out
