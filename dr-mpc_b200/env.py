"""Batched combined environment: mirrors reference
environment/HA_and_PT/human_avoidance_and_path_tracking_env.py (`HAAndPTEnv(config_master, seed_addition)`,
`.reset() -> S`, `.step(action) -> (S', reward_return, done_return, info, eps_done_info)`) for E envs at once.

Same names and dict keys as the reference; every value carries a leading batch dim E instead of 1 and is a
torch CUDA tensor that views the engine's output slab in HBM (no copy), ready for a batched actor/critic.
"""
import ctypes
import numpy as np
from . import _lib
from .config import EngineConfig
from .paths import continuous_task_paths, pack_paths
from .mpc import MPC_STATUS_DTYPE

DONE_BITS = {'collision': 1 << 0, 'safety_human_raise': 1 << 1, 'actuation_termination': 1 << 2, 'timeout': 1 << 3,
             'goal': 1 << 4, 'deviation_end_of_path': 1 << 5, 'corridor_hit': 1 << 6, 'safety_corridor_raise': 1 << 7,
             'done_HA': 1 << 16, 'done_PT': 1 << 17, 'done': 1 << 18, 'done_for_replay': 1 << 19}
from .parallel import STAT_NAMES


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view engine-owned HBM without copying. torch keeps this object
    alive for as long as the tensor (or any view of it) lives; `owner` in turn keeps the engine handle, and with it the
    memory, alive: a view never dangles."""

    def __init__(self, ptr, shape, typestr, owner=None):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class _Handle:
    """Owns the dre_env handle; destroyed when the env object AND every tensor viewing its memory are gone."""

    def __init__(self, L, h):
        self.L, self.h = L, h

    def __del__(self):
        if self.h is not None and self.L is not None:
            self.L.dre_env_destroy(self.h)
            self.h = None


_TYPESTR = {"f8": "<f8", "f4": "<f4", "i4": "<i4", "u4": "<u4", "u1": "|u1", "i8": "<i8"}


def _view(ptr, shape, kind, device, owner=None):
    import torch
    return torch.as_tensor(_DevArray(ptr, shape, _TYPESTR[kind], owner), device=device)


class BatchedHAAndPTEnv:
    def __init__(self, config_master=None, num_envs=None, seed_addition=0, device="cuda:0", **overrides):
        import torch
        if isinstance(config_master, EngineConfig):
            import dataclasses
            cfg = dataclasses.replace(config_master)    # the caller's object is not modified (num_envs / seed_addition below)
            for k, v in overrides.items():
                setattr(cfg, k, v)
            self.config_master = None
        elif config_master is not None:
            cfg = EngineConfig.from_config_master(config_master, **overrides)
            self.config_master = config_master
        else:
            cfg = EngineConfig(**overrides)
            self.config_master = None
        if num_envs is not None:
            cfg.num_envs = int(num_envs)
        cfg.seed = (cfg.seed + seed_addition) & 0xFFFFFFFFFFFFFFFF
        self.cfg = cfg
        self.E, self.Nh, self.K, self.LB = cfg.num_envs, cfg.num_humans, cfg.horizon, cfg.lookback
        self.device = torch.device(device)
        self.time_step = cfg.time_step
        self._L = _lib.load()
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)   # make sure the primary context exists before the library allocates
        h = ctypes.c_void_p()
        c = cfg.env_config()
        _lib.check(self._L.dre_env_create(ctypes.byref(c), ctypes.byref(h)))
        self._h = h
        self._owner = _Handle(self._L, h)
        self.paths = continuous_task_paths(cfg.circle_radius, cfg.time_step, cfg.node_separation)
        tab, lens = pack_paths(self.paths)
        _lib.check(self._L.dre_env_set_paths(self._h, tab.ctypes.data_as(ctypes.c_void_p), tab.shape[0], tab.shape[2],
                                             lens.ctypes.data_as(ctypes.c_void_p)))
        self._bind_views()
        self._host_slab = None

    # ---- zero-copy views ----
    def _slab_layout(self, o):
        E, Nh, LB, H = self.E, self.Nh, self.LB, self.LB + 1
        nact = 2 * min(self.K, self.cfg.actions_in_input)
        npt = 4 * (self.cfg.lookahead + self.cfg.lookbehind)
        return [("pt_state", o.pt_state, (E, npt), "f8"), ("actions_used", o.actions_used, (E, 2), "f8"),
                ("mpc_actions", o.mpc_actions, (E, nact), "f8"),
                ("robot_node", o.robot_node, (E, 2 * LB), "f8"), ("past_robot_vel", o.past_robot_vel, (E, 2 * LB), "f8"),
                ("percent_episode", o.percent_episode, (E, 1), "f8"),
                ("spatial_edges", o.spatial_edges, (E, Nh, 2 * H), "f8"), ("human_valid", o.human_valid, (E, Nh), "f8"),
                ("detected_humans", o.detected_humans, (E,), "f8"), ("rewards", o.rewards, (E, 3), "f8"),
                ("info", o.info, (E, 8), "f8"), ("done_bits", o.done_bits, (E,), "i4"),
                ("mpc_status", o.mpc_status, (E, 8), "i4"), ("done_flags", o.done_flags, (E, 16), "u1")]

    def _bind_views(self):
        E, Nh, LB, H = self.E, self.Nh, self.LB, self.LB + 1
        d = self.device
        own = self._owner
        # the output slab is double-buffered (step t writes slab t & 1): one set of views per slab
        self._outs = []
        for b in (0, 1):
            o = _lib.StepOut()
            _lib.check(self._L.dre_env_outputs_at(self._h, b, ctypes.byref(o)))
            if b == 0:
                self._out = o
                self.slab_bytes = int(o.slab_bytes)
                self._layout = self._slab_layout(o)
            self._outs.append({name: _view(ptr, shape, kind, d, own) for name, ptr, shape, kind in self._slab_layout(o)})
        s = _lib.EnvStateView()
        _lib.check(self._L.dre_env_state(self._h, ctypes.byref(s)))
        v = lambda ptr, shape, kind: _view(ptr, shape, kind, d, own)
        self.state = {
            "hpos": v(s.hpos, (E, Nh, 2), "f8"), "hvel": v(s.hvel, (E, Nh, 2), "f4"),
            "hgoal": v(s.hgoal, (E, Nh, 2), "f8"), "hstatic": v(s.hstatic, (E, Nh), "u1"),
            "hhist": v(s.hhist, (E, Nh, H, 2), "f8"), "hvalid": v(s.hvalid, (E, Nh), "i4"),
            "rpose": v(s.rpose, (E, 4), "f8"), "rprev": v(s.rprev, (E, 4), "f8"),
            "rhist": v(s.rhist, (E, H, 4), "f8"), "rpastv": v(s.rpastv, (E, LB, 2), "f8"),
            "rvalid": v(s.rvalid, (E,), "i4"), "gtime": v(s.gtime, (E,), "f8"),
            "path_idx": v(s.path_idx, (E,), "i4"), "loc_to_start": v(s.loc_to_start, (E,), "u1"),
            "step_count": v(s.step_count, (E,), "i8"), "sr_state": v(s.sr_state, (E, 4), "i4"),
        }

    @property
    def out(self):
        """Views of the slab holding the LATEST outputs. Views handed out for the previous step stay intact during the next
        one (double buffering): `S_prime, ... = env.step(a); buf.insert(S, a, S_prime, ...); S = S_prime` needs no clone."""
        return self._outs[self._L.dre_env_current_slab(self._h)]

    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _obs(self):
        o = self.out
        return {'PT': {'PT_state': o["pt_state"], 'MPC_actions': o["mpc_actions"]},
                'HA': {'robot_node': o["robot_node"], 'past_robot_velocities': o["past_robot_vel"],
                       'percent_episode': o["percent_episode"], 'spatial_edges': o["spatial_edges"],
                       'detected_human_num': o["detected_humans"], 'human_valid_history': o["human_valid"]}}

    # ---- reference interface ----
    def _mask(self, env_mask):
        import torch
        if env_mask is None:
            return None
        return torch.as_tensor(env_mask).to(device=self.device, dtype=torch.uint8).contiguous()

    def reset(self, hpos_init=None, hgoal_init=None, env_mask=None):
        """HAAndPTEnv.reset (env.py:243-280). Optional [E,Nh,2] tensors replace the device-side spawning; `env_mask` [E]
        (bool / uint8) resets only the selected envs, the others keep their episode."""
        import torch
        conv = lambda t: torch.as_tensor(t).to(device=self.device, dtype=torch.float64).contiguous().reshape(self.E, self.Nh, 2) if t is not None else None
        hp, hg, m = conv(hpos_init), conv(hgoal_init), self._mask(env_mask)   # held until the launches are queued
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(self._L.dre_env_reset_masked(self._h, ptr(m), ptr(hp), ptr(hg), self._stream()))
        return self._obs()

    def step(self, action, update=True, soft_reset=False, model_info=None):
        """HAAndPTEnv.step (env.py:305-416). `action`: [E,2] float64 CUDA tensor of (v, w)."""
        import torch
        a = action.to(device=self.device, dtype=torch.float64).contiguous()
        _lib.check(self._L.dre_env_step(self._h, ctypes.c_void_p(a.data_ptr()), self._stream()))
        return self._package()

    def crowd_step(self, write_obs=False):
        """ORCA-only step: one ORCA pass + the humans' integration, histories and goal resampling, robot frozen
        (one step of HAEnv.warm_start, human_avoidance_env.py:79-125)."""
        _lib.check(self._L.dre_env_crowd_step(self._h, int(write_obs), self._stream()))

    def _package(self):
        import torch
        o = self.out
        bits = o["done_bits"]
        r = o["rewards"]
        reward_return = {'R_PT': r[:, 0], 'R_DO': r[:, 1], 'R': r[:, 2]}
        fl = o["done_flags"].view(torch.bool)       # zero-copy [E,16] views: no kernels are launched to package a step
        done_return = {k: fl[:, i] for i, k in enumerate(('done_PT', 'done_HA', 'done', 'done_for_replay'))}
        info = {'done_bits': bits, 'disturbance_v': o["info"][:, 0], 'disturbance_th': o["info"][:, 1],
                'path_advancement': o["info"][:, 2], 'deviation': o["info"][:, 3], 'dmin': o["info"][:, 4],
                'interp_index': o["info"][:, 5], 'xy_dev': o["info"][:, 6], 'mpc_status': o["mpc_status"],
                # on-device soft reset (EngineConfig.soft_reset_on_device): transitions flagged here ran the scripted
                # recovery action of HAAndPTEnv.soft_reset (env.py:120-240) and do not belong in the replay buffer
                'soft_reset': fl[:, 12], 'soft_reset_stuck': fl[:, 13], 'actions_used': o["actions_used"]}
        eps_done_info = {k: fl[:, 4 + i] for i, k in enumerate(('success', 'collision', 'safety_human_raise', 'timeout',
                                                                 'deviation_end_of_path', 'corridor_hit',
                                                                 'safety_corridor_raise', 'actuation_termination'))}
        return self._obs(), reward_return, done_return, info, eps_done_info

    def observe(self, solve_mpc=False, env_mask=None):
        """Regenerate S from the current state (HAEnv.soft_reset / PTEnv.soft_reset, env.py:141-148)."""
        m = self._mask(env_mask)
        _lib.check(self._L.dre_env_observe_masked(self._h, ctypes.c_void_p(m.data_ptr()) if m is not None else None, int(solve_mpc), self._stream()))
        return self._obs()

    def switch_path(self, env_mask=None):
        """The path switch of HAAndPTEnv.soft_reset (env.py:130-141): next path, localize_to_start, MPC.reset(), for the
        selected envs. Follow with observe(solve_mpc=True, env_mask=...) like PT_env.soft_reset / HA_env.soft_reset."""
        m = self._mask(env_mask)
        _lib.check(self._L.dre_env_switch_path(self._h, ctypes.c_void_p(m.data_ptr()) if m is not None else None, self._stream()))

    def soft_reset_decide(self, commit=True):
        """One decision of HAAndPTEnv.soft_reset's loop (env.py:128-238) per env from the latest done word:
        -> (scripted [E] bool, actions [E,2], stuck [E] bool). For callers that drive the loop themselves."""
        import torch
        sc = torch.empty(self.E, dtype=torch.uint8, device=self.device)
        a = torch.empty((self.E, 2), dtype=torch.float64, device=self.device)
        st = torch.empty(self.E, dtype=torch.uint8, device=self.device)
        _lib.check(self._L.dre_env_soft_reset_decide(self._h, int(commit), ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(a.data_ptr()),
                                                     ctypes.c_void_p(st.data_ptr()), self._stream()))
        return sc.bool(), a, st.bool()

    # ---- host-buffer entry point (end-to-end drop-in for a CPU-side caller) ----
    OBS_FIELDS = ("pt_state", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid", "detected_humans")
    CONTROL_FIELDS = ("actions_used", "mpc_actions", "rewards", "info", "done_bits", "mpc_status", "done_flags")

    def _make_host_views(self):
        import torch
        if self._host_slab is None:
            self._host_slab = torch.empty(self.slab_bytes, dtype=torch.uint8).pin_memory()
        base = self._host_slab.numpy()
        mode = getattr(self, "_host_mode", 0)
        self._host_views = {}
        self.host_bytes_per_step = 0
        for name, ptr, shape, kind in self._layout:
            if mode == 2 and name in self.OBS_FIELDS:
                continue
            off = int(ptr) - int(self._out.slab)
            n = int(np.prod(shape))
            dt = np.dtype("<f4") if (mode == 1 and name in self.OBS_FIELDS) else np.dtype(_TYPESTR[kind])
            self._host_views[name] = base[off:off + n * dt.itemsize].view(dt).reshape(shape)
            self.host_bytes_per_step += n * dt.itemsize

    def configure_host_outputs(self, out_dtype="float64", outputs="all"):
        """What step_host brings back each step: out_dtype 'float32' = the observation fields rounded once to float32 on the
        device (half the bytes); outputs 'control' = only actions / rewards / info / done words / MPC status (the observations
        stay in HBM for a GPU-side consumer). Default: the whole float64 slab, like the reference's loop receives."""
        mode = 2 if outputs == "control" else (1 if out_dtype in ("float32", "f32") else 0)
        _lib.check(self._L.dre_env_set_host_outputs(self._h, mode))
        self._host_mode = mode
        self._host_views = None

    def d2h_ceiling_ms(self, reps=20):
        """Milliseconds of one plain pinned D2H copy of the slab (the floor of a step_host that returns everything)."""
        if self._host_slab is None:
            self._make_host_views()
        ms = _lib.c_d()
        _lib.check(self._L.dre_env_d2h_probe(self._h, ctypes.c_void_p(self._host_slab.data_ptr()), int(reps), ctypes.byref(ms)))
        return float(ms.value)

    def step_host(self, actions_np):
        """actions: numpy [E,2] float64 (pinned for best speed) -> dict of numpy views into one pinned host slab."""
        if self._host_slab is None or getattr(self, "_host_views", None) is None:
            self._make_host_views()
        a = np.ascontiguousarray(actions_np, dtype=np.float64)
        _lib.check(self._L.dre_env_step_host(self._h, a.ctypes.data_as(ctypes.c_void_p),
                                             ctypes.c_void_p(self._host_slab.data_ptr())))
        return self._host_views

    CVMM_ACTION_SET = ((0.05, -0.8), (0.15, -0.8), (0.6, -0.8), (0.05, -0.2), (0.15, -0.2), (0.6, -0.2), (0.15, 0.0), (0.6, 0.0),
                       (0.05, 0.2), (0.15, 0.2), (0.6, 0.2), (0.05, 0.8), (0.15, 0.8), (0.6, 0.8))   # config_general.py:88-91

    def cvmm_diverse_safety_pipeline(self, action_execute, num_steps=8, backup_actions=None, want_candidates=False):
        """HAAndPTEnv.cvmm_diverse_safety_pipeline (env.py:422-473) for every env: `action_execute` [E,2] float64 CUDA
        tensor -> (safe_actions [E,2], info). info['original_safe'] / ['num_safe'] / ['chosen'] (-1 = original kept) /
        ['random'] (no safe action: random back-up, env.py:459-464) are int32 [E] tensors; the env is not modified."""
        import torch
        a = action_execute.to(device=self.device, dtype=torch.float64).contiguous()
        bk = torch.as_tensor(np.asarray(self.CVMM_ACTION_SET if backup_actions is None else backup_actions, dtype=np.float64),
                             device=self.device).contiguous()
        out = torch.empty_like(a)
        info = torch.empty((self.E, 4), dtype=torch.int32, device=self.device)
        cand = torch.empty((self.E, bk.shape[0] + 1), dtype=torch.uint8, device=self.device) if want_candidates else None
        _lib.check(self._L.dre_env_cvmm_filter(self._h, ctypes.c_void_p(a.data_ptr()), int(num_steps), ctypes.c_void_p(bk.data_ptr()),
                                               int(bk.shape[0]), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(info.data_ptr()),
                                               ctypes.c_void_p(cand.data_ptr()) if cand is not None else None, self._stream()))
        d = {'original_safe': info[:, 0], 'num_safe': info[:, 1], 'chosen': info[:, 2], 'random': info[:, 3]}
        if cand is not None:
            d['candidate_safe'] = cand
        return out, d

    def cvmm_safety_check(self, action_execute, num_steps=8):
        """HAAndPTEnv.cvmm_safety_check (env.py:475-513) for every env: [E,2] -> (safe [E] bool, FDE_to_path [E])."""
        import torch
        a = action_execute.to(device=self.device, dtype=torch.float64).contiguous()
        safe = torch.empty(self.E, dtype=torch.uint8, device=self.device)
        dist = torch.empty(self.E, dtype=torch.float64, device=self.device)
        _lib.check(self._L.dre_env_cvmm_check(self._h, ctypes.c_void_p(a.data_ptr()), int(num_steps), ctypes.c_void_p(safe.data_ptr()),
                                              ctypes.c_void_p(dist.data_ptr()), self._stream()))
        return safe.bool(), dist

    def set_seed(self, seed):
        """Philox key of the spawning / goal resampling from the next reset() on."""
        self.cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        _lib.check(self._L.dre_env_set_seed(self._h, ctypes.c_uint64(self.cfg.seed)))

    def invalidate_orca_cache(self):
        """orca_dedup: call after writing human / robot state through `self.state` (the cached look-ahead is then stale)."""
        _lib.check(self._L.dre_env_invalidate_orca_cache(self._h, self._stream()))

    def set_host_chunks(self, chunks):
        """Number of env chunks step_host pipelines the crowd half and its D2H copies in (1 = no pipelining)."""
        _lib.check(self._L.dre_env_set_host_chunks(self._h, int(chunks)))

    def stats_comm_init(self, rank, world_size, broadcast):
        """Create the engine's own NCCL communicator for the episode statistics. `broadcast(bytes_or_None) -> bytes` must
        deliver rank 0's 128-byte id to every rank (e.g. through torch.distributed.broadcast_object_list)."""
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(self._L.dre_stats_comm_id(buf))
        ident = broadcast(bytes(buf.raw) if rank == 0 else None)
        _lib.check(self._L.dre_env_stats_comm_init(self._h, ctypes.c_char_p(ident), int(rank), int(world_size)))

    def stats_reduce(self, reset=False):
        """dre_stats_reduce: the statistics of ALL ranks (one ncclAllReduce(sum) pair over NVLink) -> (int64[16], float64[4])."""
        import torch
        c = torch.zeros(16, dtype=torch.int64, device=self.device)
        s = torch.zeros(4, dtype=torch.float64, device=self.device)
        _lib.check(self._L.dre_stats_reduce(self._h, None, ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(s.data_ptr()), int(reset), self._stream()))
        return c, s

    def stats(self, reset=False):
        """Episode statistics accumulated on device -> (int64[16], float64[4]) CUDA tensors (feed to all_reduce)."""
        import torch
        c = torch.zeros(16, dtype=torch.int64, device=self.device)
        s = torch.zeros(4, dtype=torch.float64, device=self.device)
        _lib.check(self._L.dre_env_stats(self._h, ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(s.data_ptr()), int(reset),
                                         self._stream()))
        return c, s

    def set_profiling(self, on=True):
        _lib.check(self._L.dre_env_set_profiling(self._h, int(on)))

    def kernel_ms(self, reset=False):
        """-> ({'ha_step','pt_step','mpc_solve','mpc_stats'} summed ms, profiled steps) since the last reset."""
        ms = (ctypes.c_double * 4)()
        n = ctypes.c_int64()
        _lib.check(self._L.dre_env_kernel_ms(self._h, ms, ctypes.byref(n), int(reset)))
        return dict(zip(("ha_step", "pt_step", "mpc_solve", "mpc_stats"), list(ms))), int(n.value)

    @property
    def MPC(self):
        return self

    def mpc_get_warm(self):
        T = np.empty((self.E, self.K, 4)); u = np.empty((self.E, self.K, 2)); it0 = np.empty(self.E, dtype=np.uint8)
        m = self._L.dre_env_mpc(self._h)
        _lib.check(self._L.dre_mpc_get_warm(m, T.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                                            it0.ctypes.data_as(ctypes.c_void_p)))
        return T, u, it0

    def mpc_set_warm(self, T=None, u=None, it0=None):
        m = self._L.dre_env_mpc(self._h)
        f = lambda a, dt: np.ascontiguousarray(a, dtype=dt).ctypes.data_as(ctypes.c_void_p) if a is not None else None
        T_, u_, i_ = (np.ascontiguousarray(T, dtype=np.float64) if T is not None else None,
                      np.ascontiguousarray(u, dtype=np.float64) if u is not None else None,
                      np.ascontiguousarray(it0, dtype=np.uint8) if it0 is not None else None)
        _lib.check(self._L.dre_mpc_set_warm(m, T_.ctypes.data_as(ctypes.c_void_p) if T_ is not None else None,
                                            u_.ctypes.data_as(ctypes.c_void_p) if u_ is not None else None,
                                            i_.ctypes.data_as(ctypes.c_void_p) if i_ is not None else None))
