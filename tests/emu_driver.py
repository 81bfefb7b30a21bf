"""TEST INFRASTRUCTURE: a numpy-only driver of the product's C ABI (include/dre.h) for the EMULATED build of the library
(tests/host_harness/gen_emu.py -> libdre_emu.so, where "device" memory is host memory). It does what dr_mpc_b200/env.py does --
the same structs, the same calls in the same order -- with numpy views where the product hands out zero-copy torch CUDA tensors,
so that the `-m "not gpu"` suite can put the real entry points and kernels through the reference's golden fixtures."""
import ctypes
import importlib.util
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None


def lib():
    """ctypes handle of libdre_emu.so with the argtypes of dr_mpc_b200._lib (built on first use)."""
    global _L
    if _L is None:
        spec = importlib.util.spec_from_file_location("dre_gen_emu", os.path.join(REPO, "tests", "host_harness", "gen_emu.py"))
        gen = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gen)
        so = gen.build(sanitize=os.environ.get("DRE_EMU_SANITIZE") or None)     # tools/emu_memcheck.sh: the same suite under ASan
        import dr_mpc_b200._lib as LL
        saved = (LL.LIB_PATH, LL._lib)
        try:                                   # borrow the product's own prototype table for the emulated binary
            LL.LIB_PATH, LL._lib = so, None
            _L = LL.load()
        finally:
            LL.LIB_PATH, LL._lib = saved
    return _L


def check(rc):
    import dr_mpc_b200._lib as LL
    if rc != 0:
        raise LL.DreError(f"dre call failed: {rc} ({lib().dre_strerror(rc).decode()})")


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def view(address, shape, dtype):
    """numpy view of `device` memory of the emulated library"""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if len(shape) else 1
    buf = (ctypes.c_char * (n * dt.itemsize)).from_address(int(address))
    return np.frombuffer(buf, dtype=dt).reshape(shape)


STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime", "path_idx",
         "loc_to_start")


class EmuEnv:
    """BatchedHAAndPTEnv's calls (dr_mpc_b200/env.py) on the emulated library."""

    def __init__(self, num_envs, **cfg):
        import dr_mpc_b200._lib as LL
        from dr_mpc_b200.config import EngineConfig
        from dr_mpc_b200.paths import continuous_task_paths, pack_paths
        self.LL, self.L = LL, lib()
        self.cfg = EngineConfig(num_envs=num_envs, **cfg)
        c = self.cfg.env_config()
        self.h = ctypes.c_void_p()
        check(self.L.dre_env_create(ctypes.byref(c), ctypes.byref(self.h)))
        self.paths = continuous_task_paths(self.cfg.circle_radius, self.cfg.time_step, self.cfg.node_separation)
        tab, lens = pack_paths(self.paths)
        check(self.L.dre_env_set_paths(self.h, ptr(tab), tab.shape[0], tab.shape[2], ptr(lens)))
        E, Nh, LB = num_envs, self.cfg.num_humans, self.cfg.lookback
        H = LB + 1
        self.E, self.Nh, self.LB = E, Nh, LB
        s = LL.EnvStateView()
        check(self.L.dre_env_state(self.h, ctypes.byref(s)))
        self.state = {
            "hpos": view(s.hpos, (E, Nh, 2), "f8"), "hvel": view(s.hvel, (E, Nh, 2), "f4"), "hgoal": view(s.hgoal, (E, Nh, 2), "f8"),
            "hstatic": view(s.hstatic, (E, Nh), "u1"), "hhist": view(s.hhist, (E, Nh, H, 2), "f8"), "hvalid": view(s.hvalid, (E, Nh), "i4"),
            "rpose": view(s.rpose, (E, 4), "f8"), "rprev": view(s.rprev, (E, 4), "f8"), "rhist": view(s.rhist, (E, H, 4), "f8"),
            "rpastv": view(s.rpastv, (E, LB, 2), "f8"), "rvalid": view(s.rvalid, (E,), "i4"), "gtime": view(s.gtime, (E,), "f8"),
            "path_idx": view(s.path_idx, (E,), "i4"), "loc_to_start": view(s.loc_to_start, (E,), "u1"),
            "step_count": view(s.step_count, (E,), "i8"), "sr_state": view(s.sr_state, (E, 4), "i4")}
        self._outs = [self._slab_views(b) for b in (0, 1)]

    def _slab_views(self, which):
        o = self.LL.StepOut()
        check(self.L.dre_env_outputs_at(self.h, which, ctypes.byref(o)))
        E, Nh, LB = self.E, self.Nh, self.LB
        H = LB + 1
        nact = 2 * min(self.cfg.horizon, self.cfg.actions_in_input)
        npt = 4 * (self.cfg.lookahead + self.cfg.lookbehind)
        self.slab_bytes = int(o.slab_bytes)
        self._slab0 = getattr(self, "_slab0", int(o.slab))
        spec = [("pt_state", (E, npt), "f8"), ("actions_used", (E, 2), "f8"), ("mpc_actions", (E, nact), "f8"), ("robot_node", (E, 2 * LB), "f8"),
                ("past_robot_vel", (E, 2 * LB), "f8"), ("percent_episode", (E, 1), "f8"), ("spatial_edges", (E, Nh, 2 * H), "f8"),
                ("human_valid", (E, Nh), "f8"), ("detected_humans", (E,), "f8"), ("rewards", (E, 3), "f8"), ("info", (E, 8), "f8"),
                ("done_bits", (E,), "u4"), ("mpc_status", (E, 8), "i4"), ("done_flags", (E, 16), "u1")]
        self._offsets = {n: int(getattr(o, n)) - int(o.slab) for n, _, _ in spec}
        self._spec = spec
        return {n: view(getattr(o, n), shp, dt) for n, shp, dt in spec}

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dre_env_destroy(self.h)
            self.h = None

    @property
    def out(self):
        return self._outs[self.L.dre_env_current_slab(self.h)]

    def upload(self, g, tag, prefix, sel=slice(None)):
        """teacher forcing: the reference's recorded state into the engine's state arrays (+ the MPC warm start)"""
        for name in STATE:
            key = f"{tag}_{prefix}_{name}"
            if name == "hstatic" and key not in g.files:
                self.state[name][...] = 0
                continue
            self.state[name][...] = np.asarray(g[key][sel]).reshape(self.state[name].shape)
        self.mpc_set_warm(g[f"{tag}_{prefix}_warm_T"][sel], g[f"{tag}_{prefix}_warm_u"][sel], g[f"{tag}_{prefix}_it0"][sel].astype(np.uint8))

    def mpc_set_warm(self, T, u, it0):
        m = self.L.dre_env_mpc(self.h)
        T, u, it0 = np.ascontiguousarray(T, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64), np.ascontiguousarray(it0, dtype=np.uint8)
        check(self.L.dre_mpc_set_warm(m, ptr(T), ptr(u), ptr(it0)))

    def reset(self, hpos_init=None, hgoal_init=None, env_mask=None):
        hp = np.ascontiguousarray(hpos_init, dtype=np.float64) if hpos_init is not None else None
        hg = np.ascontiguousarray(hgoal_init, dtype=np.float64) if hgoal_init is not None else None
        m = np.ascontiguousarray(env_mask, dtype=np.uint8) if env_mask is not None else None
        check(self.L.dre_env_reset_masked(self.h, ptr(m), ptr(hp), ptr(hg), None))
        return self.out

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.E, 2)
        check(self.L.dre_env_step(self.h, ptr(a), None))
        return self.out

    def step_host(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.E, 2)
        if not hasattr(self, "_host_slab"):
            self._host_slab = np.zeros(self.slab_bytes, dtype=np.uint8)
        check(self.L.dre_env_step_host(self.h, ptr(a), ptr(self._host_slab)))
        out = {}
        for n, shp, dt in self._spec:
            cnt = int(np.prod(shp)) * np.dtype(dt).itemsize
            out[n] = self._host_slab[self._offsets[n]:self._offsets[n] + cnt].view(dt).reshape(shp)
        return out

    def set_host_chunks(self, n):
        check(self.L.dre_env_set_host_chunks(self.h, int(n)))

    def flags(self):
        """done_return / info / eps_done_info flags of the latest step as in env.py:_package"""
        fl = self.out["done_flags"].astype(bool)
        names = ("done_PT", "done_HA", "done", "done_for_replay", "success", "collision", "safety_human_raise", "timeout",
                 "deviation_end_of_path", "corridor_hit", "safety_corridor_raise", "actuation_termination", "soft_reset", "soft_reset_stuck")
        return {n: fl[:, i] for i, n in enumerate(names)}

    def cvmm(self, actions, num_steps, action_set):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        bk = np.ascontiguousarray(action_set, dtype=np.float64)
        out = np.zeros_like(a)
        info = np.zeros((self.E, 4), dtype=np.int32)
        cand = np.zeros((self.E, bk.shape[0] + 1), dtype=np.uint8)
        check(self.L.dre_env_cvmm_filter(self.h, ptr(a), int(num_steps), ptr(bk), int(bk.shape[0]), ptr(out), ptr(info), ptr(cand), None))
        return out, {"original_safe": info[:, 0], "num_safe": info[:, 1], "chosen": info[:, 2], "random": info[:, 3], "candidate_safe": cand}

    def mpc_get_warm(self):
        m = self.L.dre_env_mpc(self.h)
        K = self.cfg.horizon
        T, u, it0 = np.empty((self.E, K, 4)), np.empty((self.E, K, 2)), np.empty(self.E, dtype=np.uint8)
        check(self.L.dre_mpc_get_warm(m, ptr(T), ptr(u), ptr(it0)))
        return T, u, it0

    def invalidate_orca_cache(self):
        check(self.L.dre_env_invalidate_orca_cache(self.h, None))

    def observe(self, solve_mpc=False, env_mask=None):
        m = np.ascontiguousarray(env_mask, dtype=np.uint8) if env_mask is not None else None
        check(self.L.dre_env_observe_masked(self.h, ptr(m), int(solve_mpc), None))
        return self.out

    def switch_path(self, env_mask=None):
        m = np.ascontiguousarray(env_mask, dtype=np.uint8) if env_mask is not None else None
        check(self.L.dre_env_switch_path(self.h, ptr(m), None))

    def soft_reset_decide(self, commit=True):
        sc, a, st = np.zeros(self.E, np.uint8), np.zeros((self.E, 2)), np.zeros(self.E, np.uint8)
        check(self.L.dre_env_soft_reset_decide(self.h, int(commit), ptr(sc), ptr(a), ptr(st), None))
        return sc.astype(bool), a, st.astype(bool)

    def cvmm_check(self, actions, num_steps=8):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        safe, dist = np.zeros(self.E, np.uint8), np.zeros(self.E)
        check(self.L.dre_env_cvmm_check(self.h, ptr(a), int(num_steps), ptr(safe), ptr(dist), None))
        return safe.astype(bool), dist

    def crowd_step(self, write_obs=False):
        check(self.L.dre_env_crowd_step(self.h, int(write_obs), None))

    def set_host_outputs(self, mode):
        check(self.L.dre_env_set_host_outputs(self.h, int(mode)))
        self._host_mode = int(mode)

    def step_host_mode(self, actions):
        """step_host with the configured output mode: {name: array} of what travelled (float32 observations in mode 1, control
        fields only in mode 2), decoded like BatchedHAAndPTEnv._make_host_views"""
        mode = getattr(self, "_host_mode", 0)
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.E, 2)
        if not hasattr(self, "_host_slab"):
            self._host_slab = np.zeros(self.slab_bytes, dtype=np.uint8)
        check(self.L.dre_env_step_host(self.h, ptr(a), ptr(self._host_slab)))
        obs = ("pt_state", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid", "detected_humans")
        out = {}
        for n, shp, dt in self._spec:
            if mode == 2 and n in obs:
                continue
            d = np.dtype("f4") if (mode == 1 and n in obs) else np.dtype(dt)
            cnt = int(np.prod(shp)) * d.itemsize
            out[n] = self._host_slab[self._offsets[n]:self._offsets[n] + cnt].view(d).reshape(shp).copy()
        return out

    def stats(self, reset=False):
        c, s = np.zeros(16, np.int64), np.zeros(4)
        check(self.L.dre_env_stats(self.h, ptr(c), ptr(s), int(reset), None))
        return c, s


OBS_KEYS = ("pt_state", "mpc_actions", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid", "detected_humans")


class EmuReplay:
    """DeviceReplayBuffer's calls (dr_mpc_b200/replay.py) on the emulated library: dre_replay_create / insert / sample / count."""

    def __init__(self, env, capacity, dtype=np.float64):
        self.env, self.cap, self.dtype = env, int(capacity), np.dtype(dtype)
        self.L, self.LL = env.L, env.LL
        code = 0 if self.dtype == np.float64 else 1
        self.h = ctypes.c_void_p()
        check(self.L.dre_replay_create(env.h, self.cap, code, ctypes.byref(self.h)))
        v = self.LL.ReplayView()
        widths = (ctypes.c_int32 * 8)()
        nbytes = ctypes.c_size_t()
        check(self.L.dre_replay_storage(self.h, ctypes.byref(v), widths, ctypes.byref(nbytes)))
        self.widths = list(widths)
        self.ring = {"S": {k: view(getattr(v.S, k), (self.cap, w), self.dtype) for k, w in zip(OBS_KEYS, self.widths)},
                     "S_prime": {k: view(getattr(v.S_prime, k), (self.cap, w), self.dtype) for k, w in zip(OBS_KEYS, self.widths)},
                     "rewards": view(v.rewards, (self.cap, 1), self.dtype), "actions": view(v.actions, (self.cap, 2), self.dtype),
                     "done": view(v.done, (self.cap, 1), self.dtype)}

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dre_replay_destroy(self.h)
            self.h = None

    def insert(self, actions=None, mask=None, skip_scripted=True):
        a = np.ascontiguousarray(actions, dtype=np.float64) if actions is not None else None
        m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
        check(self.L.dre_replay_insert(self.h, ptr(a), ptr(m), int(bool(skip_scripted)), None))

    def count(self):
        v, hd, last = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(self.L.dre_replay_count(self.h, ctypes.byref(v), ctypes.byref(hd), ctypes.byref(last), None))
        return int(v.value), int(hd.value), int(last.value)

    def sample(self, indices, dtype=None):
        dt = np.dtype(dtype or self.dtype)
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        n = idx.size
        b = self.LL.ReplayView()
        out = {"S": {}, "S_prime": {}}
        for grp, fields in (("S", b.S), ("S_prime", b.S_prime)):
            for k, w in zip(OBS_KEYS, self.widths):
                out[grp][k] = np.zeros((n, w), dtype=dt)
                setattr(fields, k, out[grp][k].ctypes.data)
        out["actions"], out["rewards"], out["done"] = np.zeros((n, 2), dt), np.zeros((n, 1), dt), np.zeros((n, 1), dt)
        b.actions, b.rewards, b.done = out["actions"].ctypes.data, out["rewards"].ctypes.data, out["done"].ctypes.data
        b.rows, b.dtype = n, 0 if dt == np.float64 else 1
        check(self.L.dre_replay_sample(self.h, ptr(idx), n, b.dtype, ctypes.byref(b), None))
        return out
