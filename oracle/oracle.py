"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package must never import it.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("rvo2_ref.c", "mpc_ref.c", "env_ref.c") if os.path.exists(os.path.join(_HERE, f))]
    if (not force) and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


class OrcaStats(ctypes.Structure):
    _fields_ = [("n_lines", ctypes.c_int32), ("lp2_fail", ctypes.c_int32), ("n_lp1", ctypes.c_int32),
                ("lp1_work", ctypes.c_int32), ("branch_mask", ctypes.c_uint32), ("n_lp3_proj", ctypes.c_int32)]


ORCA_STATS_DTYPE = np.dtype([("n_lines", "i4"), ("lp2_fail", "i4"), ("n_lp1", "i4"), ("lp1_work", "i4"),
                             ("branch_mask", "u4"), ("n_lp3_proj", "i4")])


class MpcParams(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int32), ("dt", ctypes.c_double), ("vmax", ctypes.c_double), ("wmax", ctypes.c_double),
                ("node_inc", ctypes.c_double), ("max_iterations", ctypes.c_int32), ("abs_change_tol", ctypes.c_double),
                ("rel_change_tol", ctypes.c_double), ("pose_jac_exact", ctypes.c_int32), ("max_retries", ctypes.c_int32)]


MPC_STATS_DTYPE = np.dtype([("iterations", "i4"), ("retries", "i4"), ("lm_solves", "i4"), ("cost_evals", "i4"),
                            ("term", "i4"), ("pad", "i4"), ("final_cost", "f8")])


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def orca_solve(self8, others5, neighbor_dist=10.0, max_neighbors=None, time_horizon=5.0, time_step=0.25):
    """Single ORCA solve (fp32). Returns (v[2] float32, stats dict)."""
    self8 = np.ascontiguousarray(self8, dtype=np.float32)
    others5 = np.ascontiguousarray(others5, dtype=np.float32).reshape(-1, 5)
    n = others5.shape[0]
    out = np.zeros(2, dtype=np.float32)
    st = OrcaStats()
    f = lib().orca_ref_solve
    f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                  ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    f(n, _p(self8), _p(others5), neighbor_dist, n if max_neighbors is None else max_neighbors, time_horizon,
      time_step, _p(out), ctypes.byref(st))
    return out, {k: getattr(st, k) for k, _ in OrcaStats._fields_}


def orca_crowd_pass(hpos, hvel, goal, rpos, is_static=None, robot_visible=True, neighbor_dist=10.0,
                    time_horizon=5.0, time_step=0.25, radius=0.485, vmax=1.0, nthreads=0, want_stats=False):
    """Batched per-human ORCA pass with the reference's semantics. Shapes [E,Nh,2] / [E,2]."""
    hpos = np.ascontiguousarray(hpos, dtype=np.float64)
    hvel = np.ascontiguousarray(hvel, dtype=np.float32)
    goal = np.ascontiguousarray(goal, dtype=np.float64)
    rpos = np.ascontiguousarray(rpos, dtype=np.float64)
    E, Nh, _ = hpos.shape
    vnew = np.zeros((E, Nh, 2), dtype=np.float32)
    stats = np.zeros((E, Nh), dtype=ORCA_STATS_DTYPE) if want_stats else None
    st8 = np.ascontiguousarray(is_static, dtype=np.uint8) if is_static is not None else None
    f = lib().orca_ref_crowd_pass
    f.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_float] * 5 + \
                 [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    f(E, Nh, _p(hpos), _p(hvel), _p(goal), _p(rpos), _p(st8), int(robot_visible), neighbor_dist, time_horizon,
      time_step, radius, vmax, _p(vnew), _p(stats), nthreads)
    return (vnew, stats) if want_stats else vnew


def mpc_params(K=4, dt=0.25, vmax=1.0, wmax=1.0, node_separation=0.2, max_iterations=1500, abs_change_tol=1e-4,
               rel_change_tol=1e-4, pose_jac_exact=False, max_retries=20):
    return MpcParams(K, dt, vmax, wmax, (vmax * 0.99) * dt / node_separation, max_iterations, abs_change_tol,
                     rel_change_tol, int(pose_jac_exact), max_retries)


class MpcBatch:
    """Persistent warm-start state of N independent reference MPC instances (mpc.py:51-56,192-205)."""

    def __init__(self, N, params, paths, lens):
        self.N, self.P = N, params
        self.paths = np.ascontiguousarray(paths, dtype=np.float64)   # [P,3,Lstride]
        self.lens = np.ascontiguousarray(lens, dtype=np.int32)
        K = params.K
        self.warm_T = np.zeros((N, K, 4))
        self.warm_u = np.zeros((N, K, 2))
        self.iteration0 = np.ones(N, dtype=np.int32)

    def reset(self, mask=None):
        if mask is None:
            self.iteration0[:] = 1
        else:
            self.iteration0[np.asarray(mask, dtype=bool)] = 1

    def act(self, path_id, interp_idx, pose, nthreads=0):
        N, K = self.N, self.P.K
        path_id = np.ascontiguousarray(path_id, dtype=np.int32)
        interp_idx = np.ascontiguousarray(interp_idx, dtype=np.float64)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        out = np.zeros((N, 2 * K))
        stats = np.zeros(N, dtype=MPC_STATS_DTYPE)
        f = lib().mpc_ref_act_batch
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 9 + [ctypes.c_int]
        f(ctypes.byref(self.P), N, _p(self.paths), self.paths.shape[2], _p(self.lens), _p(path_id), _p(interp_idx),
          _p(pose), _p(self.warm_T), _p(self.warm_u), _p(self.iteration0), _p(out), _p(stats), nthreads)
        return out, stats


# ------------------------------------------------------------------------------------------------------------
# flat env oracle (oracle/env_ref.c)
# ------------------------------------------------------------------------------------------------------------
class EnvCfg(ctypes.Structure):
    _fields_ = [("E", ctypes.c_int32), ("Nh", ctypes.c_int32), ("LB", ctypes.c_int32),
                ("neighbor_dist", ctypes.c_float), ("time_horizon", ctypes.c_float), ("orca_radius", ctypes.c_float),
                ("max_speed_self", ctypes.c_float), ("max_neighbors", ctypes.c_int32), ("robot_visible", ctypes.c_int32),
                ("dt", ctypes.c_double), ("time_limit", ctypes.c_double), ("human_radius", ctypes.c_double),
                ("robot_radius", ctypes.c_double), ("circle_radius", ctypes.c_double), ("circle_radius_humans", ctypes.c_double),
                ("end_goal_changing", ctypes.c_int32)] + \
               [(n, ctypes.c_double) for n in ("collision_penalty", "safety_dist", "safety_penalty", "dist_radius", "dist_factor_v",
                                               "dist_factor_th", "act_min_vel", "act_penalty", "pt_scaling", "goal_radius",
                                               "goal_angle", "goal_reward", "path_adv_scaling", "dev_xy", "dev_th", "dev_scale",
                                               "corridor_penalty", "corridor_max_dev", "eop_penalty", "safety_corridor_penalty",
                                               "safety_corridor_max_dev")] + \
               [("lookahead", ctypes.c_int32), ("lookbehind", ctypes.c_int32), ("mpc_actions_in_input", ctypes.c_int32),
                ("auto_path_switch", ctypes.c_int32), ("seed", ctypes.c_uint64), ("env_id_offset", ctypes.c_int64),
                ("mpc", MpcParams), ("paths", ctypes.c_void_p), ("lens", ctypes.c_void_p), ("num_paths", ctypes.c_int32),
                ("path_stride", ctypes.c_int32), ("soft_reset", ctypes.c_int32), ("sensor_restriction", ctypes.c_int32),
                ("num_static", ctypes.c_int32), ("sensor_range", ctypes.c_double), ("static_pos", ctypes.c_double * 8)]


_ENV_ARRAYS = ["hpos", "hgoal", "hhist", "hvel", "hstatic", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
               "path_idx", "loc_to_start", "step_count", "reset_epoch", "warm_T", "warm_u", "iteration0",
               "pt_state", "mpc_actions", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid",
               "detected_humans", "rewards", "info", "done_bits", "mpc_status", "sr_state", "actions_used", "soft_flags"]


class EnvArrays(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _ENV_ARRAYS]


class EnvRef:
    """E reference environments held in numpy arrays with the engine's state / output layout."""

    def __init__(self, E, Nh, paths, lens, K=4, lookback=6, seed=0x00D24D9C, env_id_offset=0, auto_path_switch=True,
                 end_goal_changing=True, neighbor_dist=10.0, actions_in_input=4, soft_reset=False, sensor_restriction=False,
                 sensor_range=4.0, static_humans=(), **over):
        self.E, self.Nh, self.K, self.LB = E, Nh, K, lookback
        H, LB = lookback + 1, lookback
        self.paths = np.ascontiguousarray(paths, dtype=np.float64)
        self.lens = np.ascontiguousarray(lens, dtype=np.int32)
        c = EnvCfg()
        c.E, c.Nh, c.LB = E, Nh, lookback
        c.neighbor_dist, c.time_horizon, c.orca_radius, c.max_speed_self = neighbor_dist, 5.0, 0.3 + 0.01 + 0.175, 1.0
        c.max_neighbors, c.robot_visible = 0, 1
        c.dt, c.time_limit, c.human_radius, c.robot_radius, c.circle_radius, c.circle_radius_humans = 0.25, 50.0, 0.3, 0.3, 4.0, 5.0
        c.end_goal_changing = int(end_goal_changing)
        c.collision_penalty, c.safety_dist, c.safety_penalty = -15.0, 0.22, -15.0
        c.dist_radius, c.dist_factor_v, c.dist_factor_th = 1.5, 0.8 * 7, 0.5 * 7
        c.act_min_vel, c.act_penalty = 0.025, -20.0
        c.pt_scaling, c.goal_radius, c.goal_angle, c.goal_reward = 0.5, 0.3, 0.5, 0.0
        c.path_adv_scaling, c.dev_xy, c.dev_th, c.dev_scale = 2.0, 1.0, 0.05, 0.9
        c.corridor_penalty, c.corridor_max_dev, c.eop_penalty = -20.0, 1.6, -10.0
        c.safety_corridor_penalty, c.safety_corridor_max_dev = -20.0, 1.4
        c.lookahead, c.lookbehind, c.mpc_actions_in_input, c.auto_path_switch = 20, 5, actions_in_input, int(auto_path_switch)
        c.seed, c.env_id_offset = seed, env_id_offset
        c.mpc = mpc_params(K=K)
        c.soft_reset, c.sensor_restriction, c.sensor_range = int(soft_reset), int(sensor_restriction), float(sensor_range)
        c.num_static = len(static_humans)
        for i, xy in enumerate(static_humans):
            c.static_pos[2 * i], c.static_pos[2 * i + 1] = float(xy[0]), float(xy[1])
        for k, v in over.items():
            setattr(c, k, v)
        c.paths, c.lens = self.paths.ctypes.data, self.lens.ctypes.data
        c.num_paths, c.path_stride = self.paths.shape[0], self.paths.shape[2]
        self.cfg = c
        nact = 2 * min(K, actions_in_input)
        z = np.zeros
        self.a = {
            "hpos": z((E, Nh, 2)), "hgoal": z((E, Nh, 2)), "hhist": z((E, Nh, H, 2)), "hvel": z((E, Nh, 2), np.float32),
            "hstatic": z((E, Nh), np.uint8), "hvalid": z((E, Nh), np.int32), "rpose": z((E, 4)), "rprev": z((E, 4)),
            "rhist": z((E, H, 4)), "rpastv": z((E, LB, 2)), "rvalid": z(E, np.int32), "gtime": z(E), "path_idx": z(E, np.int32),
            "loc_to_start": z(E, np.uint8), "step_count": z(E, np.int64), "reset_epoch": z(E, np.int32),
            "warm_T": z((E, K, 4)), "warm_u": z((E, K, 2)), "iteration0": np.ones(E, np.int32),
            "pt_state": z((E, 4 * 25)), "mpc_actions": z((E, nact)), "robot_node": z((E, 2 * LB)), "past_robot_vel": z((E, 2 * LB)),
            "percent_episode": z((E, 1)), "spatial_edges": z((E, Nh, 2 * H)), "human_valid": z((E, Nh)), "detected_humans": z(E),
            "rewards": z((E, 3)), "info": z((E, 8)), "done_bits": z(E, np.uint32), "mpc_status": z(E, MPC_STATS_DTYPE),
            "sr_state": z((E, 4), np.int32), "actions_used": z((E, 2)), "soft_flags": z((E, 2), np.uint8),
        }
        self.A = EnvArrays(*[self.a[n].ctypes.data for n in _ENV_ARRAYS])
        L = lib()
        L.env_ref_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.env_ref_observe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.env_ref_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]

    def reset(self, hpos_init=None, hgoal_init=None, nthreads=0):
        p = np.ascontiguousarray(hpos_init, dtype=np.float64) if hpos_init is not None else None
        g = np.ascontiguousarray(hgoal_init, dtype=np.float64) if hgoal_init is not None else None
        lib().env_ref_reset(ctypes.byref(self.cfg), ctypes.byref(self.A), _p(p), _p(g), nthreads)
        return self.a

    def crowd_step(self, nthreads=0):
        """ORCA-only step: one pass + integration, robot frozen (BASELINE config 4). Returns (arrays, #LP3 solves)."""
        L = lib()
        L.env_ref_crowd_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        n = ctypes.c_int64(0)
        L.env_ref_crowd_step(ctypes.byref(self.cfg), ctypes.byref(self.A), nthreads, ctypes.byref(n))
        return self.a, int(n.value)

    def step(self, actions, nthreads=0):
        act = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.E, 2)
        lib().env_ref_step(ctypes.byref(self.cfg), ctypes.byref(self.A), _p(act), nthreads)
        return self.a
