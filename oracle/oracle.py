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
