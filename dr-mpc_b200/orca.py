"""Batched ORCA policy: mirrors reference scripts/policy/orca.py (`ORCA(config).predict(JointState)
-> ActionXY`) and its caller CrowdSim.get_human_actions (crowd_sim.py:264-284) for E envs x Nh
humans at once. One call = one ORCA solve for every human of every env."""
import ctypes
import numpy as np
from . import _lib
from .config import EngineConfig


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _dptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _hptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class BatchedORCA:
    name = 'ORCA'

    def __init__(self, config=None):
        """`config`: an EngineConfig, or the reference's config_HA-like object (orca.py:8,61 reads
        config.orca.*), or None for the reference defaults."""
        if config is None:
            config = EngineConfig()
        elif not isinstance(config, EngineConfig):
            ha = config
            config = EngineConfig(neighbor_dist=ha.orca.neighbor_dist, time_horizon=ha.orca.time_horizon,
                                  safety_space=ha.orca.safety_space, human_radius=ha.humans.radius,
                                  human_v_max=ha.humans.v_max, robot_visible=bool(ha.robot.visible),
                                  time_step=ha.env.time_step)
        self.config = config
        self.time_step = config.time_step
        self.max_speed = 1
        self.safety_space = config.safety_space
        self.last_stats = None
        self._L = _lib.load()

    def predict(self, hpos, hvel, goal, rpos, is_static=None, want_stats=False):
        """hpos/goal [E,Nh,2] float64, hvel [E,Nh,2] float32, rpos [E,2] float64 -> vnew [E,Nh,2] float32.
        torch CUDA tensors run in place on the caller's stream; numpy arrays go through the host entry point."""
        p = self.config.orca_params()
        p.time_step = self.time_step
        E, Nh = int(hpos.shape[0]), int(hpos.shape[1])
        if _is_torch(hpos):
            import torch
            assert hpos.is_cuda and hpos.dtype == torch.float64 and hvel.dtype == torch.float32
            hpos, hvel, goal, rpos = hpos.contiguous(), hvel.contiguous(), goal.contiguous(), rpos.contiguous()
            vnew = torch.empty((E, Nh, 2), dtype=torch.float32, device=hpos.device)
            stats = torch.empty((E, Nh, 6), dtype=torch.int32, device=hpos.device) if want_stats else None
            st8 = is_static.to(torch.uint8).contiguous() if is_static is not None else None
            stream = ctypes.c_void_p(torch.cuda.current_stream(hpos.device).cuda_stream)
            with torch.cuda.device(hpos.device):
                _lib.check(self._L.dre_orca_predict(ctypes.byref(p), E, Nh, _dptr(hpos), _dptr(hvel), _dptr(goal),
                                                    _dptr(rpos), _dptr(st8), _dptr(vnew), _dptr(stats), stream))
            self.last_stats = stats
            return vnew
        hpos = np.ascontiguousarray(hpos, dtype=np.float64)
        hvel = np.ascontiguousarray(hvel, dtype=np.float32)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        rpos = np.ascontiguousarray(rpos, dtype=np.float64)
        st8 = np.ascontiguousarray(is_static, dtype=np.uint8) if is_static is not None else None
        vnew = np.empty((E, Nh, 2), dtype=np.float32)
        stats = np.empty((E, Nh, 6), dtype=np.int32) if want_stats else None
        _lib.check(self._L.dre_orca_predict_host(ctypes.byref(p), E, Nh, _hptr(hpos), _hptr(hvel), _hptr(goal), _hptr(rpos),
                                                 _hptr(st8), _hptr(vnew), _hptr(stats)))
        self.last_stats = stats
        return vnew
