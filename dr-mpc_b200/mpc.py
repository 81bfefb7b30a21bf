"""Batched path-tracking MPC: mirrors reference environment/path_tracking/utils/mpc.py
(`MPC(config_master, K, time_step, continuous_task, max_v, max_w)`, `.reset()`, `.act(dict) -> ndarray(2K)`)
for N independent problems, each with its own warm start (mpc.py:51-56,192-205)."""
import ctypes
import numpy as np
from . import _lib
from .config import EngineConfig
from .orca import _is_torch, _dptr, _hptr
from .paths import pack_paths

MPC_STATUS_DTYPE = np.dtype([("iterations", "i4"), ("retries", "i4"), ("lm_solves", "i4"), ("cost_evals", "i4"),
                             ("term", "i4"), ("zero_steps", "i4"), ("final_cost", "f8")])


class BatchedMPC:
    def __init__(self, config_master=None, K=4, time_step=0.25, continuous_task=True, max_v=1.0, max_w=1.0,
                 num_problems=1, node_separation=None, pose_jac_exact=False, max_retries=20):
        if node_separation is None:
            node_separation = config_master.config_PT.env.node_separation if config_master is not None else 0.2
        self.config_master = config_master
        self.K, self.time_step, self.continuous_task = K, time_step, continuous_task
        self.max_v_scalar, self.max_w_scalar = max_v, max_w
        self.N = int(num_problems)
        self.params = EngineConfig(horizon=K, time_step=time_step, v_max=max_v, w_max=max_w,
                                   node_separation=node_separation, mpc_pose_jac_exact=pose_jac_exact,
                                   mpc_max_retries=max_retries).mpc_params()
        self._L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._L.dre_mpc_create(ctypes.byref(self.params), self.N, ctypes.byref(h)))
        self._h = h
        self._paths_set = False
        self.last_status = None

    def __del__(self):
        if getattr(self, "_h", None) and self._L is not None:
            self._L.dre_mpc_destroy(self._h)
            self._h = None

    def set_paths(self, paths):
        """paths: list of 3xL float64 arrays (the reference's path layout)."""
        tab, lens = pack_paths(paths)
        _lib.check(self._L.dre_mpc_set_paths(self._h, _hptr(tab), tab.shape[0], tab.shape[2], _hptr(lens)))
        self._paths_set = True

    def reset(self, mask=None):
        """MPC.reset (mpc.py:55-56): next act() starts from the iteration0 initial guess. `mask` selects problems."""
        if mask is None:
            stream = None
            try:
                import torch
                if torch.cuda.is_available():   # same stream as act(): a reset queued on another stream would race the solve
                    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            except ImportError:
                pass
            _lib.check(self._L.dre_mpc_reset(self._h, None, stream))
        elif _is_torch(mask):
            import torch
            m = mask.to(torch.uint8).contiguous()
            _lib.check(self._L.dre_mpc_reset(self._h, _dptr(m), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.current_stream().synchronize()
        else:
            w_it0 = np.empty(self.N, dtype=np.uint8)
            _lib.check(self._L.dre_mpc_get_warm(self._h, None, None, _hptr(w_it0)))
            w_it0[np.asarray(mask, dtype=bool)] = 1
            _lib.check(self._L.dre_mpc_set_warm(self._h, None, None, _hptr(w_it0)))

    def act(self, state_gen_input, want_status=False):
        """state_gen_input: {'interp_index': [N], 'pose': [N,3], 'path_id': [N] int32 (optional),
        'path': 3xL array (optional, registers a single shared path like the reference's dict)}.
        Returns controls [N, 2K] = [v0,w0,...,v_{K-1},w_{K-1}] per problem."""
        if 'path' in state_gen_input and not self._paths_set:
            self.set_paths([np.asarray(state_gen_input['path'], dtype=np.float64)])
        interp, pose = state_gen_input['interp_index'], state_gen_input['pose']
        pid = state_gen_input.get('path_id')
        if _is_torch(pose):
            import torch
            interp = interp.to(torch.float64).contiguous()
            pose = pose.to(torch.float64).contiguous()
            pid = pid.to(torch.int32).contiguous() if pid is not None else None
            out = torch.empty((self.N, 2 * self.K), dtype=torch.float64, device=pose.device)
            status = torch.empty((self.N, 8), dtype=torch.int32, device=pose.device) if want_status else None
            stream = ctypes.c_void_p(torch.cuda.current_stream(pose.device).cuda_stream)
            _lib.check(self._L.dre_mpc_act(self._h, _dptr(pid), _dptr(interp), _dptr(pose), _dptr(out), _dptr(status), stream))
            self.last_status = status
            return out
        interp = np.ascontiguousarray(np.broadcast_to(np.asarray(interp, dtype=np.float64), (self.N,)))
        pose = np.ascontiguousarray(np.asarray(pose, dtype=np.float64).reshape(self.N, 3))
        pid = np.ascontiguousarray(pid, dtype=np.int32) if pid is not None else None
        out = np.empty((self.N, 2 * self.K), dtype=np.float64)
        status = np.empty(self.N, dtype=MPC_STATUS_DTYPE) if want_status else None
        _lib.check(self._L.dre_mpc_act_host(self._h, _hptr(pid), _hptr(interp), _hptr(pose), _hptr(out), _hptr(status)))
        self.last_status = status
        return out

    # warm-start access for parity tests / checkpoints (problem-major [N,K,4], [N,K,2], [N])
    def get_warm(self):
        T = np.empty((self.N, self.K, 4)); u = np.empty((self.N, self.K, 2)); it0 = np.empty(self.N, dtype=np.uint8)
        _lib.check(self._L.dre_mpc_get_warm(self._h, _hptr(T), _hptr(u), _hptr(it0)))
        return T, u, it0

    def set_warm(self, T=None, u=None, it0=None):
        T = np.ascontiguousarray(T, dtype=np.float64) if T is not None else None
        u = np.ascontiguousarray(u, dtype=np.float64) if u is not None else None
        it0 = np.ascontiguousarray(it0, dtype=np.uint8) if it0 is not None else None
        _lib.check(self._L.dre_mpc_set_warm(self._h, _hptr(T), _hptr(u), _hptr(it0)))
