"""Literal drop-in adapters: the reference's OWN call signatures, objects and return types (one env, numpy with a leading
dim of 1, Python floats / bools / sets) on top of the C ABI of libdre.so. A maintainer swaps three imports (plus, if wanted,
the replay buffer's):

    scripts/policy/policy_factory.py:8      policy_factory['orca'] = dr_mpc_b200.compat.ORCA
    environment/HA_and_PT/...env.py:47      self.MPC = dr_mpc_b200.compat.MPC(config_master, K, dt, continuous_task, v_max, w_max)
    scripts/online_continuous_task.py:39    self.env = dr_mpc_b200.compat.HAAndPTEnv(config_master, seed_addition=run)

and the rest of the reference (HAEnv / PTEnv around the first two, the SAC loop around the third) runs unchanged. These are
the E = 1 doors; the batched classes (BatchedORCA / BatchedMPC / BatchedHAAndPTEnv) are where the throughput is.
There is no CPU fallback: every call goes through libdre.so and raises if it is missing.

Reference interfaces mirrored (file:line under /root/reference):
  ORCA(config).predict(JointState) -> ActionXY                      scripts/policy/orca.py:7-117
  MPC(config_master, K, time_step, continuous_task, max_v, max_w)   environment/path_tracking/utils/mpc.py:18-56
     .reset(), .act({'path','interp_index','pose'}) -> ndarray(2K)  mpc.py:55-62,206
  HAAndPTEnv(config_master, seed_addition)                          environment/HA_and_PT/human_avoidance_and_path_tracking_env.py:31-41
     .reset() -> S                                                  :243-280
     .step(ActionRot, update, soft_reset, model_info) -> 5-tuple    :305-416
     .soft_reset(done_return, info, original_Sprime) -> S | None    :120-240
     .cvmm_diverse_safety_pipeline(np(2,), num_steps, backup) -> (ActionRot, info)   :422-473
     .cvmm_safety_check(np(2,), num_steps) -> (bool, float)         :475-513
     .create_video_continuous_task / .no_video / .render            :88-118,515 (no-ops: rendering is out of scope)
  ReplayBuffer(obs_example, config_master, buffer_size)             scripts/utils/storage.py:6-35
     .insert(S, A, S_prime, R, done), .sample_from_buffer(batch_size, indices_to_sample=None), .grab_all_data(batch_size),
     .full(), .valid_entries, .replay_buffer                        storage.py:37-38,54-145
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import _lib
from .config import EngineConfig

# environment/human_avoidance/utils/action.py:5-6, state.py:6-7 -- same field names, so either side's tuples work
ActionXY = namedtuple('ActionXY', ['vx', 'vy'])
ActionRot = namedtuple('ActionRot', ['v', 'w'])
FullState = namedtuple('FullState', ['px', 'py', 'vx', 'vy', 'radius', 'gx', 'gy', 'v_pref', 'theta'])
ObservableState = namedtuple('ObservableState', ['px', 'py', 'vx', 'vy', 'radius'])


def _hp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _action_types():
    """The reference's own ActionXY / ActionRot when its package is importable (isinstance checks in agent.py:157-162),
    ours otherwise."""
    try:
        from environment.human_avoidance.utils.action import ActionXY as AXY, ActionRot as AR
        return AXY, AR
    except ImportError:
        return ActionXY, ActionRot


# ---------------------------------------------------------------------------------------------------------------
# ORCA policy (boundary B)
# ---------------------------------------------------------------------------------------------------------------
def orca_pack_joint_state(state, safety_space):
    """JointState -> (self8 float32[8], others5 float32[n,5]): exactly what orca.py:84-111 hands to its private rvo2
    simulator (positions / velocities as they are, preferred velocity = goal direction normalised only beyond 1 m, radii
    + 0.01 + safety_space, ego max speed = v_pref), cast to float like the Cython boundary."""
    s = state.self_state
    velocity = np.array((s.gx - s.px, s.gy - s.py))
    speed = np.linalg.norm(velocity)
    pref_vel = velocity / speed if speed > 1 else velocity
    self8 = np.array([s.px, s.py, s.vx, s.vy, pref_vel[0], pref_vel[1], s.radius + 0.01 + safety_space, s.v_pref], dtype=np.float32)
    others = np.array([[h.px, h.py, h.vx, h.vy, h.radius + 0.01 + safety_space] for h in state.human_states],
                      dtype=np.float32).reshape(len(state.human_states), 5)
    return self8, others


class ORCA:
    """policy_factory['orca'] replacement. `config` is config_HA (reads config.orca.*); `time_step` is set by the owner
    (agent.py:31)."""

    def __init__(self, config):
        self.config = config
        self.name = 'ORCA'
        self.trainable = False
        self.phase = self.model = self.device = self.env = None
        self.last_state = None
        self.time_step = None
        self.max_neighbors = None
        self.radius = None
        self.max_speed = 1
        self.sim = None
        self.safety_space = self.config.orca.safety_space
        self._L = _lib.load()

    def _solve(self, params, self8, others5):
        """One C-ABI call (dre_orca_solve_host) for N joint states: self8 [N,8], others5 [N,n,5] -> [N,2] float32."""
        N, n = self8.shape[0], others5.shape[1]
        out = np.empty((N, 2), dtype=np.float32)
        _lib.check(self._L.dre_orca_solve_host(ctypes.byref(params), N, n, _hp(self8), _hp(others5), _hp(out), None))
        return out

    def _params(self, n_others):
        o = self.config.orca
        if self.time_step is None:
            raise ValueError("ORCA.time_step must be set by the owner before predict() (agent.py:31)")
        return _lib.OrcaParams(float(o.neighbor_dist), float(o.time_horizon), float(self.time_step), 0.0, 1.0, int(n_others), 1, 0)

    def predict(self, state):
        self.max_neighbors = len(state.human_states)          # orca.py:77-78
        self.radius = state.self_state.radius
        self8, others = orca_pack_joint_state(state, self.safety_space)
        v = self._solve(self._params(self.max_neighbors), np.ascontiguousarray(self8[None]), np.ascontiguousarray(others[None]))
        AXY, _ = _action_types()
        action = AXY(float(v[0, 0]), float(v[0, 1]))
        self.last_state = state
        return action

    def predict_many(self, states):
        """Not in the reference: every human's joint state of one crowd step in ONE call (CrowdSim.get_human_actions
        issues them one by one, crowd_sim.py:264-284). All states must have the same number of neighbours."""
        packed = [orca_pack_joint_state(s, self.safety_space) for s in states]
        self8 = np.ascontiguousarray(np.stack([p[0] for p in packed]))
        others = np.ascontiguousarray(np.stack([p[1] for p in packed]))
        v = self._solve(self._params(others.shape[1]), self8, others)
        AXY, _ = _action_types()
        return [AXY(float(a), float(b)) for a, b in v]

    @staticmethod
    def reach_destination(state):
        s = state.self_state
        return bool(np.linalg.norm((s.py - s.gy, s.px - s.gx)) < s.radius)


# ---------------------------------------------------------------------------------------------------------------
# MPC (boundary B)
# ---------------------------------------------------------------------------------------------------------------
class MPC:
    """environment.path_tracking.utils.MPC replacement: one problem, warm start kept inside the engine."""

    def __init__(self, config_master, K, time_step, continuous_task, max_v, max_w):
        from .mpc import BatchedMPC
        self.config_master = config_master
        self.K = K
        self.time_step = time_step
        self.continuous_task = continuous_task
        self.max_v_scalar, self.max_w_scalar = max_v, max_w
        self.travel_dist_in_DT = (self.max_v_scalar * 0.99) * self.time_step
        self.node_idx_increment = self.travel_dist_in_DT / self.config_master.config_PT.env.node_separation
        self._mpc = BatchedMPC(config_master, K, time_step, continuous_task, max_v, max_w, num_problems=1)
        self._paths = []            # every distinct 3xL path seen so far, in arrival order
        self._keys = {}
        self.iteration0 = True
        self.last_status = None

    def reset(self):
        self.iteration0 = True
        self._mpc.reset()

    def _path_id(self, path):
        p = np.ascontiguousarray(path, dtype=np.float64)
        key = (p.shape, p.tobytes())
        if key not in self._keys:
            self._keys[key] = len(self._paths)
            self._paths.append(p)
            self._mpc.set_paths(self._paths)     # path table re-uploaded (rare: the continuous task has four paths)
        return self._keys[key]

    def act(self, state_gen_input):
        pid = self._path_id(state_gen_input['path'])
        u = self._mpc.act({'path_id': np.array([pid], dtype=np.int32),
                           'interp_index': np.array([float(state_gen_input['interp_index'])]),
                           'pose': np.asarray(state_gen_input['pose'], dtype=np.float64).reshape(1, 3)}, want_status=True)
        self.iteration0 = False
        self.last_status = self._mpc.last_status[0]
        return np.array(u[0], dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------------
# Combined env (boundary A)
# ---------------------------------------------------------------------------------------------------------------
_REASONS = (('collision', 1 << 0), ('safety_human_raise', 1 << 1), ('actuation_termination', 1 << 2), ('timeout', 1 << 3),
            ('goal', 1 << 4), ('deviation_end_of_path', 1 << 5), ('corridor_hit', 1 << 6), ('safety_corridor_raise', 1 << 7))
_HA_REASONS = ('collision', 'safety_human_raise', 'actuation_termination', 'timeout')


class _RobotView:
    """env.robot of the reference (shared Robot object): read-only pose / radius of the engine's robot."""

    def __init__(self, owner):
        self._o = owner
        self.kinematics = 'unicycle'

    def _pose(self):
        return self._o._env.state["rpose"][0, :3].cpu().numpy()

    px = property(lambda self: float(self._pose()[0]))
    py = property(lambda self: float(self._pose()[1]))
    theta = property(lambda self: float(self._pose()[2]))
    radius = property(lambda self: float(self._o._env.cfg.robot_radius))

    def get_pose_np(self):
        return np.array(self._pose(), dtype=np.float64)

    def get_position_np(self):
        return np.array(self._pose()[:2], dtype=np.float64)


class HAAndPTEnv:
    """HAAndPTEnv(config_master, seed_addition) over one engine env. Reference semantics kept literally: the path switch
    on goal / end of path happens inside soft_reset() (auto_path_switch off), soft_reset() issues its scripted
    self.step(action, soft_reset=True) calls one by one, every returned array is a fresh numpy copy owned by the caller."""

    def __init__(self, config_master, seed_addition=0, device="cuda:0", **engine_overrides):
        from .env import BatchedHAAndPTEnv
        self.config_master = config_master
        self.config_general = config_master.config_general
        self.config_PT = config_master.config_PT
        self.config_HA = config_master.config_HA
        self.policy = self.config_general.model.policy
        self.time_step = self.config_general.env.time_step
        self.use_PEB = self.config_general.model.use_PEB
        self.continuous_task = self.config_general.env.continuous_task
        if not self.continuous_task:
            raise NotImplementedError               # reset_episodic_task / step_episodic_task raise upstream too (env.py:282,418)
        self.device = self.config_general.device
        self.case_start = {'train': 0, 'test': np.iinfo(np.uint32).max - 3000}
        self.case_capacity = {'train': np.iinfo(np.uint32).max - 3000, 'test': 2999}
        self.case_counter = {'train': 0, 'test': 0}
        over = dict(auto_path_switch=False, soft_reset_on_device=False)
        over.update(engine_overrides)
        cfg = EngineConfig.from_config_master(config_master, num_envs=1, **over)
        self._seed0 = cfg.seed + seed_addition
        self._env = BatchedHAAndPTEnv(cfg, num_envs=1, device=device)
        self.MPC = self._env if self.policy in ['DR-MPC', 'ResidualDRL'] else None
        self.robot = _RobotView(self)
        self.create_vid = False
        self._act = np.zeros((1, 2), dtype=np.float64)
        self.prev_robot_pose = None
        self.robot_change = None

    # ---- rendering hooks the training loop calls: video is out of scope, keep the calls harmless ----
    def create_video_continuous_task(self, vid_name, S_PT):
        self.vid_name = vid_name

    def no_video(self):
        self.create_vid = False

    def render(self):
        pass

    # ---- observation packaging: fresh numpy arrays with the reference's shapes (leading dim 1) ----
    def _S(self, views):
        c = lambda k: np.array(views[k], dtype=np.float64, copy=True)
        S_PT = {'PT_state': c("pt_state")}
        if self.MPC is not None:
            S_PT['MPC_actions'] = c("mpc_actions")
        S_HA = {'robot_node': c("robot_node"), 'past_robot_velocities': c("past_robot_vel"), 'percent_episode': c("percent_episode"),
                'spatial_edges': c("spatial_edges"), 'detected_human_num': c("detected_humans").astype(np.int64),
                'human_valid_history': c("human_valid")}
        return {'PT': S_PT, 'HA': S_HA}

    def _device_views(self):
        return {k: v.cpu().numpy() for k, v in self._env.out.items()}

    def reset(self, phase='train', test_case=None, curved_path=False):
        if test_case is None:
            case_num = self.case_start[phase] + self.case_counter[phase]
            self.case_counter[phase] = (self.case_counter[phase] + 1) % self.case_capacity[phase]
        else:
            case_num = test_case
        # np.random.seed(case_num) upstream (env.py:252); here the case number keys the Philox streams of the spawn
        self._env.set_seed(self._seed0 + 0x9E3779B97F4A7C15 * int(case_num))
        self._env.reset()
        self.prev_robot_pose = self.robot.get_pose_np()
        self.robot_change = None
        return self._S(self._device_views())

    def reset_from(self, human_positions, human_goals):
        """Not in the reference: reset() with the humans placed by the caller (parity runs: the reference spawns from the
        global np.random stream, which cannot be matched draw for draw)."""
        import torch
        hp = torch.as_tensor(np.asarray(human_positions, dtype=np.float64)[None], device=self._env.device)
        hg = torch.as_tensor(np.asarray(human_goals, dtype=np.float64)[None], device=self._env.device)
        self._env.reset(hp, hg)
        self.prev_robot_pose = self.robot.get_pose_np()
        return self._S(self._device_views())

    def step(self, action, update=True, soft_reset=False, model_info=None):
        self._act[0, 0], self._act[0, 1] = float(action[0]), float(action[1])
        prev = self.robot.get_pose_np()
        v = self._env.step_host(self._act)                      # dre_env_step_host: host buffers in, host slab out
        S_prime = self._S(v)
        bits = int(v["done_bits"][0])
        r = v["rewards"][0]
        reward_return = {'R_PT': float(r[0]), 'R_DO': float(r[1]), 'R': float(r[2])}
        fl = v["done_flags"][0]
        done_return = {'done_PT': bool(fl[0]), 'done_HA': bool(fl[1]), 'done': bool(fl[2]), 'done_for_replay': bool(fl[3])}
        if done_return['done'] and (bits & 0xFF) == (1 << 3) and self.use_PEB:      # pure timeout bootstraps (env.py:407-410)
            done_return['done_for_replay'] = False
        info = self._info(bits, v["info"][0])
        eps_done_info = {k: bool(fl[4 + i]) for i, k in enumerate(('success', 'collision', 'safety_human_raise', 'timeout',
                                                                     'deviation_end_of_path', 'corridor_hit',
                                                                     'safety_corridor_raise', 'actuation_termination'))}
        curr = self.robot.get_pose_np()
        self.robot_change = curr - prev
        self.prev_robot_pose = curr
        return S_prime, reward_return, done_return, info, eps_done_info

    def _info(self, bits, iv):
        names = {n for n, b in _REASONS if bits & b}
        done_HA = {'False'} | {n for n in names if n in _HA_REASONS}
        done_PT = {'False'} | {n for n in names if n not in _HA_REASONS}
        c = self._env.cfg
        sc = c.pt_overall_scaling
        return {'done': {'False'} | names, 'done_PT': done_PT, 'done_HA': done_HA,
                # reward_comp.py:20-32 (each scaled by overall_scaling) and human_avoidance_env.py:189-291
                'goal': sc * c.goal_reward if 'goal' in names else sc * 0.0,
                'deviation': float(iv[3]), 'path_advancement': float(iv[2]),
                'deviation_end_of_path': sc * c.end_of_path_penalty if 'deviation_end_of_path' in names else sc * 0.0,
                'safety_corridor_raise': sc * c.safety_corridor_penalty if 'safety_corridor_raise' in names else sc * 0.0,
                'corridor_hit': sc * c.corridor_penalty if 'corridor_hit' in names else sc * 0.0,
                'collision': c.collision_penalty if 'collision' in names else 0,
                'safety_human_raise': c.safety_penalty if 'safety_human_raise' in names else 0,
                'disturbance_v': float(iv[0]), 'disturbance_th': float(iv[1]),
                'actuation_termination': c.actuation_penalty if 'actuation_termination' in names else 0}

    def soft_reset(self, done_return, info, original_Sprime):
        """HAAndPTEnv.soft_reset (env.py:120-240): the loop runs here, on the host, like upstream; each decision of its body
        (which scripted action, when to stop, stuck) is the engine's soft_reset_decide on the latest step's done word."""
        _, AR = _action_types()
        Sprime = original_Sprime
        while True:
            scripted, act, stuck = self._env.soft_reset_decide(commit=True)
            if bool(stuck[0]):
                self.render()
                print("Stuck in soft reset")
                return None
            if not bool(scripted[0]):
                break
            a = act[0].cpu().numpy()
            Sprime, reward_return, done_return, info, eps_done_info = self.step(AR(float(a[0]), float(a[1])), soft_reset=True)
        if 'goal' in info['done'] or 'deviation_end_of_path' in info['done']:
            # next path, localise to its start, MPC.reset(), regenerate S (env.py:130-148)
            self._env.switch_path()
            self._env.observe(solve_mpc=self.MPC is not None)
            return self._S(self._device_views())
        return Sprime

    def cvmm_safety_check(self, action_execute, num_steps):
        import torch
        a = torch.as_tensor(np.asarray(action_execute, dtype=np.float64).reshape(1, 2), device=self._env.device)
        safe, dist = self._env.cvmm_safety_check(a, num_steps)
        return bool(safe[0]), float(dist[0])

    def cvmm_diverse_safety_pipeline(self, action_execute, num_steps, backup_actions):
        import torch
        _, AR = _action_types()
        a_np = np.asarray(action_execute, dtype=np.float64).reshape(2)
        a = torch.as_tensor(a_np.reshape(1, 2), device=self._env.device)
        bk = np.asarray(backup_actions, dtype=np.float64)
        out, d = self._env.cvmm_diverse_safety_pipeline(a, num_steps, bk)
        o = out[0].cpu().numpy()
        info = {}
        if int(d['original_safe'][0]):
            act = AR(a_np[0], a_np[1])
            info['RL_action_bool'] = f'NA: ({act.v}, {act.w})'
            return act, info
        act = AR(o[0], o[1])
        if int(d['random'][0]):
            info['RL_action_bool'] = f'no safe: ({act.v}, {act.w})'
        else:
            info['RL_action_bool'] = f"{int(d['num_safe'][0])}:{int(d['chosen'][0])}:({act.v},{act.w})"
        return act, info


class ReplayBuffer:
    """scripts/utils/storage.py:6-145 with the reference's own signatures (one transition per insert, tensors with a leading
    dim of 1, `indices_to_sample` chronological) on top of BatchedReplayBuffer's ring: a full buffer overwrites its oldest
    row instead of rolling every tensor by one (storage.py:56-75), which is invisible through this interface --
    `replay_buffer` presents the tensors in the reference's chronological order. Fourth import a maintainer can swap:
    scripts/online_continuous_task.py `from scripts.utils.storage import ReplayBuffer`."""

    def __init__(self, obs_example, config_master, buffer_size):
        from .replay import BatchedReplayBuffer
        self.config_master = config_master
        self.buffer_size = buffer_size
        self.device = config_master.config_general.device          # storage.py:12
        self._ring = BatchedReplayBuffer(obs_example, buffer_size, device=self.device)

    @property
    def valid_entries(self):
        return self._ring.valid_entries

    def full(self):
        return self._ring.full()

    @property
    def replay_buffer(self):
        """The reference's dict of tensors, rows in insertion order (storage.py:18-35). Views while the ring has not
        wrapped; gathered copies afterwards."""
        import torch
        rb = self._ring.replay_buffer
        if not self._ring.full() or self._ring._head == 0:
            return rb
        order = self._ring.physical_index(torch.arange(self.buffer_size, device=self._ring.device))
        pick = lambda v: {k: pick(x) for k, x in v.items()} if isinstance(v, dict) else v[order]
        return {k: pick(v) for k, v in rb.items()}

    def insert(self, S, A, S_prime, R, done):
        import torch
        t = lambda x, w: torch.as_tensor(x, dtype=torch.float64).reshape(1, w)
        self._ring.insert(S, t(A, 2), S_prime, t(float(R), 1), t(float(done), 1))

    def sample_from_buffer(self, batch_size, indices_to_sample=None):
        import torch
        if indices_to_sample is None:
            if self.valid_entries < batch_size:                      # storage.py:103-106 only warns
                print(f"Not enough entries in buffer to sample from. Only {self.valid_entries} entries in buffer")
            indices_to_sample = torch.randint(0, self.valid_entries, (batch_size,))
        return self._ring.sample_from_buffer(batch_size, indices_to_sample)

    def grab_all_data(self, batch_size):
        import torch
        out, cur = [], 0
        while cur < self.valid_entries:                              # storage.py:135-145
            end = min(cur + batch_size, self.valid_entries)
            out.append(self.sample_from_buffer(0, indices_to_sample=torch.arange(cur, end)))
            cur = end
        return out
