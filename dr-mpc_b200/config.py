"""Engine configuration. Defaults are the reference's (scripts/configs/config_general.py:21-38,
config_HA.py:27-36,40-56,59-77,100-106, config_PT.py:7-22,38-39,48-50); `from_config_master` reads the
same attribute bags the reference's env classes read, so an unmodified ConfigMaster() can be passed."""
from dataclasses import dataclass, field
import ctypes
from . import _lib


@dataclass
class EngineConfig:
    num_envs: int = 1
    num_humans: int = 6
    lookback: int = 6
    time_step: float = 0.25
    time_limit: float = 50.0
    # ORCA (config_HA.orca, orca.py:79-89)
    neighbor_dist: float = 10.0
    time_horizon: float = 5.0
    safety_space: float = 0.175
    max_neighbors: int = 0          # 0 = all other agents (orca.py:77)
    group_lanes: int = 0            # 0 = auto
    human_radius: float = 0.3
    human_v_max: float = 1.0
    robot_radius: float = 0.3
    robot_visible: bool = True
    circle_radius: float = 4.0
    circle_radius_humans: float = 5.0
    end_goal_changing: bool = True
    # MPC (config_PT.MPC, mpc.py:21-53,144)
    horizon: int = 4
    actions_in_input: int = 4
    v_max: float = 1.0
    w_max: float = 1.0
    node_separation: float = 0.2
    mpc_max_iterations: int = 1500
    mpc_abs_change_tol: float = 1e-4
    mpc_rel_change_tol: float = 1e-4
    mpc_pose_jac_exact: bool = False
    mpc_max_retries: int = 20
    # HA rewards (config_HA.py:27-36)
    collision_penalty: float = -15.0
    safety_dist: float = 0.22
    safety_penalty: float = -15.0
    disturbance_radius: float = 1.5
    disturbance_factor_v: float = 0.8 * 7
    disturbance_factor_th: float = 0.5 * 7
    actuation_min_vel: float = 0.025
    actuation_penalty: float = -20.0
    # PT rewards (config_PT.py:11-22)
    pt_overall_scaling: float = 0.5
    goal_radius: float = 0.3
    goal_angle: float = 0.5
    goal_reward: float = 0.0
    path_advancement_scaling: float = 2.0
    deviation_xy: float = 1.0
    deviation_th: float = 0.05
    deviation_scale: float = 0.9
    corridor_penalty: float = -20.0
    corridor_max_dev: float = 1.6
    end_of_path_penalty: float = -10.0
    safety_corridor_penalty: float = -20.0
    safety_corridor_max_dev: float = 1.4
    lookahead: int = 20
    lookbehind: int = 5
    auto_path_switch: bool = True
    soft_reset_on_device: bool = False   # run HAAndPTEnv.soft_reset as a per-env state machine inside step()
    orca_dedup: bool = False             # reuse step t's look-ahead ORCA pass as step t+1's movement pass (bit-identical)
    orca_lookahead_prune: bool = False   # look-ahead pass only for the humans the disturbance reward reads (bit-identical)
    robot_sensor_restriction: bool = False   # config_HA.robot.sensor_restriction ("TODO: if True, verify correctness" upstream)
    robot_sensor_range: float = 4.0          # config_HA.robot.sensor_range
    static_humans: tuple = ()            # ((x, y), ...) static obstacles = the LAST len(static_humans) of num_humans
    seed: int = 0x00D24D9C
    env_id_offset: int = 0

    @classmethod
    def from_config_master(cls, cm, **overrides):
        g, ha, pt = cm.config_general, cm.config_HA, cm.config_PT
        rp, pp = ha.rewards.params, pt.rewards.params
        c = cls(
            num_humans=ha.sim.human_num, lookback=g.env.lookback, time_step=g.env.time_step, time_limit=g.env.time_limit,
            neighbor_dist=ha.orca.neighbor_dist, time_horizon=ha.orca.time_horizon, safety_space=ha.orca.safety_space,
            human_radius=ha.humans.radius, human_v_max=ha.humans.v_max, robot_radius=ha.robot.radius,
            robot_visible=bool(ha.robot.visible), circle_radius=ha.sim.circle_radius,
            circle_radius_humans=ha.sim.circle_radius_humans, end_goal_changing=bool(ha.humans.end_goal_changing),
            horizon=pt.MPC.planning_horizon, actions_in_input=pt.MPC.actions_in_input, v_max=g.robot.v_max,
            w_max=g.robot.w_max, node_separation=pt.env.node_separation,
            collision_penalty=rp['collision']['penalty'], safety_dist=rp['safety_human_raise']['safety_dist'],
            safety_penalty=rp['safety_human_raise']['penalty'], disturbance_radius=rp['disturbance']['radius'],
            disturbance_factor_v=rp['disturbance']['factor_v'], disturbance_factor_th=rp['disturbance']['factor_th'],
            actuation_min_vel=rp['actuation_termination']['min_vel'], actuation_penalty=rp['actuation_termination']['penalty'],
            pt_overall_scaling=pt.rewards.overall_scaling, goal_radius=pp['goal']['goal_radius'],
            goal_angle=pp['goal']['goal_angle'], goal_reward=pp['goal']['goal_reward'],
            path_advancement_scaling=pp['path_advancement']['path_advancement_scaling'],
            deviation_xy=pp['deviation']['xy'], deviation_th=pp['deviation']['th'], deviation_scale=pp['deviation']['scale'],
            corridor_penalty=pp['corridor_hit']['penalty'], corridor_max_dev=pp['corridor_hit']['max_deviation'],
            end_of_path_penalty=pp['deviation_end_of_path']['penalty'],
            safety_corridor_penalty=pp['safety_corridor_raise']['penalty'],
            safety_corridor_max_dev=pp['safety_corridor_raise']['max_deviation'],
            lookahead=pt.PT_model_params['MLP_local'].lookahead, lookbehind=pt.PT_model_params['MLP_local'].lookbehind,
        )
        c.robot_sensor_restriction = bool(getattr(ha.robot, "sensor_restriction", False))
        c.robot_sensor_range = float(getattr(ha.robot, "sensor_range", 4.0))
        # continuous task with static humans on: two hard-coded obstacles appended after the dynamic humans
        # (crowd_sim.py:156-158,171-183; config_HA.py:46)
        st = getattr(ha.sim, "include_static_humans", None)
        if st and st.get('episode_probability', 0) > 0 and getattr(g.env, "continuous_task", True):
            c.static_humans = ((-4.0, 0.0), (0.0, 0.0))
            c.num_humans += 2
        for k, v in overrides.items():
            setattr(c, k, v)
        return c

    # ---- C structs ----
    def orca_params(self):
        return _lib.OrcaParams(self.neighbor_dist, self.time_horizon, self.time_step,
                               self.human_radius + 0.01 + self.safety_space,   # orca.py:85,88
                               self.human_v_max, self.max_neighbors, int(self.robot_visible), self.group_lanes)

    def mpc_params(self):
        return _lib.MpcParams(self.horizon, self.time_step, self.v_max, self.w_max, self.node_separation,
                              self.mpc_max_iterations, self.mpc_abs_change_tol, self.mpc_rel_change_tol,
                              int(self.mpc_pose_jac_exact), self.mpc_max_retries)

    def env_config(self):
        return _lib.EnvConfig(
            self.num_envs, self.num_humans, self.lookback, self.orca_params(), self.mpc_params(), self.time_step,
            self.time_limit, self.human_radius, self.robot_radius, self.circle_radius, self.circle_radius_humans,
            int(self.end_goal_changing), self.collision_penalty, self.safety_dist, self.safety_penalty,
            self.disturbance_radius, self.disturbance_factor_v, self.disturbance_factor_th, self.actuation_min_vel,
            self.actuation_penalty, self.pt_overall_scaling, self.goal_radius, self.goal_angle, self.goal_reward,
            self.path_advancement_scaling, self.deviation_xy, self.deviation_th, self.deviation_scale,
            self.corridor_penalty, self.corridor_max_dev, self.end_of_path_penalty, self.safety_corridor_penalty,
            self.safety_corridor_max_dev, self.lookahead, self.lookbehind, self.actions_in_input,
            int(self.auto_path_switch), int(self.soft_reset_on_device), int(self.orca_dedup), int(self.orca_lookahead_prune), int(self.robot_sensor_restriction), float(self.robot_sensor_range),
            len(self.static_humans),
            (ctypes.c_double * 8)(*([float(v) for xy in self.static_humans for v in xy] + [0.0] * (8 - 2 * len(self.static_humans)))),
            self.seed, self.env_id_offset)
