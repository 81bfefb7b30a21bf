"""ctypes binding of libdre.so (include/dre.h). There is NO CPU fallback: if the CUDA library is
missing or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRE_LIB") or os.path.join(_HERE, "libdre.so")   # DRE_LIB: A/B builds of the same library

c_i32, c_u8p, c_f, c_d, c_vp, c_u64, c_i64 = (ctypes.c_int32, ctypes.POINTER(ctypes.c_uint8), ctypes.c_float,
                                              ctypes.c_double, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64)


class DreError(RuntimeError):
    pass


class OrcaParams(ctypes.Structure):
    _fields_ = [("neighbor_dist", c_f), ("time_horizon", c_f), ("time_step", c_f), ("radius", c_f),
                ("max_speed_self", c_f), ("max_neighbors", c_i32), ("robot_visible", c_i32), ("group_lanes", c_i32)]


class OrcaStats(ctypes.Structure):
    _fields_ = [("n_lines", c_i32), ("lp2_fail", c_i32), ("n_lp1", c_i32), ("lp1_work", c_i32),
                ("branch_mask", ctypes.c_uint32), ("n_lp3_proj", c_i32)]


class MpcParams(ctypes.Structure):
    _fields_ = [("horizon", c_i32), ("time_step", c_d), ("v_max", c_d), ("w_max", c_d), ("node_separation", c_d),
                ("max_iterations", c_i32), ("abs_change_tol", c_d), ("rel_change_tol", c_d), ("pose_jac_exact", c_i32),
                ("max_retries", c_i32)]


class MpcStatus(ctypes.Structure):
    _fields_ = [("iterations", c_i32), ("retries", c_i32), ("lm_solves", c_i32), ("cost_evals", c_i32),
                ("term", c_i32), ("zero_steps", c_i32), ("final_cost", c_d)]


class EnvConfig(ctypes.Structure):
    _fields_ = [("num_envs", c_i32), ("num_humans", c_i32), ("lookback", c_i32), ("orca", OrcaParams), ("mpc", MpcParams),
                ("time_step", c_d), ("time_limit", c_d), ("human_radius", c_d), ("robot_radius", c_d),
                ("circle_radius", c_d), ("circle_radius_humans", c_d), ("end_goal_changing", c_i32),
                ("collision_penalty", c_d), ("safety_dist", c_d), ("safety_penalty", c_d),
                ("disturbance_radius", c_d), ("disturbance_factor_v", c_d), ("disturbance_factor_th", c_d),
                ("actuation_min_vel", c_d), ("actuation_penalty", c_d),
                ("pt_overall_scaling", c_d), ("goal_radius", c_d), ("goal_angle", c_d), ("goal_reward", c_d),
                ("path_advancement_scaling", c_d), ("deviation_xy", c_d), ("deviation_th", c_d), ("deviation_scale", c_d),
                ("corridor_penalty", c_d), ("corridor_max_dev", c_d), ("end_of_path_penalty", c_d),
                ("safety_corridor_penalty", c_d), ("safety_corridor_max_dev", c_d),
                ("lookahead", c_i32), ("lookbehind", c_i32), ("mpc_actions_in_input", c_i32), ("auto_path_switch", c_i32),
                ("soft_reset_on_device", c_i32), ("orca_dedup", c_i32), ("orca_lookahead_prune", c_i32), ("sensor_restriction", c_i32), ("sensor_range", c_d), ("num_static_humans", c_i32), ("static_pos", c_d * 8), ("seed", c_u64), ("env_id_offset", c_i64)]


class EnvStateView(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist",
                                    "rpastv", "rvalid", "gtime", "path_idx", "loc_to_start", "step_count", "sr_state")]


class StepOut(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("pt_state", "actions_used", "mpc_actions", "robot_node", "past_robot_vel", "percent_episode",
                                    "spatial_edges", "human_valid", "detected_humans", "rewards", "info", "done_bits",
                                    "mpc_status", "done_flags", "slab")] + [("slab_bytes", ctypes.c_size_t)]


_REPLAY_KEYS = ("pt_state", "mpc_actions", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid",
                "detected_humans")


class ReplayFields(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in _REPLAY_KEYS]


class ReplayView(ctypes.Structure):
    _fields_ = [("S", ReplayFields), ("S_prime", ReplayFields), ("rewards", c_vp), ("actions", c_vp), ("done", c_vp),
                ("rows", c_i64), ("dtype", c_i32)]


_lib = None


def load():
    """Load libdre.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DreError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(needs nvcc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    L.dre_strerror.restype = ctypes.c_char_p
    L.dre_strerror.argtypes = [ctypes.c_int]
    L.dre_last_cuda_error.restype = ctypes.c_char_p
    L.dre_version.restype = ctypes.c_int
    L.dre_orca_predict.argtypes = [P(OrcaParams), c_i32, c_i32] + [c_vp] * 8
    L.dre_orca_predict_host.argtypes = [P(OrcaParams), c_i32, c_i32] + [c_vp] * 7
    L.dre_orca_solve_host.argtypes = [P(OrcaParams), c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]
    L.dre_mpc_create.argtypes = [P(MpcParams), c_i32, P(c_vp)]
    L.dre_mpc_destroy.argtypes = [c_vp]
    L.dre_mpc_set_paths.argtypes = [c_vp, c_vp, c_i32, c_i32, c_vp]
    L.dre_mpc_reset.argtypes = [c_vp, c_vp, c_vp]
    L.dre_mpc_act.argtypes = [c_vp] + [c_vp] * 6
    L.dre_mpc_act_host.argtypes = [c_vp] + [c_vp] * 5
    L.dre_mpc_get_warm.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.dre_mpc_set_warm.argtypes = [c_vp, c_vp, c_vp, c_vp]
    if hasattr(L, "dre_env_create"):
        L.dre_env_create.argtypes = [P(EnvConfig), P(c_vp)]
        L.dre_env_destroy.argtypes = [c_vp]
        L.dre_env_set_paths.argtypes = [c_vp, c_vp, c_i32, c_i32, c_vp]
        L.dre_env_state.argtypes = [c_vp, P(EnvStateView)]
        L.dre_env_outputs.argtypes = [c_vp, P(StepOut)]
        L.dre_env_outputs_at.argtypes = [c_vp, c_i32, P(StepOut)]
        L.dre_env_current_slab.argtypes = [c_vp]
        L.dre_env_reset_masked.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp]
        L.dre_env_observe_masked.argtypes = [c_vp, c_vp, c_i32, c_vp]
        L.dre_env_switch_path.argtypes = [c_vp, c_vp, c_vp]
        L.dre_env_soft_reset_decide.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]
        L.dre_stats_comm_id.argtypes = [c_vp]
        L.dre_env_stats_comm_init.argtypes = [c_vp, c_vp, c_i32, c_i32]
        L.dre_stats_reduce.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp]
        L.dre_env_mpc.restype = c_vp
        L.dre_env_mpc.argtypes = [c_vp]
        L.dre_env_reset.argtypes = [c_vp, c_vp, c_vp, c_vp]
        L.dre_env_step.argtypes = [c_vp, c_vp, c_vp]
        L.dre_env_crowd_step.argtypes = [c_vp, c_i32, c_vp]
        L.dre_env_step_host.argtypes = [c_vp, c_vp, c_vp]
        L.dre_env_set_host_chunks.argtypes = [c_vp, c_i32]
        L.dre_env_set_host_outputs.argtypes = [c_vp, c_i32]
        L.dre_env_d2h_probe.argtypes = [c_vp, c_vp, c_i32, P(c_d)]
        L.dre_env_invalidate_orca_cache.argtypes = [c_vp, c_vp]
        L.dre_env_observe.argtypes = [c_vp, c_i32, c_vp]
        L.dre_env_stats.argtypes = [c_vp, c_vp, c_vp, c_i32, c_vp]
        L.dre_env_cvmm_filter.argtypes = [c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]
        L.dre_env_cvmm_check.argtypes = [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]
        L.dre_env_set_seed.argtypes = [c_vp, c_u64]
        L.dre_replay_create.argtypes = [c_vp, c_i64, c_i32, P(c_vp)]
        L.dre_replay_destroy.argtypes = [c_vp]
        L.dre_replay_storage.argtypes = [c_vp, P(ReplayView), P(c_i32), P(ctypes.c_size_t)]
        L.dre_replay_insert.argtypes = [c_vp, c_vp, c_vp, c_i32, c_vp]
        L.dre_replay_count.argtypes = [c_vp, P(c_i64), P(c_i64), P(c_i64), c_vp]
        L.dre_replay_clear.argtypes = [c_vp, c_vp]
        L.dre_replay_sample.argtypes = [c_vp, c_vp, c_i64, c_i32, P(ReplayView), c_vp]
        L.dre_env_set_profiling.argtypes = [c_vp, c_i32]
        L.dre_env_kernel_ms.argtypes = [c_vp, c_vp, c_vp, c_i32]
    L.dre_measure_peak.argtypes = [c_i32, P(c_d)]
    L.dre_mpc_math_selftest.argtypes = [c_i32, c_vp, c_vp, c_vp]
    L.dre_mpc_math_selftest_fn.argtypes = [c_i32, c_i32, c_vp, c_vp, c_vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        L = load()
        msg = L.dre_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.dre_last_cuda_error().decode()
        raise DreError(f"libdre call failed ({rc}): {msg}")
