"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the reference's OWN, UNMODIFIED Python
(/root/reference: scripts/policy/orca.py, environment/path_tracking/utils/mpc.py, pysteam_augmented/*, the env
classes) on the API shims in oracle/shims. Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The GPU box never runs this; it only reads the committed fixtures.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("DRE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def _config(num_humans=6):
    from scripts.configs import ConfigMaster
    cm = ConfigMaster()
    cm.config_HA.sim.human_num = num_humans
    cm.config_HA.sim.max_allowable_humans = num_humans
    return cm


# ------------------------------------------------------------------------------------------------------------
# 1. ORCA: CrowdSim.get_human_actions on hand-built crowds (crowd_sim.py:264-284 -> human.py:16-25 -> orca.py:64-117)
# ------------------------------------------------------------------------------------------------------------
def golden_orca():
    from environment.human_avoidance.human_avoidance_env import HAEnv
    from environment.human_avoidance.utils.human import Human
    rng = np.random.default_rng(20241019)
    out = {}
    for Nh, spread, C in ((1, 2.0, 24), (6, 3.0, 48), (6, 1.2, 48), (10, 4.0, 32), (10, 1.8, 32), (20, 5.0, 24),
                          (20, 2.5, 24), (50, 8.0, 8), (50, 4.5, 8)):
        cm = _config(Nh)
        env = HAEnv()
        env.configure(cm, cm.config_HA)
        hpos = np.zeros((C, Nh, 2)); hvel = np.zeros((C, Nh, 2), dtype=np.float32); goal = np.zeros((C, Nh, 2))
        rpos = np.zeros((C, 2)); vnew = np.zeros((C, Nh, 2), dtype=np.float32)
        for c in range(C):
            ang = rng.uniform(0, 2 * np.pi, Nh); rad = spread * np.sqrt(rng.uniform(0, 1, Nh))
            P = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
            V = (rng.uniform(-1, 1, (Nh, 2)) * rng.uniform(0, 1, (Nh, 1))).astype(np.float32)
            if c % 4 == 0:
                G = P + rng.uniform(-0.6, 0.6, (Nh, 2))          # goal closer than 1 m: un-normalised pref velocity
            elif c % 4 == 1:
                G = -P                                           # circle crossing
            else:
                G = rng.uniform(-6, 6, (Nh, 2))
            if c % 5 == 0 and Nh >= 2:
                P[1] = P[0] + np.array([0.5, 0.1])               # overlapping pair -> collision branch
            if c % 7 == 0 and Nh >= 3:
                V[:] = 0                                         # everybody at rest
            R = rng.uniform(-spread, spread, 2)
            env.robot.set(R[0], R[1], 0, 4, 0, 0, 0.3)
            env.humans = []
            for i in range(Nh):
                h = Human(cm.config_HA, 'humans')
                h.set(P[i, 0], P[i, 1], G[i, 0], G[i, 1], float(V[i, 0]), float(V[i, 1]), 0)
                env.humans.append(h)
            acts = env.get_human_actions()
            hpos[c], hvel[c], goal[c], rpos[c] = P, V, G, R
            vnew[c] = np.array([[a.vx, a.vy] for a in acts], dtype=np.float32)
        key = f"n{Nh}_s{spread}"
        out[key + "_hpos"], out[key + "_hvel"], out[key + "_goal"], out[key + "_rpos"], out[key + "_vnew"] = hpos, hvel, goal, rpos, vnew
    np.savez_compressed(os.path.join(OUT, "orca_crowds.npz"), **out)
    print("orca_crowds.npz", {k: v.shape for k, v in out.items() if k.endswith("_vnew")})


# ------------------------------------------------------------------------------------------------------------
# 2. MPC: MPC.act sequences (mpc.py:58-206) along the four continuous-task paths
# ------------------------------------------------------------------------------------------------------------
def _warm_of(mpc, K):
    """(warm_T [K,4] = c,s,x,y of the stored inverse poses, warm_u [K,2], iteration0) before an act() call."""
    T = np.zeros((K, 4)); u = np.zeros((K, 2))
    if not mpc.iteration0 and mpc.next_se3_state_vars != 0:
        for k in range(K):
            M = mpc.next_se3_state_vars[k].value.matrix()
            T[k] = [M[0, 0], M[1, 0], M[0, 3], M[1, 3]]
            u[k] = mpc.next_vspace_state_vars[k].value[:, 0]
    return T, u, int(mpc.iteration0)


def golden_mpc():
    from environment.path_tracking.utils.mpc import MPC
    from environment.path_tracking.path_tracking_env import PTEnv
    from environment.path_tracking.utils import RewardComputationPT, Unicycle
    from pysteam.pysteam.solver import solver as SV
    SV.TRACE = []
    cm = _config()
    pt = PTEnv(cm, RewardComputationPT(cm.config_PT.rewards, True))
    pt.reset(np.array([0, -4.0, np.pi]))
    out = {f"path{i}": p for i, p in enumerate(pt.paths)}
    for K, steps in ((4, 70), (10, 24), (20, 24), (40, 16)):
        rec = {k: [] for k in ("path_id", "interp", "pose", "warm_T", "warm_u", "it0", "u", "iters", "cost", "retries")}
        for pid in range(4):
            mpc = MPC(cm, K, 0.25, True, 1.0, 1.0)
            rng = np.random.default_rng(100 * K + pid)
            path = pt.paths[pid]
            pose = path[:, 0] + np.array([0.1, -0.2, 0.15])
            for step in range(steps):
                if step == steps // 2:
                    mpc.reset()                                       # iteration0 again mid-sequence
                    pose = pose + np.array([0.25, 0.2, -0.3])
                _, interp_idx, _ = pt.localize_on_path(pt.kd_trees[pid], pose[:2], path, localize_to_start=(step < 20))
                wT, wu, it0 = _warm_of(mpc, K)
                SV.TRACE.clear()
                u = mpc.act({'path': path, 'interp_index': interp_idx, 'pose': pose})
                rec["path_id"].append(pid); rec["interp"].append(float(interp_idx)); rec["pose"].append(pose.copy())
                rec["warm_T"].append(wT); rec["warm_u"].append(wu); rec["it0"].append(it0); rec["u"].append(u.copy())
                rec["iters"].append(SV.TRACE[-1][0]); rec["cost"].append(SV.TRACE[-1][1]); rec["retries"].append(len(SV.TRACE) - 1)
                a = u[:2] + rng.normal(0, 0.05, 2)
                a[0] = np.clip(a[0], 0, 1); a[1] = np.clip(a[1], -1, 1)
                pose = Unicycle.step_external(pose, a, 0.25)
        for k, v in rec.items():
            out[f"K{K}_{k}"] = np.array(v)
    # ---- forced failures: solves whose LM step fails, so that the reference's own `except: pose_error_scaling *= 5;
    # continue` (mpc.py:188-190) runs. Candidates are found with the flat oracle (test infrastructure too), then EVERY
    # recorded number comes from the reference's MPC.act. Two kinds per path:
    #   "stale": solve, reset(), then a failing FIRST solve -> the retry deep-copies the warm start stored before the reset
    #            (mpc.py:72-75: iteration0 is cleared before the solve)
    #   "warm" : solve, then a far-away pose -> a failing warm-started solve retries from the same warm start
    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_oracle_front", os.path.join(HERE, "oracle.py"))
    O = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(O)
    O.build()
    stride = max(p.shape[1] for p in pt.paths)
    tab = np.zeros((4, 3, stride)); lens = np.zeros(4, dtype=np.int32)
    for i, p in enumerate(pt.paths):
        tab[i, :, :p.shape[1]] = p; lens[i] = p.shape[1]
    for K, per_kind in ((4, 4), (10, 2)):
        rec = {k: [] for k in ("path_id", "interp", "pose", "warm_T", "warm_u", "it0", "u", "iters", "cost", "retries", "kind")}
        for pid in range(4):
            path = pt.paths[pid]
            start_pose = path[:, 5] + np.array([0.05, -0.05, 0.05])
            _, start_idx, _ = pt.localize_on_path(pt.kd_trees[pid], start_pose[:2], path, localize_to_start=True)

            def armed(reset):
                m = MPC(cm, K, 0.25, True, 1.0, 1.0)
                m.act({'path': path, 'interp_index': start_idx, 'pose': start_pose})
                if reset:
                    m.reset()
                return m
            for kind, reset in (("stale", True), ("warm", False)):
                wT, wu, _ = _warm_of(armed(False), K)
                rng = np.random.default_rng(7000 + 10 * K + pid + (100 if reset else 0))
                N = 3000
                node = rng.integers(0, path.shape[1] - 1, N)
                cand = path[:, node].T + rng.uniform(-0.7, 0.7, (N, 3))
                interp = np.array([pt.localize_on_path(pt.kd_trees[pid], c[:2], path, localize_to_start=False)[1] for c in cand], dtype=np.float64)
                mb = O.MpcBatch(N, O.mpc_params(K=K), tab, lens)
                mb.warm_T[:] = wT; mb.warm_u[:] = wu; mb.iteration0[:] = 1 if reset else 0
                _, st = mb.act(np.full(N, pid, np.int32), interp, cand)
                picks = np.nonzero(st["retries"] > 0)[0][:per_kind]
                assert len(picks) > 0, (K, pid, kind)
                for i in picks:
                    m = armed(reset)
                    wT_i, wu_i, it0_i = _warm_of(m, K)
                    if reset:   # _warm_of reports zeros while iteration0 is set; the stale warm start is still inside the object
                        for k in range(K):
                            Mx = m.next_se3_state_vars[k].value.matrix()
                            wT_i[k] = [Mx[0, 0], Mx[1, 0], Mx[0, 3], Mx[1, 3]]
                            wu_i[k] = m.next_vspace_state_vars[k].value[:, 0]
                    SV.TRACE.clear()
                    u = m.act({'path': path, 'interp_index': float(interp[i]), 'pose': cand[i]})
                    assert len(SV.TRACE) >= 2, "the reference did not take its retry path on this candidate"
                    rec["path_id"].append(pid); rec["interp"].append(float(interp[i])); rec["pose"].append(cand[i].copy())
                    rec["warm_T"].append(wT_i); rec["warm_u"].append(wu_i); rec["it0"].append(it0_i); rec["u"].append(u.copy())
                    rec["iters"].append(SV.TRACE[-1][0]); rec["cost"].append(SV.TRACE[-1][1]); rec["retries"].append(len(SV.TRACE) - 1)
                    rec["kind"].append(0 if reset else 1)
        for k, v in rec.items():
            out[f"F{K}_{k}"] = np.array(v)
        print(f"forced failures K={K}:", len(rec["u"]), "solves, retries", rec["retries"])
    SV.TRACE = None
    np.savez_compressed(os.path.join(OUT, "mpc_sequences.npz"), **out)
    print("mpc_sequences.npz", {k: v.shape for k, v in out.items() if k.endswith("_u")})


# ------------------------------------------------------------------------------------------------------------
# 3. Full env: HAAndPTEnv.reset/step (HA_and_PT env.py:243-416) teacher-forcing fixtures
# ------------------------------------------------------------------------------------------------------------
def _snapshot(env, K):
    ha, pt, mpc = env.HA_env, env.PT_env, env.MPC
    Nh = len(ha.humans)
    s = {
        "hpos": np.array([[h.px, h.py] for h in ha.humans]), "hvel": np.array([[h.vx, h.vy] for h in ha.humans], dtype=np.float32),
        "hgoal": np.array([[h.gx, h.gy] for h in ha.humans]), "hstatic": np.array([bool(h.isObstacle) for h in ha.humans], dtype=np.uint8),
        "hhist": ha.human_history_global.copy(), "hvalid": ha.human_valid_history[:Nh].astype(np.int32),
        "rpose": np.array([env.robot.px, env.robot.py, env.robot.theta, 0.0]),
        "rprev": np.concatenate([env.prev_robot_pose, [0.0]]),
        "rhist": np.concatenate([ha.robot_pose_history_global, np.zeros((ha.lookback + 1, 1))], 1),
        "rpastv": ha.robot_past_velocities.copy(), "rvalid": np.int32(ha.robot_valid_history),
        "gtime": np.float64(ha.global_time), "path_idx": np.int32(pt.path_idx), "loc_to_start": np.uint8(pt.localize_to_start),
    }
    s["warm_T"], s["warm_u"], s["it0"] = _warm_of(mpc, K)
    return s


def _flatten_S(S):
    return {"pt_state": S['PT']['PT_state'][0], "mpc_actions": S['PT']['MPC_actions'][0], "robot_node": S['HA']['robot_node'][0],
            "past_robot_vel": S['HA']['past_robot_velocities'][0], "percent_episode": S['HA']['percent_episode'][0],
            "spatial_edges": S['HA']['spatial_edges'][0], "human_valid": S['HA']['human_valid_history'][0],
            "detected_humans": np.float64(S['HA']['detected_human_num'][0])}


_BITS = {'collision': 1 << 0, 'safety_human_raise': 1 << 1, 'actuation_termination': 1 << 2, 'timeout': 1 << 3,
         'goal': 1 << 4, 'deviation_end_of_path': 1 << 5, 'corridor_hit': 1 << 6, 'safety_corridor_raise': 1 << 7}


def golden_env():
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot
    out = {}
    for tag, Nh, seed, steps in (("h6", 6, 0, 500), ("h10", 10, 1, 90), ("h20", 20, 2, 50)):
        cm = _config(Nh)
        env = HAAndPTEnv(cm, seed_addition=seed)
        env.case_counter['train'] = seed
        K = cm.config_PT.MPC.planning_horizon
        S = env.reset()
        rng = np.random.default_rng(seed)
        rec_pre, rec_post, rec_out = [], [], []
        out[f"{tag}_reset_hpos"] = None
        # the spawn (needed to replay reset() on the device): positions/goals before warm start are not kept by the
        # reference, so re-run the spawning with the same seed
        np.random.seed(seed)
        env2 = HAAndPTEnv(cm, seed_addition=seed)
        env2.case_counter['train'] = seed
        env2.config_HA.sim.warm_start = False
        S2 = env2.reset()
        env2.config_HA.sim.warm_start = True
        out[f"{tag}_reset_hpos"] = np.array([[h.px, h.py] for h in env2.HA_env.humans])
        out[f"{tag}_reset_hgoal"] = np.array([[h.gx, h.gy] for h in env2.HA_env.humans])
        for k, v in _flatten_S(S).items():
            out[f"{tag}_reset_{k}"] = np.asarray(v)
        post_reset = _snapshot(env, K)
        for k, v in post_reset.items():
            out[f"{tag}_postreset_{k}"] = np.asarray(v)
        done_return, info = {'done': False}, None
        n = 0
        while n < steps:
            if done_return['done']:
                S = env.soft_reset(done_return, info, S)
                if S is None:
                    break
            a = np.array(S['PT']['MPC_actions'][0, :2], dtype=np.float64)
            # perturb the MPC action now and then so that the robot meets humans, leaves the corridor, stalls ...
            phase = (n // 15) % 4
            if phase == 1:
                a = a + rng.normal(0, 0.3, 2)
            elif phase == 3 and Nh == 6:
                a = np.array([0.0, 0.0]) if (n // 5) % 2 else a * 0.3
            a[0] = float(np.clip(a[0], 0, 1)); a[1] = float(np.clip(a[1], -1, 1))
            pre = _snapshot(env, K)
            S, R, D, info, eps = env.step(ActionRot(a[0], a[1]))
            done_return = D
            post = _snapshot(env, K)
            bits = sum(b for name, b in _BITS.items() if name in info['done'])
            o = _flatten_S(S)
            o.update({"action": a, "rewards": np.array([R['R_PT'], R['R_DO'], R['R']]), "bits": np.int64(bits),
                      "done_PT": np.uint8(D['done_PT']), "done_HA": np.uint8(D['done_HA']),
                      "disturbance_v": np.float64(info['disturbance_v']), "disturbance_th": np.float64(info['disturbance_th']),
                      "path_advancement": np.float64(info['path_advancement']), "deviation": np.float64(info['deviation'])})
            rec_pre.append(pre); rec_post.append(post); rec_out.append(o)
            n += 1
        for name, rec in (("pre", rec_pre), ("post", rec_post), ("out", rec_out)):
            for k in rec[0]:
                out[f"{tag}_{name}_{k}"] = np.stack([np.asarray(r[k]) for r in rec])
        allbits = out[f"{tag}_out_bits"]
        print(tag, "steps", len(rec_pre), "done events", {k: int(((allbits & b) != 0).sum()) for k, b in _BITS.items()})
    np.savez_compressed(os.path.join(OUT, "env_rollouts.npz"), **out)
    print("env_rollouts.npz", os.path.getsize(os.path.join(OUT, "env_rollouts.npz")), "bytes")


# ------------------------------------------------------------------------------------------------------------
# 4. HAAndPTEnv.soft_reset (HA_and_PT env.py:120-240): EVERY env.step call of a run, including the scripted ones the
#    reference issues from inside soft_reset(), with the action it chose. env.step is wrapped on the instance (the
#    reference code is untouched) so the calls made by soft_reset() are seen too.
# ------------------------------------------------------------------------------------------------------------
def golden_soft_reset():
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot
    out = {}
    for tag, Nh, seed, policy_steps in (("h6", 6, 3, 420), ("h10", 10, 4, 160), ("h6s", 6, 5, 60)):
        cm = _config(Nh)
        env = HAAndPTEnv(cm, seed_addition=seed)
        env.case_counter['train'] = seed
        K = cm.config_PT.MPC.planning_horizon
        S = env.reset()
        rng = np.random.default_rng(seed)
        if tag == "h6s":
            # a static human (crowd_sim.py:270-272) parked on the path: the robot runs into its safety range, waits more
            # than 10 steps, backs up (env.py:187-190) and takes the 7 extra back-up steps (:227-233)
            h0 = env.HA_env.humans[0]
            node = env.PT_env.paths[0][:, 22]
            h0.isObstacle = True
            h0.px, h0.py, h0.vx, h0.vy, h0.gx, h0.gy = float(node[0]), float(node[1]), 0.0, 0.0, float(node[0]), float(node[1])
        rec = []
        inner = env.step

        def recording_step(action, update=True, soft_reset=False, model_info=None):
            pre = _snapshot(env, K)
            r = inner(action, update=update, soft_reset=soft_reset, model_info=model_info)
            S2, R, D, info, eps = r
            bits = sum(b for name, b in _BITS.items() if name in info['done'])
            rec.append({"pre": pre, "post": _snapshot(env, K), "action": np.array([action.v, action.w], dtype=np.float64),
                        "soft": np.uint8(bool(soft_reset)), "bits": np.int64(bits), "done": np.uint8(D['done']),
                        "done_HA": np.uint8(D['done_HA']), "rewards": np.array([R['R_PT'], R['R_DO'], R['R']])})
            return r

        env.step = recording_step
        done_return, info = {'done': False}, None
        n = 0
        while n < policy_steps:
            if done_return['done']:
                S = env.soft_reset(done_return, info, S)
                if S is None:
                    break
            a = np.array(S['PT']['MPC_actions'][0, :2], dtype=np.float64)
            phase = (n // 20) % 6
            if phase == 1:
                a = a + rng.normal(0, 0.3, 2)
            elif phase == 3 and tag != "h6s":
                a = np.array([1.0, -0.35]) if (n // 120) % 2 == 0 else np.array([1.0, 0.5])   # leave the corridor
            elif phase == 5 and tag == "h6":
                a = np.array([0.0, 0.0]) if (n // 5) % 2 else a * 0.3                          # stall: actuation termination
            a[0] = float(np.clip(a[0], 0, 1)); a[1] = float(np.clip(a[1], -1, 1))
            S, R, D, info, eps = env.step(ActionRot(a[0], a[1]))
            done_return = D
            n += 1
        for k in rec[0]["pre"]:
            out[f"{tag}_pre_{k}"] = np.stack([np.asarray(r["pre"][k]) for r in rec])
            out[f"{tag}_post_{k}"] = np.stack([np.asarray(r["post"][k]) for r in rec])
        for k in ("action", "soft", "bits", "done", "done_HA", "rewards"):
            out[f"{tag}_{k}"] = np.stack([np.asarray(r[k]) for r in rec])
        soft = out[f"{tag}_soft"].astype(bool)
        acts = out[f"{tag}_action"][soft]
        kinds = {"wait": int(((acts[:, 0] == 0) & (acts[:, 1] == 0)).sum()), "backup": int((acts[:, 0] == -0.3).sum()),
                 "turn": int(((acts[:, 0] == 0) & (acts[:, 1] != 0)).sum()), "creep": int((acts[:, 0] == 0.5).sum())}
        allbits = out[f"{tag}_bits"]
        print(tag, "calls", len(rec), "scripted", int(soft.sum()), kinds,
              "done events", {k: int(((allbits & b) != 0).sum()) for k, b in _BITS.items()})
    np.savez_compressed(os.path.join(OUT, "soft_reset.npz"), **out)
    print("soft_reset.npz", os.path.getsize(os.path.join(OUT, "soft_reset.npz")), "bytes")


# ------------------------------------------------------------------------------------------------------------
# 5. CVMM safety filter: HAAndPTEnv.cvmm_safety_check / cvmm_diverse_safety_pipeline (HA_and_PT env.py:422-513) on the
#    states of a run, for the policy's action, a noisy one and an aggressive one.
# ------------------------------------------------------------------------------------------------------------
def golden_cvmm():
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot
    out = {}
    for tag, Nh, seed, steps in (("h6", 6, 6, 150), ("h10", 10, 7, 90)):
        cm = _config(Nh)
        action_set = cm.config_general.safety_module.params['action_set']
        nsteps = cm.config_general.safety_module.params['lookahead_steps']
        env = HAAndPTEnv(cm, seed_addition=seed)
        env.case_counter['train'] = seed
        K = cm.config_PT.MPC.planning_horizon
        S = env.reset()
        rng = np.random.default_rng(seed)
        rec = []
        done_return, info = {'done': False}, None
        for n in range(steps):
            if done_return['done']:
                S = env.soft_reset(done_return, info, S)
                if S is None:
                    break
            mpc_a = np.array(S['PT']['MPC_actions'][0, :2], dtype=np.float64)
            pre = _snapshot(env, K)
            for kind in range(3):
                if kind == 0:
                    a = mpc_a.copy()
                elif kind == 1:
                    a = np.clip(mpc_a + rng.normal(0, 0.35, 2), [0, -1], [1, 1])
                else:
                    a = np.array([rng.uniform(0.6, 1.0), rng.choice([-1.0, 1.0]) * rng.uniform(0.3, 1.0)])
                cand = [env.cvmm_safety_check(a, nsteps)] + [env.cvmm_safety_check(action_set[i], nsteps) for i in range(action_set.shape[0])]
                st = np.random.get_state()
                act, inf = env.cvmm_diverse_safety_pipeline(a, nsteps, action_set)
                np.random.set_state(st)      # the fall-back draw must not perturb the run's goal resampling
                rec.append({"pre": pre, "action": a, "cand_safe": np.array([c[0] for c in cand], dtype=np.uint8),
                            "cand_dist": np.array([c[1] for c in cand]), "safe_action": np.array([act.v, act.w]),
                            "random": np.uint8(inf['RL_action_bool'].startswith('no safe'))})
            a = mpc_a + (rng.normal(0, 0.25, 2) if (n // 12) % 2 else 0.0)
            S, R, done_return, info, eps = env.step(ActionRot(float(np.clip(a[0], 0, 1)), float(np.clip(a[1], -1, 1))))
        for k in rec[0]["pre"]:
            out[f"{tag}_pre_{k}"] = np.stack([np.asarray(r["pre"][k]) for r in rec])
        for k in ("action", "cand_safe", "cand_dist", "safe_action", "random"):
            out[f"{tag}_{k}"] = np.stack([np.asarray(r[k]) for r in rec])
        cs = out[f"{tag}_cand_safe"]
        print(tag, "queries", len(rec), "original safe", int(cs[:, 0].sum()), "none safe", int(out[f"{tag}_random"].sum()),
              "mean safe back-ups", float(cs[:, 1:].sum(1).mean()))
    out["action_set"] = np.asarray(action_set, dtype=np.float64)
    out["lookahead_steps"] = np.int64(nsteps)
    np.savez_compressed(os.path.join(OUT, "cvmm.npz"), **out)
    print("cvmm.npz", os.path.getsize(os.path.join(OUT, "cvmm.npz")), "bytes")


# ------------------------------------------------------------------------------------------------------------
# 6. Robot sensor restriction (config_HA.robot.sensor_restriction = True, sensor_range 4 m): HAAndPTEnv.reset / step with
#    humans dropping out of and coming back into range (human_avoidance_env.py:53-72,102-105,161-164,331-334)
# ------------------------------------------------------------------------------------------------------------
def golden_sensor():
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot
    out = {}
    for tag, Nh, seed, steps in (("h6", 6, 8, 150), ("h10", 10, 9, 90)):
        cm = _config(Nh)
        cm.config_HA.robot.sensor_restriction = True
        env = HAAndPTEnv(cm, seed_addition=seed)
        env.case_counter['train'] = seed
        K = cm.config_PT.MPC.planning_horizon
        S = env.reset()
        for k, v in _flatten_S(S).items():
            out[f"{tag}_reset_{k}"] = np.asarray(v)
        np.random.seed(seed)
        env2 = HAAndPTEnv(cm, seed_addition=seed)
        env2.case_counter['train'] = seed
        env2.config_HA.sim.warm_start = False
        env2.reset()
        env2.config_HA.sim.warm_start = True
        out[f"{tag}_reset_hpos"] = np.array([[h.px, h.py] for h in env2.HA_env.humans])
        out[f"{tag}_reset_hgoal"] = np.array([[h.gx, h.gy] for h in env2.HA_env.humans])
        rng = np.random.default_rng(seed)
        rec_pre, rec_post, rec_out = [], [], []
        done_return, info = {'done': False}, None
        for n in range(steps):
            if done_return['done']:
                S = env.soft_reset(done_return, info, S)
                if S is None:
                    break
            a = np.array(S['PT']['MPC_actions'][0, :2], dtype=np.float64)
            if (n // 15) % 3 == 1:
                a = a + rng.normal(0, 0.3, 2)
            a[0] = float(np.clip(a[0], 0, 1)); a[1] = float(np.clip(a[1], -1, 1))
            pre = _snapshot(env, K)
            S, R, done_return, info, eps = env.step(ActionRot(a[0], a[1]))
            o = _flatten_S(S)
            o.update({"action": a, "rewards": np.array([R['R_PT'], R['R_DO'], R['R']])})
            rec_pre.append(pre); rec_post.append(_snapshot(env, K)); rec_out.append(o)
        for name, rec in (("pre", rec_pre), ("post", rec_post), ("out", rec_out)):
            for k in rec[0]:
                out[f"{tag}_{name}_{k}"] = np.stack([np.asarray(r[k]) for r in rec])
        det = out[f"{tag}_out_detected_humans"]
        hv = out[f"{tag}_post_hvalid"]
        print(tag, "steps", len(rec_pre), "detected humans min/mean/max", det.min(), det.mean(), det.max(), "re-entries",
              int(((hv[1:] == 1) & (hv[:-1] == 0)).sum()), "nobody visible", int((hv.sum(1) == 0).sum()))
    np.savez_compressed(os.path.join(OUT, "sensor_restriction.npz"), **out)
    print("sensor_restriction.npz", os.path.getsize(os.path.join(OUT, "sensor_restriction.npz")), "bytes")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["orca", "mpc", "env", "soft_reset", "cvmm", "sensor"]
    if "orca" in which:
        golden_orca()
    if "mpc" in which:
        golden_mpc()
    if "env" in which:
        golden_env()
    if "soft_reset" in which:
        golden_soft_reset()
    if "cvmm" in which:
        golden_cvmm()
    if "sensor" in which:
        golden_sensor()
