"""Error behaviour of the C ABI (include/dre.h): return codes instead of exceptions, argument validation before any
launch, call-order checks (step before reset), dre_strerror text. Mirrors how the reference reports misuse: config
asserts and NotImplementedError (config_general.py:34,37; HA_and_PT env.py:282-283) rather than silent fallbacks."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_invalid_arguments_and_call_order():
    import torch
    import dr_mpc_b200 as d
    from dr_mpc_b200 import _lib
    L = _lib.load()
    torch.zeros(1, device="cuda:0")
    h = ctypes.c_void_p()
    # --- creation-time validation ---
    for bad in (dict(num_envs=0), dict(num_humans=0), dict(num_humans=65), dict(lookback=0), dict(horizon=41),
                dict(soft_reset_on_device=True, auto_path_switch=False), dict(num_humans=3, static_humans=((0, 0), (1, 1), (2, 2)))):
        c = d.EngineConfig(**{"num_envs": 8, **bad}).env_config()
        assert L.dre_env_create(ctypes.byref(c), ctypes.byref(h)) == -1, bad
    assert L.dre_strerror(-1).decode() and L.dre_strerror(-4).decode() == "call order violated"
    # --- call order: paths, then reset, then step ---
    c = d.EngineConfig(num_envs=8, num_humans=4).env_config()
    assert L.dre_env_create(ctypes.byref(c), ctypes.byref(h)) == 0
    a = torch.zeros((8, 2), dtype=torch.float64, device="cuda:0")
    assert L.dre_env_reset(h, None, None, None) == -4                       # no paths yet
    tab, lens = d.paths.pack_paths(d.paths.continuous_task_paths())
    assert L.dre_env_set_paths(h, tab.ctypes.data_as(ctypes.c_void_p), 4, tab.shape[2], lens.ctypes.data_as(ctypes.c_void_p)) == 0
    assert L.dre_env_step(h, ctypes.c_void_p(a.data_ptr()), None) == -4     # step before reset
    assert L.dre_env_step(h, None, None) == -1
    host = np.zeros(16)
    assert L.dre_env_step_host(h, host.ctypes.data_as(ctypes.c_void_p), host.ctypes.data_as(ctypes.c_void_p)) == -4
    assert L.dre_env_reset(h, None, None, None) == 0
    assert L.dre_env_step(h, ctypes.c_void_p(a.data_ptr()), None) == 0
    assert L.dre_env_set_host_chunks(h, 0) == -1 and L.dre_env_set_host_chunks(h, 33) == -1 and L.dre_env_set_host_chunks(h, 4) == 0
    out = torch.zeros((8, 2), dtype=torch.float64, device="cuda:0")
    bk = torch.zeros((40, 2), dtype=torch.float64, device="cuda:0")
    assert L.dre_env_cvmm_filter(h, ctypes.c_void_p(a.data_ptr()), 8, ctypes.c_void_p(bk.data_ptr()), 40, ctypes.c_void_p(out.data_ptr()), None, None, None) == -1
    assert L.dre_env_cvmm_filter(h, ctypes.c_void_p(a.data_ptr()), 0, ctypes.c_void_p(bk.data_ptr()), 14, ctypes.c_void_p(out.data_ptr()), None, None, None) == -1
    torch.cuda.synchronize()
    assert L.dre_env_destroy(h) == 0
    # --- MPC object ---
    m = ctypes.c_void_p()
    assert L.dre_mpc_create(ctypes.byref(_lib.MpcParams(0, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)), 8, ctypes.byref(m)) == -1
    assert L.dre_mpc_create(ctypes.byref(_lib.MpcParams(4, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)), 0, ctypes.byref(m)) == -1
    assert L.dre_mpc_create(ctypes.byref(_lib.MpcParams(4, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)), 8, ctypes.byref(m)) == 0
    bad_lens = np.array([1, 41, 125, 41], dtype=np.int32)                  # a path needs >= 2 nodes
    assert L.dre_mpc_set_paths(m, tab.ctypes.data_as(ctypes.c_void_p), 4, tab.shape[2], bad_lens.ctypes.data_as(ctypes.c_void_p)) == -1
    assert L.dre_mpc_destroy(m) == 0


def test_python_mirror_raises_on_failure():
    import dr_mpc_b200 as d
    from dr_mpc_b200 import _lib
    with pytest.raises(_lib.DreError):
        d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=65), num_envs=4)


def test_env_does_not_modify_the_callers_config():
    """Two envs built from ONE EngineConfig with different seed additions / sizes: the object the caller holds keeps its
    values (the reference builds several envs from one ConfigMaster the same way, online_continuous_task.py:39)."""
    import dr_mpc_b200 as d
    cfg = d.EngineConfig(num_humans=5)
    seed, n = cfg.seed, cfg.num_envs
    a = d.BatchedHAAndPTEnv(cfg, num_envs=8, seed_addition=3)
    b = d.BatchedHAAndPTEnv(cfg, num_envs=16, seed_addition=3, soft_reset_on_device=True)
    assert (cfg.seed, cfg.num_envs, cfg.soft_reset_on_device) == (seed, n, False)
    assert a.cfg.seed == b.cfg.seed == seed + 3 and (a.E, b.E) == (8, 16) and b.cfg.soft_reset_on_device
