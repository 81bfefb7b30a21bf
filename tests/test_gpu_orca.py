"""GPU parity: the CUDA ORCA pass (through the C ABI) against the reference golden vectors and the C oracle.
fp32 on both sides, FMA contraction off, same operation order -> bit-exact; LP-branch divergences are counted."""
import numpy as np
import pytest
from conftest import ORCA_KEYS

pytestmark = pytest.mark.gpu


def _engine(group_lanes=0, **kw):
    import dr_mpc_b200 as d
    return d.BatchedORCA(d.EngineConfig(group_lanes=group_lanes, **kw))


@pytest.mark.parametrize("lanes", [4, 8, 16, 32, 1, 1 | 0x400])   # 1: thread-per-agent variant (0x400: projected lines in global scratch)
@pytest.mark.parametrize("key", ORCA_KEYS)
def test_cuda_matches_reference_golden_bit_exact(golden_orca, key, lanes):
    g = golden_orca
    orca = _engine(lanes)
    v = orca.predict(g[key + "_hpos"], g[key + "_hvel"], g[key + "_goal"], g[key + "_rpos"])      # host-buffer C-ABI entry
    ref = g[key + "_vnew"]
    assert np.array_equal(v.view(np.uint32), ref.view(np.uint32)), np.abs(v - ref).max()


@pytest.mark.parametrize("E,Nh,spread", [(1024, 10, 4.0), (2048, 20, 5.0), (512, 50, 8.0), (4096, 6, 1.5), (333, 1, 1.0), (7, 63, 6.0), (5, 64, 6.0), (1, 2, 0.4)])
def test_cuda_matches_oracle_random_crowds(oracle, E, Nh, spread):
    import torch
    rng = np.random.default_rng(E + Nh)
    ang = rng.uniform(0, 2 * np.pi, (E, Nh)); rad = spread * np.sqrt(rng.uniform(0, 1, (E, Nh)))
    hpos = np.stack([rad * np.cos(ang), rad * np.sin(ang)], -1)
    goal = np.where(rng.uniform(size=(E, 1, 1)) < 0.5, -hpos, hpos + rng.uniform(-0.8, 0.8, (E, Nh, 2)))
    hvel = (rng.uniform(-1, 1, (E, Nh, 2)) * rng.uniform(0, 1, (E, Nh, 1))).astype(np.float32)
    rpos = rng.uniform(-spread, spread, (E, 2))
    static = rng.uniform(size=(E, Nh)) < 0.05
    v_ref, st_ref = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, is_static=static, want_stats=True)
    orca = _engine()
    dev = "cuda:0"
    v = orca.predict(torch.tensor(hpos, device=dev), torch.tensor(hvel, device=dev), torch.tensor(goal, device=dev),
                     torch.tensor(rpos, device=dev), is_static=torch.tensor(static, device=dev), want_stats=True)
    torch.cuda.synchronize()
    v = v.cpu().numpy(); st = orca.last_stats.cpu().numpy()
    mism = (v.view(np.uint32) != v_ref.view(np.uint32)).any(-1)
    branch_div = (st[..., 1] != st_ref["lp2_fail"]) | (st[..., 2] != st_ref["n_lp1"]) | (st[..., 4].view(np.uint32) != st_ref["branch_mask"])
    print(f"E={E} Nh={Nh}: velocity mismatches {int(mism.sum())}/{mism.size}, LP-branch divergences {int(branch_div.sum())}, "
          f"LP3 solves {int((st_ref['lp2_fail'] < st_ref['n_lines']).sum())}")
    assert mism.sum() == 0 and branch_div.sum() == 0
    assert (v[static] == 0).all()


def test_max_neighbors_and_neighbor_dist_cutoff(oracle):
    """neighbor_dist = 3 (grid-cutoff case of SURVEY 8(d) C4) and a binding max_neighbors."""
    import dr_mpc_b200 as d
    rng = np.random.default_rng(3)
    E, Nh = 256, 30
    hpos = rng.uniform(-6, 6, (E, Nh, 2)); goal = -hpos
    hvel = rng.uniform(-0.7, 0.7, (E, Nh, 2)).astype(np.float32)
    rpos = rng.uniform(-6, 6, (E, 2))
    v_ref = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, neighbor_dist=3.0)
    v = d.BatchedORCA(d.EngineConfig(neighbor_dist=3.0)).predict(hpos, hvel, goal, rpos)
    assert np.array_equal(v.view(np.uint32), v_ref.view(np.uint32))


def test_full_size_properties():
    """C3/C4 sizes: properties that need no oracle -- determinism, speed bound, lone agents head for their goal,
    translation invariance of the whole crowd (positions and goals shifted by an fp32-exact offset)."""
    import torch
    import dr_mpc_b200 as d
    dev = "cuda:0"
    for E, Nh in ((16384, 20), (8192, 50)):
        gen = torch.Generator(device=dev).manual_seed(E)
        hpos = (torch.rand((E, Nh, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 12
        hpos = hpos.float().double()                                   # fp32-representable so a shift by 64 stays exact-ish
        goal = -hpos
        hvel = (torch.rand((E, Nh, 2), generator=gen, device=dev) - 0.5)
        rpos = (torch.rand((E, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 12
        orca = d.BatchedORCA(d.EngineConfig())
        v1 = orca.predict(hpos, hvel, goal, rpos, want_stats=True)
        st = orca.last_stats
        v2 = orca.predict(hpos, hvel, goal, rpos)
        assert torch.equal(v1, v2)
        speed = v1.double().norm(dim=-1)
        feasible = st[..., 1] == st[..., 0]
        # LP2-feasible results lie in the max-speed disc. LP3 (infeasible) results can leave it: with two almost
        # anti-parallel half-planes RVO2's fp32 discriminant in linearProgram1 cancels catastrophically (bit-exactly
        # reproduced, see tests/test_oracle_orca.py::test_lp3_overshoot_is_reference_behaviour); it must stay rare.
        assert torch.isfinite(v1).all() and (speed[feasible] <= 1.0 + 1e-5).all()
        assert (speed > 1.01).double().mean() < 1e-3
    # lone humans far from everything: new velocity == preferred velocity
    E, Nh = 4096, 4
    hpos = torch.zeros((E, Nh, 2), device=dev, dtype=torch.float64)
    hpos[:, :, 0] = torch.arange(Nh, device=dev) * 100.0
    goal = hpos + torch.tensor([3.0, 4.0], device=dev, dtype=torch.float64)
    v = d.BatchedORCA(d.EngineConfig()).predict(hpos, torch.zeros((E, Nh, 2), device=dev), goal,
                                                 torch.full((E, 2), 1e4, device=dev, dtype=torch.float64))
    assert torch.allclose(v, torch.tensor([0.6, 0.8], device=dev).expand_as(v), atol=1e-7)


def test_full_size_symmetries_are_bit_exact():
    """C3 / C4 shard sizes, no oracle needed: the solve commutes with quarter turns, reflections and permutations of the humans
    BIT FOR BIT (two-term dot products / determinants, norms and comparisons map to themselves in IEEE arithmetic, neighbours
    are ranked by distance; tests/test_oracle_shims.py holds the same property for the CPU restatement), LP3 solves included."""
    import torch
    import dr_mpc_b200 as d
    dev = "cuda:0"
    quarter = lambda a: torch.stack([-a[..., 1], a[..., 0]], -1).contiguous()
    mirror = lambda a: (a * torch.tensor([1, -1], device=dev, dtype=a.dtype)).contiguous()
    for E, Nh, spread in ((16384, 20, 5.0), (8192, 50, 8.0)):
        gen = torch.Generator(device=dev).manual_seed(7 * E + Nh)
        hpos = (torch.rand((E, Nh, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 2 * spread
        goal = (torch.rand((E, Nh, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 2 * spread
        hvel = (torch.rand((E, Nh, 2), generator=gen, device=dev) - 0.5) * 1.4
        rpos = (torch.rand((E, 2), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 2 * spread
        orca = d.BatchedORCA(d.EngineConfig())
        v0 = orca.predict(hpos, hvel, goal, rpos, want_stats=True)
        st = orca.last_stats
        assert (st[..., 1] < st[..., 0]).double().mean() > 0.1          # the infeasible (LP3) branch is well represented
        for name, g in (("quarter turn", quarter), ("half turn", lambda a: (-a).contiguous()), ("reflection", mirror)):
            v = orca.predict(g(hpos), g(hvel), g(goal), g(rpos))
            assert torch.equal(v, g(v0)), (E, Nh, name)              # exact float equality (a zero may change its sign)
        perm = torch.randperm(Nh, generator=torch.Generator().manual_seed(Nh)).to(dev)
        v = orca.predict(hpos[:, perm].contiguous(), hvel[:, perm].contiguous(), goal[:, perm].contiguous(), rpos)
        # same bits, except where two neighbours sit at EXACTLY the same fp32 distance: RVO2's insertNeighbor then keeps index
        # order, which a permutation changes (a dozen such agents are expected among 8192 x 50)
        differ = (v.view(torch.int32) != v0[:, perm].contiguous().view(torch.int32)).any(-1).double().mean().item()
        assert differ < 1e-4, (E, Nh, "permutation", differ)
