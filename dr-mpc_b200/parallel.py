"""Multi-GPU plumbing: one process per GPU, environments sharded in contiguous blocks, NO collective on the step
path (envs never interact: reference HAAndPTEnv holds all state per instance). The only exchange is one all-reduce
(sum) of the episode statistics -- the 8 `eps_done_info` counters of reference HA_and_PT env.py:391-404 plus reward
sums and solver counters -- per reporting interval, over NCCL (NVLink 5 / NVSwitch) on GPUs or gloo on CPU."""
import os

STAT_NAMES = ['steps', 'success', 'collision', 'safety_human_raise', 'timeout', 'deviation_end_of_path', 'corridor_hit',
              'safety_corridor_raise', 'actuation_termination', 'mpc_retries', 'mpc_gave_up', 'orca_lp3']
_BIT_OF = {'success': 1 << 4, 'collision': 1 << 0, 'safety_human_raise': 1 << 1, 'timeout': 1 << 3,
           'deviation_end_of_path': 1 << 5, 'corridor_hit': 1 << 6, 'safety_corridor_raise': 1 << 7,
           'actuation_termination': 1 << 2}


def shard_envs(total_envs, world_size, rank):
    """Contiguous block of envs owned by `rank`: (global offset, count). The first `total % world` ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world_size, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def episode_stats_from_step(done_bits, rewards, mpc_status=None):
    """Host/torch restatement of what the device accumulates per step: -> (int64[16] counters, float64[4] sums).
    done_bits [E] int, rewards [E,3] = (R_PT, R_HA, R)."""
    import torch
    bits = done_bits.to(torch.int64)
    c = torch.zeros(16, dtype=torch.int64, device=bits.device)
    c[0] = bits.numel()
    for i, name in enumerate(STAT_NAMES[1:9], start=1):
        c[i] = ((bits & _BIT_OF[name]) != 0).sum()
    if mpc_status is not None:
        c[9] = mpc_status[:, 1].sum()
        c[10] = (mpc_status[:, 4] < 0).sum()
    s = torch.zeros(4, dtype=torch.float64, device=bits.device)
    r = rewards.to(torch.float64)
    s[0], s[1], s[2] = r[:, 2].sum(), r[:, 0].sum(), r[:, 1].sum()
    return c, s


def reduce_episode_stats(counters, sums, group=None):
    """In-place all-reduce (sum) of the statistics over all ranks; a no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return counters, sums


def stats_dict(counters, sums):
    c = counters.detach().cpu().tolist()
    s = sums.detach().cpu().tolist()
    out = {n: int(c[i]) for i, n in enumerate(STAT_NAMES)}
    out.update({"sum_R": s[0], "sum_R_PT": s[1], "sum_R_HA": s[2]})
    return out


def init_stats_comm(env, group=None):
    """Give `env` the engine's own NCCL communicator for dre_stats_reduce (include/dre.h): rank 0's id travels through
    torch.distributed (any backend), the all-reduce itself is issued by libdre over NCCL / NVLink."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("init_stats_comm needs an initialised torch.distributed process group")
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def bcast(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0, group=group)
        return box[0]
    env.stats_comm_init(rank, world, bcast)
