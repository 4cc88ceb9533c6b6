"""Multi-GPU plumbing: one process per GPU, environments sharded by contiguous global-id slices.

Environments are independent (the reference `Game` shares nothing between tables), so the data path has
no collective at all: rank r of W owns the global env ids [r*N/W, (r+1)*N/W) and steps them alone.  Decks
and random-agent draws are keyed by the GLOBAL env id, so results do not depend on W.  The only exchange
is the end-of-run (or per-interval) sum of a small statistics vector -- payoff sums, counters, and for the
q-learning example the [10, 7] Q-delta -- through `torch.distributed.all_reduce` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).
"""
import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("steps", "hands", "showdowns", "games", "invalid", "poisoned")


def shard_envs(global_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_id_base, num_local_envs) of `rank`: contiguous slices, the remainder spread over the first ranks."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    q, r = divmod(int(global_envs), world)
    return rank * q + min(rank, r), q + (1 if rank < r else 0)


def init_from_env(backend: str = None, device_index: int = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun);
    returns (rank, local_rank, world).  Single-process runs return (0, 0, 1) without initialising anything."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            idx = local_rank if device_index is None else device_index
            torch.cuda.set_device(idx)
            kw["device_id"] = torch.device("cuda", idx)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def reduce_stats(stats: Dict, device=None, group=None) -> Dict:
    """Sum a GameBatch.stats() dict over all ranks (identity without a process group)."""
    P = len(stats["payoff_sum"])
    if device is None and dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())  # NCCL reduces device tensors only
    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS] + [float(x) for x in stats["payoff_sum"]],
                       dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {k: int(vec[i].item()) for i, k in enumerate(STAT_KEYS)}
    out["payoff_sum"] = vec[len(STAT_KEYS):len(STAT_KEYS) + P].tolist()
    return out


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of a small tensor (e.g. the q-learning Q-delta); identity for one rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
