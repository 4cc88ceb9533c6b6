"""World-size-2 gloo tests of the multi-GPU host logic: env sharding by contiguous
global-id slices and the one collective of the path, the statistics / Q-delta sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pokerl_b200 import distributed as D
    r, lr, w = D.init_from_env(backend="gloo")
    base, n = D.shard_envs(n_global, r, w)
    stats = dict(steps=n * 10, hands=n * 3, showdowns=n, games=r + 1, invalid=0, poisoned=0,
                 payoff_sum=[float(base), -float(base), 0.5 * (r + 1)])
    red = D.reduce_stats(stats)
    q = torch.full((10, 7), float(r + 1), dtype=torch.float64)
    D.all_reduce_sum_(q)
    out[rank] = (base, n, red, q.sum().item())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [1 << 20, 1000003])
def test_two_rank_sharding_and_stat_reduction(n_global):
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_global, out), nprocs=world, join=True)
    (b0, n0, red0, q0), (b1, n1, red1, q1) = out[0], out[1]
    assert b0 == 0 and b1 == n0 and n0 + n1 == n_global and abs(n0 - n1) <= 1  # contiguous, complete, balanced
    assert red0 == red1
    assert red0["steps"] == 10 * n_global and red0["hands"] == 3 * n_global and red0["games"] == 3
    assert red0["payoff_sum"] == [float(b0 + b1), -float(b0 + b1), 1.5]
    assert q0 == q1 == 70 * 3.0


def test_shard_envs_partitions_any_world():
    from pokerl_b200.distributed import shard_envs
    for n in (1, 7, 4096, (1 << 22) + 5):
        for w in (1, 2, 4, 8):
            spans = [shard_envs(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(w - 1))
    with pytest.raises(ValueError):
        shard_envs(10, 2, 2)


def test_single_process_is_identity():
    from pokerl_b200 import distributed as D
    s = dict(steps=5, hands=2, showdowns=1, games=0, invalid=0, poisoned=0, payoff_sum=[1.5, -1.5])
    assert D.reduce_stats(s) == s
