"""Aggregate pinned host <-> device copy rate of the box when 1, 2, 4, 8 GPUs copy at the same time, in the bench's
end-to-end pattern (per step and batch 0.5 MB host->device and 1 MB device->host, eight independent streams per GPU).
Run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/pcie_probe_multi.py

It is what bounds `e2e` at N = 8: no kernel of this repo runs here."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
STREAMS, ROUNDS = 8, 64
H2D, D2H = 512 << 10, 1 << 20
streams = [torch.cuda.Stream() for _ in range(STREAMS)]
h_in = [torch.empty((ROUNDS, H2D), dtype=torch.uint8).pin_memory() for _ in range(STREAMS)]
h_out = [torch.empty((ROUNDS, D2H), dtype=torch.uint8).pin_memory() for _ in range(STREAMS)]
d_in = [torch.empty(H2D, dtype=torch.uint8, device=dev) for _ in range(STREAMS)]
d_out = [torch.empty(D2H, dtype=torch.uint8, device=dev) for _ in range(STREAMS)]


def burst():
    for r in range(ROUNDS):
        for j, st in enumerate(streams):
            with torch.cuda.stream(st):
                d_in[j].copy_(h_in[j][r], non_blocking=True)
                h_out[j][r].copy_(d_out[j], non_blocking=True)
    for st in streams:
        st.synchronize()


burst()
flag = torch.zeros(1, device=dev)
active = 1
while active <= world:
    dist.all_reduce(flag)  # device-side barrier
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank < active:
        burst()
    dt = time.perf_counter() - t0
    mine = torch.tensor([(H2D + D2H) * STREAMS * ROUNDS / dt / 1e9 if rank < active else 0.0], device=dev, dtype=torch.float64)
    lo = mine.clone() if rank < active else torch.full_like(mine, 1e30)
    dist.all_reduce(mine)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{active} GPU(s) copying: aggregate {mine.item():7.1f} GB/s (both directions), slowest GPU {lo.item():6.1f} GB/s, "
              f"= {mine.item() / 1.5:6.1f} G env-steps/s at 1.5 B per env-step", flush=True)
    active *= 2
dist.destroy_process_group()
