"""Steady-state DRAM traffic of the pipelined step launches, for ncu's range replay:

    ncu --replay-mode range --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
        --csv --log-file gpurun_out/range_r02.csv python profiles/range_probe.py

The profiled range is LAUNCHES back-to-back step launches (three 2^20-env heads-up batches round-robin on one stream, the
bench's headline configuration) after the steady-state pre-roll: the launches overlap as they do in the bench, and no
cache flush separates them, so dram__bytes / LAUNCHES is what one step really moves."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

LAUNCHES = int(os.environ.get("PKB_RANGE_LAUNCHES", "12"))
P = int(os.environ.get("PKB_RANGE_PLAYERS", "2"))
nine = P == 9
gs = []
for j in range(3):
    g = GameBatch(1 << 20, P, [20 * (i + 1) for i in range(9)] if nine else 100, 2, 1, seed=j, auto_reset=True)
    g.reset()
    gs.append(g)
w = [10, 30, 30, 5, 5, 5, 45] if nine else None
t = 0
for _ in range(64 * 3):
    gs[t % 3].step_random(t, w)
    t += 1
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(LAUNCHES):
    gs[t % 3].step_random(t, w)
    t += 1
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", LAUNCHES, "launches in the profiled range")
