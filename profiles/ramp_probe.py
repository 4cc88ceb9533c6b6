"""Timed-region overhead of the pipelined step launches: device time of K back-to-back launches (3 handles round-robin,
one stream) for several K, each bracketed by a synchronize, to separate the launch interval from ramp-up and drain."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

gs = []
for j in range(3):
    g = GameBatch(1 << 20, 2, 100, 2, 1, seed=j, auto_reset=True)
    g.reset()
    gs.append(g)
t = 0
for _ in range(64 * 3):
    gs[t % 3].step_random(t)
    t += 1
torch.cuda.synchronize()
res = {}
for K in (3, 6, 12, 30, 60, 150, 300):
    best = 1e9
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(K):
            gs[t % 3].step_random(t)
            t += 1
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3)
    res[K] = best
    print(f"K={K:4d} launches: {best:8.1f} us total, {best / K:6.2f} us per launch")
k1, k2 = 60, 300
slope = (res[k2] - res[k1]) / (k2 - k1)
print(f"interval {slope:.2f} us; fixed overhead at K=60: {res[60] - 60 * slope:.1f} us")
