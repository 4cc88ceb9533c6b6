"""Device time per Game.step when a launch steps every tile K times in shared memory (pkb_step_random_n), three 2^20-env
handles round-robin on one stream (the bench's streaming configuration).

    python profiles/repeat_probe.py [players ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

for P in [int(x) for x in sys.argv[1:]] or [2, 4, 9]:
    start = 100 if P != 9 else [20 * (i + 1) for i in range(9)]
    w = None if P != 9 else [10, 30, 30, 5, 5, 5, 45]
    gs = []
    for j in range(3):
        g = GameBatch(1 << 20, P, start, 2, 1, seed=j, auto_reset=True)
        g.reset()
        gs.append(g)
    t = 0
    for _ in range(64):
        for g in gs:
            g.step_random(t, w)
        t += 1
    B = 2 * (25 * P + 20) + 6
    for K in (1, 2, 4, 8, 16):
        launches = max(6, 240 // K)
        best = 1e30
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for i in range(launches):
                gs[i % 3].step_random(t + (i // 3) * K, w, repeat=K)
            b.record()
            torch.cuda.synchronize()
            t += (launches // 3 + 1) * K
            best = min(best, a.elapsed_time(b) * 1e3 / (launches * K))
        print(f"P={P} K={K:2d}: {best:7.2f} us per step of 2^20 envs  {(1 << 20) / best / 1e3:6.2f} G env-steps/s  ({B * (1 << 20) / best / 1e3 / 6543.4:.3f} of HBM peak in algorithmic bytes)", flush=True)
    del gs
