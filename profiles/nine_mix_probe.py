"""Nine-player bench workload (BASELINE configs[3]): device time per step in windows of 20 steps from the reset on, to see
where the hand mix settles (the bench pre-rolls into that region)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

N = int(os.environ.get("PKB_NINE_ENVS", str(1 << 22)))
g = GameBatch(N, 9, [20 * (i + 1) for i in range(9)], 2, 1, seed=1, auto_reset=True)
w = [10, 30, 30, 5, 5, 5, 45]
g.reset()
t = 0
for win in range(12):
    g.stats_reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        g.step_random(t, w)
        t += 1
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    s = g.stats()
    print(f"steps {t - 20:3d}-{t - 1:3d}: {us:7.1f} us/step  {N / us / 1e3:.2f} G env-steps/s  ({496 * N / us / 1e3 / 6543.4:.3f})  "
          f"hands/step {s['hands'] / s['steps']:.3f} showdowns/step {s['showdowns'] / s['steps']:.3f} games/step {s['games'] / s['steps']:.4f}")
