"""Time of the exhaustive evaluation of all C(52,k) hands on the device (BASELINE.json configs[1]: k = 7)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200.judger import eval_exhaustive  # noqa: E402

for k in (5, 6, 7):
    eval_exhaustive(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist, a, b = eval_exhaustive(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = int(sum(hist))
    print(f"k={k}: {n} hands in {dt * 1e3:.2f} ms = {n / dt / 1e9:.1f} G evals/s (enumerated on chip, histogram + two checksums)")
