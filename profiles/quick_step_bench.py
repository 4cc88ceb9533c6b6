"""Quick device-time probe of the fused sample+step kernel (L2 flushed between steps).

    [PKB_LIB=/path/to/variant.so] python profiles/quick_step_bench.py [--players P] [--envs N] [--steps K]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--players", type=int, default=2)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--weights", type=str, default="")
args = ap.parse_args()
w = [int(x) for x in args.weights.split(",")] if args.weights else None
start = 100 if args.players != 9 else [20 * (i + 1) for i in range(9)]
g = GameBatch(args.envs, args.players, start, 2, 1, seed=0, auto_reset=True)
g.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(32):
    g.step_random(t, w)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
for i, (a, b) in enumerate(ev):
    flush.zero_()
    a.record()
    g.step_random(32 + i, w)
    b.record()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(args.steps):
    g.step_random(1000 + i, w)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_wall = time.perf_counter() - t0
ms = sorted(a.elapsed_time(b) for a, b in ev)
mean = sum(ms) / len(ms)
B = 2 * (25 * args.players + 20) + 6
print(f"slices={os.environ.get('PKB_SLICES', '1')} back-to-back (no flush): host issue {t_issue/args.steps*1e6:.1f}us/step, wall {t_wall/args.steps*1e6:.1f}us/step")
print(f"lib={os.environ.get('PKB_LIB', 'default')} P={args.players} N={args.envs} mean={mean*1e3:.1f}us median={ms[len(ms)//2]*1e3:.1f}us "
      f"min={ms[0]*1e3:.1f}us  {args.envs/mean/1e6:.2f} G env-steps/s  {B*args.envs/mean/1e6:.0f} GB/s algorithmic "
      f"({B*args.envs/mean/1e6/6543.4:.3f} of 6543.4)")
