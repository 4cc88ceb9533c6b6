"""Device time per step of the fused sample+step kernel the way bench.py measures its headline: three resident batches
(258 MB of state > L2) stepped round-robin, one stream each, so the tail of one launch overlaps the next batch's launch.

    [PKB_LIB=/path/to/variant.so] python profiles/overlap_bench.py [--players P] [--envs N] [--steps K]
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
ap.add_argument("--steps", type=int, default=600)
ap.add_argument("--batches", type=int, default=3)
args = ap.parse_args()
start = 100 if args.players != 9 else [20 * (i + 1) for i in range(9)]
w = None if args.players != 9 else [10, 30, 30, 5, 5, 5, 45]
main = torch.cuda.current_stream()
streams = [torch.cuda.Stream() for _ in range(args.batches)]
gs = []
for j in range(args.batches):
    g = GameBatch(args.envs, args.players, start, 2, 1, seed=j, auto_reset=True)
    g.reset()
    gs.append(g)
torch.cuda.synchronize()


def run(t0, k):
    for st in streams:
        st.wait_stream(main)
    for t in range(t0, t0 + k):
        j = t % args.batches
        with torch.cuda.stream(streams[j]):
            gs[j].step_random(t, w)
    for st in streams:
        main.wait_stream(st)


run(0, 64)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(main)
run(64, args.steps)
b.record(main)
torch.cuda.synchronize()
us = a.elapsed_time(b) * 1e3 / args.steps
B = 2 * (25 * args.players + 20) + 6
print(f"lib={os.path.basename(os.environ.get('PKB_LIB', 'default'))} P={args.players} N={args.envs} overlapped: {us:.1f} us/step  "
      f"{args.envs / us / 1e3:.2f} G env-steps/s  {B * args.envs / us / 1e3:.0f} GB/s ({B * args.envs / us / 1e3 / 6543.4:.3f})")
