"""Device time per step when step launches follow each other on ONE stream (what a caller that steps a GameBatch in a
loop gets): (a) one handle stepped back to back, (b) three handles round-robin (288 MB of state > L2), and the same
three handles on three streams for comparison.

    [PKB_LIB=variant.so] [PKB_NO_PDL=1] python profiles/chain_bench.py [--players P] [--envs N] [--steps K]
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
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--batches", type=int, default=3)
args = ap.parse_args()
start = 100 if args.players != 9 else [20 * (i + 1) for i in range(9)]
w = None if args.players != 9 else [10, 30, 30, 5, 5, 5, 45]
gs = []
for j in range(args.batches):
    g = GameBatch(args.envs, args.players, start, 2, 1, seed=j, auto_reset=True)
    g.reset()
    gs.append(g)
for t in range(64):
    gs[t % args.batches].step_random(t, w)
torch.cuda.synchronize()
B = 2 * (25 * args.players + 20) + 6


def timed(fn, k):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / k


t_idx = [64]


def one_handle(k):
    for _ in range(k):
        gs[0].step_random(t_idx[0], w)
        t_idx[0] += 1


def round_robin(k):
    for _ in range(k):
        gs[t_idx[0] % args.batches].step_random(t_idx[0], w)
        t_idx[0] += 1


main = torch.cuda.current_stream()
streams = [torch.cuda.Stream() for _ in range(args.batches)]


def multi_stream(k):
    for st in streams:
        st.wait_stream(main)
    for _ in range(k):
        j = t_idx[0] % args.batches
        with torch.cuda.stream(streams[j]):
            gs[j].step_random(t_idx[0], w)
        t_idx[0] += 1
    for st in streams:
        main.wait_stream(st)


tag = f"lib={os.path.basename(os.environ.get('PKB_LIB', 'default'))} pdl={'off' if os.environ.get('PKB_NO_PDL') else 'on'} P={args.players} N={args.envs}"
for name, fn in (("one handle, one stream", one_handle), ("3 handles round-robin, one stream", round_robin),
                 ("3 handles, 3 streams", multi_stream)):
    fn(30)
    us = min(timed(fn, args.steps) for _ in range(3))
    print(f"{tag} {name}: {us:.1f} us/step  {args.envs / us / 1e3:.2f} G env-steps/s  ({B * args.envs / us / 1e3 / 6543.4:.3f} of HBM peak)")
