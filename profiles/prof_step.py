"""Minimal driver for ncu: a few launches of the fused sample+step kernel on the bench workload
(heads-up, 2^20 envs, uniform random agents, auto-reset), plus the mask evaluator.

    python profiles/prof_step.py [--players P] [--envs N] [--steps T]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402
from pokerl_b200.judger import eval_masks_batch, random_hands  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--players", type=int, default=2)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--eval", action="store_true")
args = ap.parse_args()

nine = args.players == 9  # BASELINE.json configs[3]: unequal stacks, all-in heavy action mix
g = GameBatch(args.envs, args.players, [20 * (i + 1) for i in range(9)] if nine else 100, 2, 1, seed=1 if nine else 0, auto_reset=True)
w = [10, 30, 30, 5, 5, 5, 45] if nine else None
g.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(args.steps):
    flush.zero_()
    g.step_random(t, w)
if args.eval:
    masks = random_hands(0, 1 << 24)
    for _ in range(3):
        eval_masks_batch(masks)
torch.cuda.synchronize()
print("ok", g.stats())
