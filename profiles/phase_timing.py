"""Phase cycle counters of the step kernel (needs a -DPKB_TIMING build selected with PKB_LIB)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402
from pokerl_b200 import _native as nat  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = GameBatch(1 << 20, P, 100, 2, 1, seed=0, auto_reset=True)
g.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(20):
    g.step_random(t)
torch.cuda.synchronize()
L = C.CDLL(os.environ["PKB_LIB"])
buf = (C.c_ulonglong * 8)()
L.pkb_debug_timing(buf)
span = (C.c_ulonglong * 3)()
L.pkb_debug_span(span)
g.step_random(99)
torch.cuda.synchronize()
L.pkb_debug_span(span)
blk = (C.c_ulonglong * (2048 * 4))()
L.pkb_debug_blocks(blk)
import numpy as np
B = np.array(list(blk), dtype=np.int64).reshape(2048, 4)
B = B[B[:, 1] > 0]
life = (B[:, 1] - B[:, 0]) / 1e3
print("blocks", len(B), "lifetime us: min %.1f p50 %.1f p90 %.1f max %.1f" % (life.min(), np.median(life), np.percentile(life, 90), life.max()))
for nt in sorted(set(B[:, 3])):
    sel = B[:, 3] == nt
    print(f"  blocks with {nt} tiles: {sel.sum()}, lifetime mean {life[sel].mean():.1f} max {life[sel].max():.1f}")
order = np.argsort(-life)[:8]
print("  slowest blocks (id, sm, tiles, us):", [(int(i), int(B[i, 2]), int(B[i, 3]), round(float(life[i]), 1)) for i in order])
sm_end = {}
for i in range(len(B)):
    sm_end.setdefault(int(B[i, 2]), []).append(round(float((B[i, 1] - B[:, 0].min()) / 1e3), 1))
ends = sorted(sm_end.items(), key=lambda kv: -max(kv[1]))
print("  SMs finishing last:", ends[:4], " first:", ends[-3:])
print(f"one launch: block starts spread over {(span[1] - span[0]) / 1e3:.1f} us, first start to last end {(span[2] - span[0]) / 1e3:.1f} us")
L.pkb_debug_timing(buf)
K = 20
for t in range(K):
    flush.zero_()
    g.step_random(100 + t)
torch.cuda.synchronize()
L.pkb_debug_timing(buf)
v = list(buf)
tiles = v[6]
blocks = K * min(148 * int(os.environ.get("PKB_BLOCKS_PER_SM", "3")), tiles // K)
print(f"globaltimer: mean block lifetime {v[7] / blocks / 1e3:.1f} us -> SM clock {v[5] / v[7] * 1e3:.0f} MHz")
names = ["load wait", "pass A", "pass C", "pass D", "store issue"]
print(f"blocks {blocks // K} per launch, tiles {tiles // K} per launch, block lifetime {v[5] / blocks:.0f} cycles = {v[5] / blocks / 1.965e3:.1f} us")
for n, c in zip(names, v[:5]):
    print(f"  {n:12s} {c / tiles:9.0f} cycles per tile ({100 * c / sum(v[:5]):4.1f}%)")
