"""Per-warp timeline of one launch of the warp-tile step kernel (needs a -DPKB_TIMING build selected with PKB_LIB).

Prints when warps and SMs finish, how many tiles each warp took and where a tile's cycles go."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
start = 100 if P != 9 else [20 * (i + 1) for i in range(9)]   # nine players: the stacks and action mix of BASELINE configs[3]
weights = None if P != 9 else [10, 30, 30, 5, 5, 5, 45]
g = GameBatch(1 << 20, P, start, 2, 1, seed=0, auto_reset=True)
g.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(20):
    g.step_random(t, weights)
torch.cuda.synchronize()
L = C.CDLL(os.environ["PKB_LIB"])
CHAIN = int(os.environ.get("PKB_TIMING_CHAIN", "1"))  # > 1: that many back-to-back launches (pipelined); the table shows the last
for rep in range(3):
    flush.zero_()
    for k in range(CHAIN):
        g.step_random(99 + rep * CHAIN + k, weights)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (4096 * 16))()
    L.pkb_debug_warps(buf)
    W = np.array(list(buf), dtype=np.int64).reshape(4096, 16)
    W = W[W[:, 1] > 0]
    t0 = W[:, 0].min()
    end = (W[:, 1] - t0) / 1e3
    start = (W[:, 0] - t0) / 1e3
    print(f"launch {rep}: warps {len(W)}, starts spread {start.max():.1f} us; ends: min {end.min():.1f} p10 {np.percentile(end, 10):.1f} p50 {np.median(end):.1f} "
          f"p90 {np.percentile(end, 90):.1f} p99 {np.percentile(end, 99):.1f} max {end.max():.1f} us")
    for nt in sorted(set(W[:, 8])):
        sel = W[:, 8] == nt
        print(f"   warps with {nt} tiles: {sel.sum():5d}  end mean {end[sel].mean():.1f} max {end[sel].max():.1f}")
    sm_end, sm_tiles = {}, {}
    for r, e in zip(W, end):
        sm_end[int(r[2])] = max(sm_end.get(int(r[2]), 0), e)
        sm_tiles[int(r[2])] = sm_tiles.get(int(r[2]), 0) + int(r[8])
    se = np.array(sorted(sm_end.values()))
    print(f"   SM finish: min {se.min():.1f} p50 {np.median(se):.1f} p90 {np.percentile(se, 90):.1f} max {se.max():.1f} us; tiles per SM min {min(sm_tiles.values())} max {max(sm_tiles.values())}")
    late = sorted(sm_end.items(), key=lambda kv: -kv[1])[:5]
    print("   last SMs (sm, end us, tiles):", [(k, round(v, 1), sm_tiles[k]) for k, v in late])
    tiles = W[:, 8].sum()
    names = ["load wait", "pass A", "pass C", "pass D", "store+next"]
    tot = W[:, 3:8].sum()
    for i, n in enumerate(names):
        print(f"   {n:10s} {W[:, 3 + i].sum() / tiles:8.0f} cycles per tile ({100 * W[:, 3 + i].sum() / tot:4.1f}%)")
    cnt = np.minimum(W[:, 8], 1).sum(), np.minimum(np.maximum(W[:, 8] - 1, 0), 1).sum(), np.maximum(W[:, 8] - 2, 0).sum()
    print("   load wait by round (cycles per tile): first %.0f second %.0f later %.0f" % tuple(W[:, 12 + i].sum() / max(1, cnt[i]) for i in range(3)))
    order = np.argsort(-end)[:6]
    print("   slowest warps (end us, sm, tiles, cycles A/C/D, C/D rounds, longest C pass, max showdowns):",
          [(round(float(end[i]), 1), int(W[i, 2]), int(W[i, 8]), int(W[i, 4]), int(W[i, 5]), int(W[i, 6]), int(W[i, 9]), int(W[i, 10]), int(W[i, 11])) for i in order])
    print(f"   C/D rounds per tile {W[:, 9].sum() / tiles:.3f}; longest C pass: p50 {np.median(W[:, 10]):.0f} p99 {np.percentile(W[:, 10], 99):.0f} max {W[:, 10].max()} cycles; max showdowns in a tile {W[:, 11].max()}")
