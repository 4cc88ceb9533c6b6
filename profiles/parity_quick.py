"""Quick bit-exact check of the step kernel against the oracle for whichever library PKB_LIB selects
(used when sweeping kernel variants): heads-up and 9-player batches, fused random-agent steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import pyoracle as ora  # noqa: E402
from pokerl_b200 import GameBatch  # noqa: E402

ok = True
for P, N, T in ((2, 20000, 64), (9, 5000, 64)):
    g = GameBatch(N, P, 100, 2, 1, seed=1, auto_reset=True, device=0)
    o = ora.OracleBatch(N, P, 100, 2, 1, seed=1, auto_reset=True)
    g.reset(); o.reset()
    for t in range(T):
        g.step_random(t)
    o.run_random(0, T)
    ge, oe = g.export(), o.export()
    for k in ("credits", "bets", "pending", "payoffs", "minraise"):
        ok &= np.array_equal(np.ascontiguousarray(ge[k]).view(np.uint64), np.ascontiguousarray(oe[k]).view(np.uint64))
    for k in ("states", "turn", "hand", "dealer", "sb", "bb", "active", "episode"):
        ok &= np.array_equal(ge[k], oe[k])
    ok &= g.stats()["steps"] == N * T
print("parity", "OK" if ok else "MISMATCH", os.environ.get("PKB_LIB", "default"))
sys.exit(0 if ok else 1)
