"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): step kernel P=2 and P=9 (fused random agent
and external actions with skips), masked reset, observation kernels, the batch evaluators."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402
from pokerl_b200 import _native as nat  # noqa: E402
from pokerl_b200.judger import eval_exhaustive, eval_masks_batch, random_hands  # noqa: E402

for P, N in ((2, 3000), (9, 700)):
    g = GameBatch(N, P, 100 if P == 2 else [20 * (i + 1) for i in range(9)], 2, 1, seed=3, auto_reset=True)
    g.reset()
    for t in range(12):
        g.step_random(t, None if t % 2 else [1, 3, 3, 1, 1, 1, 4])
    for t in range(6):
        a = g.sample_actions(100 + t).clone()
        a[::5] = nat.ACTION_SKIP
        g.step(a)
    m = torch.zeros(N, dtype=torch.uint8, device="cuda")
    m[::7] = 1
    g.reset(m)
    g.cards(); g.hand_features(); g.hand_features(0); g.valid_actions(); g.game_over()
    print(P, g.stats())
masks = random_hands(1, 100001)
out = eval_masks_batch(masks)
print(int(out.sum()), eval_exhaustive(5)[0][1:4])
torch.cuda.synchronize()
print("sanitize_small done")
