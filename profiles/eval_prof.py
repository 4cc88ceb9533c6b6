"""One warm launch then a few more of the batch evaluator on 2^24 random hands, for ncu (-k regex:eval_masks -s 2 -c 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200.judger import eval_masks_batch, random_hands  # noqa: E402

masks = random_hands(0, 1 << 24)
for _ in range(4):
    out = eval_masks_batch(masks)
torch.cuda.synchronize()
print("ok", int(out[:4].sum()))
