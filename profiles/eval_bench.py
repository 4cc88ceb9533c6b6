"""7-card evals/s of the batch evaluator on 2^27 random hands resident in HBM (12 B per eval: 8 in, 4 out).

    [PKB_LIB=/path/to/variant.so] [PKB_EVAL_IMPL=table] python profiles/eval_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import _native as nat  # noqa: E402
from pokerl_b200.judger import random_hands  # noqa: E402

n = 1 << 27
masks = random_hands(0, n)
out = torch.empty(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def eval_masks_batch(masks, out):
    nat.check(nat.lib().pkb_eval_masks(nat.ptr(masks), n, nat.ptr(out), stream))


for _ in range(3):
    eval_masks_batch(masks, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K = 10
for _ in range(K):
    eval_masks_batch(masks, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
print(f"lib={os.path.basename(os.environ.get('PKB_LIB', 'default'))} impl={os.environ.get('PKB_EVAL_IMPL', 'hash')}: {ms * 1e3:.0f} us per 2^27 hands, "
      f"{n / ms / 1e6:.1f} G evals/s, {12 * n / ms / 1e6:.0f} GB/s ({12 * n / ms / 1e6 / 6543.4:.3f})")
