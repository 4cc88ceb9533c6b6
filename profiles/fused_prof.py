import os, sys
sys.path.insert(0, '/root/repo')
import torch
from pokerl_b200 import GameBatch
g = GameBatch(1 << 20, 2, 100, 2, 1, seed=0, auto_reset=True); g.reset()
for t in range(64): g.step_random(t)
torch.cuda.synchronize()
for i in range(4): g.step_random(100 + 8 * i, repeat=8)
torch.cuda.synchronize()
