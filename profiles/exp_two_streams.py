"""Experiment: do the three step kernels of independent env slices overlap when they free-run on
separate streams?  Compares one 2^20-env handle with H handles of 2^20/H envs on H streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pokerl_b200 import GameBatch

N, K = 1 << 20, 300
for H in (1, 2, 4, 8):
    gs = [GameBatch(N // H, 2, 100, 2, 1, seed=0, env_id_base=i * (N // H), auto_reset=True) for i in range(H)]
    ss = [torch.cuda.Stream() for _ in range(H)]
    for g in gs:
        g.reset()
    torch.cuda.synchronize()
    for t in range(20):
        for g, s in zip(gs, ss):
            with torch.cuda.stream(s):
                g.step_random(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(K):
        for g, s in zip(gs, ss):
            with torch.cuda.stream(s):
                g.step_random(20 + t)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H={H}: {dt/K*1e6:.1f} us per 2^20-env step (host issue {t_issue/K*1e6:.1f} us)  {N*K/dt/1e9:.2f} G env-steps/s")
