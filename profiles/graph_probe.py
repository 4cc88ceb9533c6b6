"""Is the step loop launch-bound?  K back-to-back launches timed three ways: plain launches, the same
launches captured in a CUDA graph, and the sum of the kernels' own lifetimes (globaltimer, -DPKB_TIMING builds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

NB, K = 3, 60
bs = [GameBatch(1 << 20, 2, 100, 2, 1, seed=0, env_id_base=j << 20, auto_reset=True) for j in range(NB)]
for b in bs:
    b.reset()
for t in range(30):
    bs[t % NB].step_random(t)
torch.cuda.synchronize()


def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def plain():
    for t in range(K):
        bs[t % NB].step_random(100 + t)


print(f"plain launches : {timed(plain) / K * 1e3:.1f} us per step")
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    plain()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    plain()
print(f"CUDA graph     : {timed(graph.replay) / K * 1e3:.1f} us per step")

# one stream per batch: consecutive steps belong to different batches and may overlap (tail of one, head of the next)
streams = [torch.cuda.Stream() for _ in range(NB)]


def overlapped():
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for t in range(K):
        with torch.cuda.stream(streams[t % NB]):
            bs[t % NB].step_random(300 + t)
    for s in streams:
        cur.wait_stream(s)


print(f"{NB} streams      : {timed(overlapped) / K * 1e3:.1f} us per step")
