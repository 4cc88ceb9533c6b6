"""Would slicing ONE 2^20-env batch into S independent sub-batches on S streams hide the per-launch tail?
Emulated with S GameBatch handles (the envs are independent, so this is the same computation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

N, K = 1 << 20, 60
for S in (1, 2, 3, 4, 6, 8):
    per = (N + S - 1) // S
    bs = [GameBatch(per, 2, 100, 2, 1, seed=0, env_id_base=j * per, auto_reset=True) for j in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    for b in bs:
        b.reset()
    torch.cuda.synchronize()

    def run(t0):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for t in range(K):
            for j in range(S):
                with torch.cuda.stream(streams[j]):
                    bs[j].step_random(t0 + t)
        for s in streams:
            cur.wait_stream(s)

    run(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(4):
        run(1000 + r * K)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (4 * K) * 1e3
    print(f"S={S}: {us:.1f} us per step of {per * S} envs = {per * S / us / 1e3:.2f} G env-steps/s")
    del bs
