"""Step 1 of the at-size parity chain (runs on the GPU box): play BASELINE.json configs[2] (heads-up, 2^20 envs) and
configs[3] (nine players, 2^22 envs, unequal stacks, all-in heavy mix) with the fused random-agent kernel at FULL size
and record, for a sample of envs spread over the whole batch, every step's action, flag byte and valid mask plus the
final state.  Step 2 (tests/golden/replay_at_size.py, build container) replays the recordings into the LIVE reference.

    python tests/golden/dump_gpu_at_size.py [outdir]      # writes at_size_p2.npz, at_size_p9.npz (default gpurun_out/)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pokerl_b200 import GameBatch  # noqa: E402

CONFIGS = {
    # name: P, envs, start_credits, seed, weights, steps, sampled envs
    "p2": (2, 1 << 20, 100, 0, None, 64, 4096),
    "p9": (9, 1 << 22, [20 * (i + 1) for i in range(9)], 1, [10, 30, 30, 5, 5, 5, 45], 48, 1024),
}


def sample_ids(n_envs, count):
    """`count` env ids spread over the batch: evenly spaced, plus the first and last tiles' edges."""
    ids = set(np.linspace(0, n_envs - 1, count - 8, dtype=np.int64).tolist())
    ids |= {0, 1, 127, 128, n_envs - 129, n_envs - 128, n_envs - 2, n_envs - 1}
    extra = 0
    while len(ids) < count:  # collisions of the two sets: fill up
        ids.add((extra * 7919 + 13) % n_envs)
        extra += 1
    return np.array(sorted(ids)[:count], dtype=np.int64)


def dump(name, outdir):
    P, N, start, seed, weights, T, S = CONFIGS[name]
    g = GameBatch(N, P, start, 2, 1, seed=seed, auto_reset=True)
    g.reset()
    ids = sample_ids(N, S)
    idx = torch.as_tensor(ids, device=g.device)
    actions = np.empty((T, S), dtype=np.uint8)
    flags = np.empty((T, S), dtype=np.uint8)
    valid = np.empty((T, S), dtype=np.uint8)
    valid0 = g.valid_actions()[idx].cpu().numpy()
    for t in range(T):
        f, v = g.step_random(t, weights, record_actions=True)
        actions[t] = g.actions_taken[idx].to(torch.uint8).cpu().numpy()
        flags[t] = f[idx].cpu().numpy()
        valid[t] = v[idx].cpu().numpy()
    e = g.export()
    final = {k: np.ascontiguousarray(v[..., ids]) for k, v in e.items()}
    for k in ("credits", "bets", "pending", "payoffs", "minraise"):
        final[k] = final[k].view(np.uint64)  # bit patterns
    st = g.stats()
    path = os.path.join(outdir, f"at_size_{name}.npz")
    np.savez_compressed(path, env_ids=ids, actions=actions, flags=flags, valid=valid, valid0=valid0,
                        config=np.array([P, N, seed, T], dtype=np.int64),
                        start_credits=np.array(start if isinstance(start, list) else [start] * P, dtype=np.float64),
                        weights=np.array(weights if weights else [1] * 7, dtype=np.uint32),
                        **{"final_" + k: v for k, v in final.items()})
    print(f"{name}: P={P} N={N} steps={T} sampled={S} stats={st} -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    for name in CONFIGS:
        dump(name, out)
