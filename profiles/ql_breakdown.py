"""Where a learner step of the batched q-learning example spends its time (2^23 four-player tables on one GPU)."""
import importlib.util
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

spec = importlib.util.spec_from_file_location("ql", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "q_learning.py"))
ql = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ql)
from pokerl_b200.agents import RandomAgent  # noqa: E402
from pokerl_b200.envs import PokerGameEnv  # noqa: E402

N = int(os.environ.get("PKB_QL_ENVS", str(1 << 23)))
dev = torch.device("cuda", 0)
env = PokerGameEnv(agents=[RandomAgent() for _ in range(3)], num_players=4, num_envs=N, seed=0, check_actions=False)
gen = torch.Generator(device=dev).manual_seed(1)
Q = torch.zeros((10, 7), dtype=torch.float64, device=dev)
env.reset()
s, mask = ql.features(env)
times = {}


def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    times[name] = times.get(name, 0.0) + a.elapsed_time(b)
    return out


K = 6
for it in range(K + 2):
    if it == 2:
        times.clear()
    cdf, best = timed("policy_tables", lambda: ql.policy_tables(Q, 0.5))
    u = timed("predict", lambda: ql.predict(cdf, s, mask, gen))
    a32 = timed("to int32", lambda: u.to(torch.int32))
    _, r, done, hand = timed("env.step (one launch + StateView)", lambda: env.step(a32))
    timed("env.reset(mask) (one launch + StateView)", lambda: env.reset(env_mask=done))
    ns = timed("hand_features", lambda: (env.game.batch.hand_features(0).long() >> 20) & 0xF)
    nmask = timed("valid_actions", lambda: env.game.batch.valid_actions(0).long())
    delta = timed("td_update", lambda: ql.td_update(Q, best, s, u, r, hand, ns, nmask, 1.0))
    timed("apply", lambda: ql.apply_update(Q, delta, 0.01))
    s, mask = ns, nmask
tot = sum(times.values())
for k, v in sorted(times.items(), key=lambda kv: -kv[1]):
    print(f"{k:45s} {v / K:8.2f} ms per learner step ({100 * v / tot:4.1f} %)")
print(f"{'total':45s} {tot / K:8.2f} ms; {N} tables")
