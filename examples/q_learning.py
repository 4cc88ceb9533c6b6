"""Batched version of the reference's examples/q_learning.py (BASELINE.json configs[4]).

Same learner -- tabular Q over (hand ranking of the agent's cards, move), softmax exploration mixed with
uniform exploration (q_learning.py:27-41, 69-93) -- but N tables per GPU advance in lockstep against
device-side RandomAgents, the hand-ranking feature comes from the showdown evaluator on 2/5/6/7 cards
(`pkb_hand_features`, q_learning.py:29-33), and the per-step Q updates of all tables are summed into one
[10, 7] delta that is all-reduced over the ranks (NCCL over NVLink) before it is applied.  The learner sees only
(hand ranking, valid mask) -- 10 x 128 situations -- so its softmax policy and its max-Q targets are worked out once per
step for every situation (`policy_tables`) and every table just looks its own up: `PokerGameEnv.step` and `reset` are one
launch each, the rest is a handful of [N]-sized torch ops.

    python examples/q_learning.py --envs 65536 --steps 200
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/q_learning.py --envs 8388608
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import distributed as D  # noqa: E402
from pokerl_b200.agents import RandomAgent  # noqa: E402
from pokerl_b200.enums import HandRanking, PokerMoves  # noqa: E402
from pokerl_b200.envs import PokerGameEnv  # noqa: E402


def features(env):
    """(hand ranking of seat 0's visible cards int64 [N], valid-move bit mask int64 [N])  -- q_learning.py:29-33"""
    ranking = (env.game.batch.hand_features(0).long() >> 20) & 0xF
    return ranking, env.game.batch.valid_actions(0).long()


_MASK_BITS = None


def mask_bits(dev):
    """bool [128, 7]: which moves each 7-bit valid mask allows (FOLD is always among them)."""
    global _MASK_BITS
    if _MASK_BITS is None or _MASK_BITS.device != dev:
        m = torch.arange(128, device=dev)
        _MASK_BITS = ((m.unsqueeze(1) >> torch.arange(7, device=dev)) & 1).bool()
    return _MASK_BITS


def policy_tables(Q, eps):
    """The learner sees only (hand ranking, valid mask): 10 x 128 distinct situations.  For every one of them the move
    distribution of q_learning.py:35-41, 81-82 -- softmax(Q[s, valid]) with probability 1 - eps, uniform over the valid
    moves otherwise -- as a CDF [10, 128, 7], and max over the valid moves of Q[s] [10, 128] (q_learning.py:88)."""
    valid = mask_bits(Q.device)                                    # [128, 7]
    logits = Q.unsqueeze(1).masked_fill(~valid.unsqueeze(0), float("-inf"))  # [10, 128, 7]
    any_valid = valid.any(dim=1).view(1, 128, 1)
    soft = torch.softmax(torch.where(any_valid, logits, torch.zeros_like(logits)), dim=2)
    unif = valid.double() / valid.sum(dim=1, keepdim=True).clamp(min=1)
    p = eps * unif.unsqueeze(0) + (1.0 - eps) * soft
    best = torch.where(any_valid.squeeze(2), logits.max(dim=2).values, torch.zeros((), dtype=Q.dtype, device=Q.device))
    return p.cumsum(dim=2), best


def predict(cdf, s, mask, gen):
    """One move per table, drawn from its situation's CDF (the per-table softmax + multinomial of the scalar example)."""
    row = cdf[s, mask]                                             # [N, 7]
    u = torch.rand(s.shape[0], device=s.device, dtype=row.dtype, generator=gen).unsqueeze(1) * row[:, -1:]
    return (u >= row).sum(dim=1).clamp(max=6)


def td_update(Q, best, s, u, r, hand, ns, nmask, gamma):
    """q_learning.py:88-89 for every table, summed into one delta and one visit count per (ranking, move) cell; returns the
    [2 * 70] vector to be all-reduced."""
    td_target = r + (~hand).double() * gamma * best[ns, nmask]
    cell = s * PokerMoves.NUM_MOVES + u
    n_cells = Q.numel()
    return torch.cat([torch.bincount(cell, weights=td_target - Q.view(-1)[cell], minlength=n_cells),
                      torch.bincount(cell, minlength=n_cells).double()])


def apply_update(Q, delta, alpha):
    d, n = delta[:Q.numel()].view_as(Q), delta[Q.numel():].view_as(Q)
    Q.add_(alpha * d / n.clamp(min=1.0))


def run(num_envs=4096, num_steps=100, num_opponents=3, alpha=0.01, gamma=1.0, seed=0, log_every=20, quiet=False):
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    base, n_local = D.shard_envs(num_envs * world, rank, world)
    # the learner only plays moves out of the valid mask, so the per-step host check for rejected moves is off
    env = PokerGameEnv(agents=[RandomAgent() for _ in range(num_opponents)], num_players=num_opponents + 1,
                       num_envs=n_local, env_id_base=base, seed=seed, device=local_rank, check_actions=False)
    gen = torch.Generator(device=dev).manual_seed(seed * 1000003 + rank)
    Q = torch.zeros((HandRanking.NONE, PokerMoves.NUM_MOVES), dtype=torch.float64, device=dev)  # q_learning.py:27
    env.reset()
    s, mask = features(env)
    score = torch.zeros(2, dtype=torch.float64, device=dev)  # sum of rewards, finished episodes
    for step in range(num_steps):
        eps = 1 - (step / num_steps) ** 2  # q_learning.py:67
        cdf, best = policy_tables(Q, eps)
        u = predict(cdf, s, mask, gen)
        _, r, done, hand = env.step(u.to(torch.int32))
        env.reset(env_mask=done)
        ns, nmask = features(env)
        delta = td_update(Q, best, s, u, r, hand, ns, nmask, gamma)
        score += torch.stack([r.sum(), done.double().sum()])
        D.all_reduce_sum_(delta)  # the path's only collective
        apply_update(Q, delta, alpha)
        s, mask = ns, nmask
        if not quiet and rank == 0 and (step + 1) % log_every == 0:
            tot = score  # this rank's tables
            print(f"step {step + 1}: mean reward per finished episode {tot[0].item() / max(tot[1].item(), 1):+.3f}; "
                  f"Q[HIGH]={Q[HandRanking.HIGH].cpu().numpy().round(3)}")
    D.all_reduce_sum_(score)
    stats = D.reduce_stats(env.game.batch.stats(), device=dev)
    return Q.cpu(), score.cpu(), stats


def bench_block(envs_per_gpu, rank, local_rank, world, timed, steps=6, num_opponents=3, seed=0):
    """bench.py's `q_learning` block (BASELINE.json configs[4]): `steps` learner steps over envs_per_gpu tables on this
    rank -- PokerGameEnv.step against device-side RandomAgents, the hand-ranking features, the TD update summed into
    one [10, 7] delta and its all-reduce over the ranks -- timed on the device (max over ranks is taken by the caller's
    `timed`).  Returns the JSON block (only meaningful on rank 0)."""
    dev = torch.device("cuda", local_rank)
    base, n_local = D.shard_envs(envs_per_gpu * world, rank, world)
    # the learner only ever plays moves out of the valid mask, so the per-step host check for rejected moves is off
    env = PokerGameEnv(agents=[RandomAgent() for _ in range(num_opponents)], num_players=num_opponents + 1,
                       num_envs=n_local, env_id_base=base, seed=seed, device=local_rank, check_actions=False)
    gen = torch.Generator(device=dev).manual_seed(seed * 1000003 + rank)
    Q = torch.zeros((HandRanking.NONE, PokerMoves.NUM_MOVES), dtype=torch.float64, device=dev)
    env.reset()
    state = {}
    state["s"], state["mask"] = features(env)
    alpha, gamma = 0.01, 1.0
    score = torch.zeros(2, dtype=torch.float64, device=dev)

    def learner_steps(k, eps=0.5):
        for _ in range(k):
            s, mask = state["s"], state["mask"]
            cdf, best = policy_tables(Q, eps)
            u = predict(cdf, s, mask, gen)
            _, r, done, hand = env.step(u.to(torch.int32))
            env.reset(env_mask=done)
            ns, nmask = features(env)
            delta = td_update(Q, best, s, u, r, hand, ns, nmask, gamma)
            score.add_(torch.stack([r.sum(), done.double().sum()]))
            D.all_reduce_sum_(delta)  # NCCL over NVLink: the Q-delta and visit counts of every rank
            apply_update(Q, delta, alpha)
            state["s"], state["mask"] = ns, nmask

    learner_steps(2)  # warm-up (allocator, NCCL channel set-up)
    before = env.game.batch.stats()["steps"]
    ms = timed(lambda: learner_steps(steps))
    game_steps = env.game.batch.stats()["steps"] - before
    tot = torch.tensor([float(game_steps)], dtype=torch.float64, device=dev)
    D.all_reduce_sum_(tot)
    D.all_reduce_sum_(score)
    P = num_opponents + 1
    per_step_bytes = 2 * (25 * P + 20) + 6
    return {
        "metric": "env_steps_per_sec", "value": envs_per_gpu * world * steps / (ms / 1e3), "unit": "PokerGameEnv steps/s (learner steps x tables)",
        "game_steps_per_sec": tot.item() / (ms / 1e3), "game_steps_per_env_step": tot.item() / (envs_per_gpu * world * steps),
        "learner_steps": steps, "ms_per_learner_step": ms / steps, "scaling": "strong",
        "config": {"workload": "examples/q_learning.py batched (BASELINE.json configs[4])", "num_players": P,
                   "global_envs": envs_per_gpu * world, "envs_per_gpu": envs_per_gpu, "opponents": "RandomAgent x3 on device",
                   "collective": "all_reduce(sum) of the [2, 10, 7] f64 Q-delta + visit counts per learner step" +
                                 (" (NCCL)" if world > 1 else " (single rank: identity)")},
        "algorithmic_bytes_per_game_step": per_step_bytes,
        "game_step_gbs_per_gpu": per_step_bytes * tot.item() / world / (ms / 1e3) / 1e9,
        "mean_reward_per_finished_episode": score[0].item() / max(score[1].item(), 1.0),
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="tables per GPU")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    Q, score, stats = run(args.envs, args.steps, seed=args.seed)
    if int(os.environ.get("RANK", "0")) == 0:
        print("final Q[FULL]     =", Q[HandRanking.FULL].numpy().round(3))
        print("final Q[HIGH]     =", Q[HandRanking.HIGH].numpy().round(3))
        print(f"episodes {int(score[1])}, mean reward {score[0].item() / max(score[1].item(), 1):+.3f}, stats {stats}")
