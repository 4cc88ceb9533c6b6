"""Batched version of the reference's examples/q_learning.py (BASELINE.json configs[4]).

Same learner -- tabular Q over (hand ranking of the agent's cards, move), softmax exploration mixed with
uniform exploration (q_learning.py:27-41, 69-93) -- but N tables per GPU advance in lockstep against
device-side RandomAgents, the hand-ranking feature comes from the showdown evaluator on 2/5/6/7 cards
(`pkb_hand_features`, q_learning.py:29-33), and the per-step Q updates of all tables are summed into one
[10, 7] delta that is all-reduced over the ranks (NCCL over NVLink) before it is applied.

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
    """(hand ranking of seat 0's visible cards [N], valid-action matrix bool [N, 7])  -- q_learning.py:29-33"""
    ranking = (env.game.batch.hand_features(0).long() >> 20) & 0xF
    mask = env.game.batch.valid_actions(0)
    valid = ((mask.unsqueeze(1) >> torch.arange(7, device=mask.device)) & 1).bool()
    return ranking, valid


def predict(Q, s, valid, eps, gen):
    """softmax(Q[s, valid]) with probability 1 - eps, uniform over the valid moves otherwise (q_learning.py:35-41, 81-82)"""
    logits = Q[s].masked_fill(~valid, float("-inf"))
    p_soft = torch.softmax(logits, dim=1)
    p_unif = valid.double() / valid.sum(dim=1, keepdim=True)
    explore = torch.rand(s.shape[0], device=s.device, generator=gen) < eps
    p = torch.where(explore.unsqueeze(1), p_unif, p_soft)
    return torch.multinomial(p, 1, generator=gen)[:, 0]


def run(num_envs=4096, num_steps=100, num_opponents=3, alpha=0.01, gamma=1.0, seed=0, log_every=20, quiet=False):
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    base, n_local = D.shard_envs(num_envs * world, rank, world)
    env = PokerGameEnv(agents=[RandomAgent() for _ in range(num_opponents)], num_players=num_opponents + 1,
                       num_envs=n_local, env_id_base=base, seed=seed, device=local_rank)
    gen = torch.Generator(device=dev).manual_seed(seed * 1000003 + rank)
    Q = torch.zeros((HandRanking.NONE, PokerMoves.NUM_MOVES), dtype=torch.float64, device=dev)  # q_learning.py:27
    env.reset()
    s, valid = features(env)
    score = torch.zeros(2, dtype=torch.float64, device=dev)  # sum of rewards, finished episodes
    for step in range(num_steps):
        eps = 1 - (step / num_steps) ** 2  # q_learning.py:67
        u = predict(Q, s, valid, eps, gen)
        _, r, done, hand = env.step(u.to(torch.int32))
        if bool(done.any()):
            env.reset(env_mask=done)
        ns, nvalid = features(env)
        # q_learning.py:88-89 for every table, summed into one delta (and one visit count) per cell
        best_next = Q[ns].masked_fill(~nvalid, float("-inf")).max(dim=1).values
        td_target = r + (~hand).double() * gamma * best_next
        cell = s * PokerMoves.NUM_MOVES + u
        delta = torch.zeros(Q.numel() * 2, dtype=torch.float64, device=dev)
        delta.index_add_(0, cell, td_target - Q.view(-1)[cell])
        delta.index_add_(0, Q.numel() + cell, torch.ones_like(td_target))
        score += torch.stack([r.sum(), done.double().sum()])
        D.all_reduce_sum_(delta)  # the path's only collective
        d, n = delta[:Q.numel()].view_as(Q), delta[Q.numel():].view_as(Q)
        Q += alpha * d / n.clamp(min=1.0)
        s, valid = ns, nvalid
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
    state = {"s": None, "valid": None, "env_steps": 0}
    state["s"], state["valid"] = features(env)
    alpha, gamma = 0.01, 1.0
    score = torch.zeros(2, dtype=torch.float64, device=dev)

    def learner_steps(k, eps=0.5):
        for _ in range(k):
            s, valid = state["s"], state["valid"]
            u = predict(Q, s, valid, eps, gen)
            _, r, done, hand = env.step(u.to(torch.int32))
            env.reset(env_mask=done)
            ns, nvalid = features(env)
            best_next = Q[ns].masked_fill(~nvalid, float("-inf")).max(dim=1).values
            td_target = r + (~hand).double() * gamma * best_next
            cell = s * PokerMoves.NUM_MOVES + u
            delta = torch.zeros(Q.numel() * 2, dtype=torch.float64, device=dev)
            delta.index_add_(0, cell, td_target - Q.view(-1)[cell])
            delta.index_add_(0, Q.numel() + cell, torch.ones_like(td_target))
            score.add_(torch.stack([r.sum(), done.double().sum()]))
            D.all_reduce_sum_(delta)  # NCCL over NVLink: the Q-delta and visit counts of every rank
            d, n = delta[:Q.numel()].view_as(Q), delta[Q.numel():].view_as(Q)
            Q.add_(alpha * d / n.clamp(min=1.0))
            state["s"], state["valid"] = ns, nvalid

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
