"""GPU (-m gpu), round 2: the at-size parity chain, Game.reset(dealer=k), the batched StateView, pipelined launches."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F64_KEYS = ("credits", "bets", "pending", "payoffs", "minraise")
INT_KEYS = ("states", "turn", "hand", "dealer", "sb", "bb", "active", "episode")
ALLIN_HEAVY = [10, 30, 30, 5, 5, 5, 45]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def states_equal(g, o, ctx):
    ge, oe = g.export(), o.export()
    for k in F64_KEYS:
        a, b = np.ascontiguousarray(ge[k]).view(np.uint64), np.ascontiguousarray(oe[k]).view(np.uint64)
        assert np.array_equal(a, b), (ctx, k, np.argwhere(a != b)[:3].tolist())
    for k in INT_KEYS:
        assert np.array_equal(ge[k], oe[k]), (ctx, k, np.argwhere(np.asarray(ge[k]) != np.asarray(oe[k]))[:3].tolist())


@pytest.mark.parametrize("name", ["p2", "p9"])
def test_full_size_run_equals_the_recording_the_live_reference_confirmed(torch_cuda, golden_dir, name):
    """BASELINE.json configs[2] / configs[3] at FULL size.  tests/golden/at_size/*.npz are recordings of exactly these
    runs (sampled envs: action, flags, valid mask of every step, final state) that tests/golden/replay_at_size.py
    replayed into the unmodified reference bit for bit (tests/golden/at_size/replay.log); the kernel must reproduce
    them."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    z = np.load(os.path.join(golden_dir, "at_size", f"at_size_{name}.npz"))
    P, N, seed, T = [int(x) for x in z["config"]]
    start = z["start_credits"].tolist()
    weights = None if name == "p2" else [int(x) for x in z["weights"]]
    g = GameBatch(N, P, int(start[0]) if name == "p2" else start, 2, 1, seed=seed, auto_reset=True)
    g.reset()
    ids = z["env_ids"]
    idx = torch.as_tensor(ids, device=g.device)
    assert np.array_equal(g.valid_actions()[idx].cpu().numpy(), z["valid0"])
    for t in range(T):
        f, v = g.step_random(t, weights, record_actions=True)
        assert np.array_equal(g.actions_taken[idx].cpu().numpy().astype(np.uint8), z["actions"][t]), (name, t, "actions")
        assert np.array_equal(f[idx].cpu().numpy(), z["flags"][t]), (name, t, "flags")
        assert np.array_equal(v[idx].cpu().numpy(), z["valid"][t]), (name, t, "valid")
    e = g.export()
    for k in F64_KEYS:
        assert np.array_equal(np.ascontiguousarray(e[k][..., ids]).view(np.uint64), z["final_" + k]), (name, k)
    for k in INT_KEYS:
        assert np.array_equal(e[k][..., ids], z["final_" + k]), (name, k)
    assert g.stats()["steps"] == N * T


@pytest.mark.parametrize("P,dealer", [(2, 1), (4, 2), (9, 8), (9, 3)])
def test_reset_with_dealer_keyword_and_auto_reset(torch_cuda, ora, P, dealer):
    """Game.reset(dealer=k) (game.py:403): the first dealer is first_playing(k + 1); with auto_reset the precomputed
    fresh state must be the one of THAT reset.  Lockstep against the oracle, masked re-resets included."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N, T = 3000, 120
    start = 30 if P < 9 else [12 * (i + 1) for i in range(9)]
    w = ALLIN_HEAVY if P == 9 else None
    for via_setter in (False, True):
        g = GameBatch(N, P, start, 2, 1, dealer=0 if via_setter else dealer, seed=17, auto_reset=True)
        if via_setter:
            g.set_reset_dealer(dealer)
        o = ora.OracleBatch(N, P, start, 2, 1, dealer=dealer, seed=17, auto_reset=True)
        g.reset(); o.reset()
        states_equal(g, o, ("after reset", P, dealer, via_setter))
        exp = g.export()
        assert set(np.unique(exp["dealer"])) == {(dealer + 1) % P}
        for t in range(T):
            g.step_random(t, w)
        o.run_random(0, T, w)
        states_equal(g, o, ("after steps", P, dealer, via_setter))
        assert g.stats()["games"] > 0  # games ended and were re-dealt from the dealer=k fresh state


def test_facade_reset_dealer_keyword(torch_cuda, ora):
    from pokerl_b200 import Game
    g = Game(num_players=4, start_credits=100, seed=3)
    g.reset(dealer=2)
    o = ora.OracleGame(4, 100, 2, 1, dealer=2, seed=3, env_id=0)
    o.reset()
    s = o.snapshot()
    assert (g.dealer_idx, g.small_blind_idx, g.big_blind_idx, g.active_player) == (s["dealer"], s["sb"], s["bb"], s["active"])
    assert g.dealer_idx == 3
    g.reset()  # like the reference, a later reset() without the keyword goes back to dealer 0 (game.py:403)
    assert g.dealer_idx == 1


@pytest.mark.parametrize("P", [2, 4, 9])
def test_batched_state_view_fields_match_the_oracle(torch_cuda, ora, P):
    """Every field and property of the batched Game.StateView (game.py:117-161) for EVERY env, after reset and along
    a random trajectory: player, valid_actions, turn, player_cards, community_cards / num_community, credits, bets,
    pending_bets, minimum_raise_value, pot (numpy's sum order), high_bet, credit, player_hand."""
    torch = torch_cuda
    from pokerl_b200 import Game
    N, seed = 700, 23
    start = 50 if P < 9 else [15 * (i + 1) for i in range(9)]
    g = Game(num_players=P, start_credits=start, num_envs=N, seed=seed, auto_reset=True)
    o = ora.OracleBatch(N, P, start, 2, 1, seed=seed, auto_reset=True)
    g.reset(); o.reset()

    def check(ctx):
        sv = g.active_state
        e = o.export()
        act = e["active"]
        assert np.array_equal(sv.player.cpu().numpy(), act), ctx
        assert np.array_equal(sv.turn.cpu().numpy(), e["turn"]), ctx
        vm = o.valid()
        assert np.array_equal(sv.valid_mask.cpu().numpy(), vm), ctx
        onehot = ((vm[:, None] >> np.arange(7)) & 1).astype(np.float64)
        assert sv.valid_actions.dtype == torch.float64 and np.array_equal(sv.valid_actions.cpu().numpy(), onehot), ctx
        for name, key in (("credits", "credits"), ("bets", "bets"), ("pending_bets", "pending")):
            a = getattr(sv, name).cpu().numpy()
            assert a.shape == (N, P) and np.array_equal(a.view(np.uint64), np.ascontiguousarray(e[key].T).view(np.uint64)), (ctx, name)
        assert np.array_equal(sv.minimum_raise_value.cpu().numpy().view(np.uint64), e["minraise"].view(np.uint64)), ctx
        pot = np.array([np.sum(np.ascontiguousarray(e["bets"][:, i])) for i in range(N)])  # the reference's np.sum(self.bets)
        assert np.array_equal(sv.pot.cpu().numpy().view(np.uint64), pot.view(np.uint64)), ctx
        assert np.array_equal(sv.high_bet.cpu().numpy().view(np.uint64), e["pending"].max(axis=0).view(np.uint64)), ctx
        assert np.array_equal(sv.credit.cpu().numpy().view(np.uint64), e["credits"][act, np.arange(N)].view(np.uint64)), ctx
        deals = np.stack([ora.deal(seed, i, int(e["episode"][i]), int(e["hand"][i]), P) for i in range(N)])  # [N, 5+2P]
        pc = np.stack([deals[np.arange(N), 5 + 2 * act], deals[np.arange(N), 6 + 2 * act]], axis=1)
        assert np.array_equal(sv.player_cards.cpu().numpy(), pc), ctx
        assert np.array_equal(sv.community_cards.cpu().numpy(), deals[:, :5]), ctx
        shown = np.where(e["turn"] == 0, 0, e["turn"] + 2)
        assert np.array_equal(sv.num_community.cpu().numpy(), shown), ctx
        assert np.array_equal(sv.player_hand.cpu().numpy(), np.concatenate([pc, deals[:, :5]], axis=1)), ctx
        assert np.array_equal(g.pot.cpu().numpy().view(np.uint64), pot.view(np.uint64)), ctx

    check("reset")
    for t in range(40):
        g.step_random()
        o.run_random(t, 1)
        if t % 4 == 0 or t > 34:
            check(t)
    # StateView(game, player): the view of a FIXED seat (game.py:122 quirk: seat 0 falls back to the active player only in scalar mode)
    sv1 = Game.StateView(g, 1)
    e = o.export()
    assert np.array_equal(sv1.credit.cpu().numpy().view(np.uint64), e["credits"][1].view(np.uint64))
    assert set(sv1.player.cpu().numpy().tolist()) == {1}


def test_pipelined_and_plain_launch_order_give_the_same_state(torch_cuda, ora):
    """Back-to-back step launches overlap (programmatic dependent launch + per-tile hand-over); the result must be the
    one of strictly serial launches -- on one handle stepped in a tight loop (a batch this small has three launches in
    flight at a time: the hand-over really is per tile) and across stream switches."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N, P, T = 1 << 18, 2, 48
    o = ora.OracleBatch(N, P, 100, 2, 1, seed=4, auto_reset=True)
    o.reset()
    o.run_random(0, T)
    for mode in ("pipelined", "plain", "two streams"):
        g = GameBatch(N, P, 100, 2, 1, seed=4, auto_reset=True)
        g.set_pipelining(mode != "plain")
        g.reset()
        side = torch.cuda.Stream()
        for t in range(T):
            if mode == "two streams" and t % 5 == 4:  # the handle follows the caller to another stream
                with torch.cuda.stream(side):
                    g.step_random(t)
            else:
                g.step_random(t)
        torch.cuda.synchronize()
        states_equal(g, o, mode)
        assert g.stats()["steps"] == N * T


def test_writes_through_the_state_view_are_honoured(torch_cuda, ora):
    """the valid mask cached in `meta` goes stale when the caller writes the planes.  pkb_step checks the
    caller's action against the planes as they are; the sampler sees them after refresh_valid(); a never-reset handle
    folds instead of inventing a move."""
    torch = torch_cuda
    from pokerl_b200 import Game, GameBatch
    from pokerl_b200 import _native as nat
    g = GameBatch(64, 2, 100, 2, 1, seed=1)
    g.reset()
    act = int(g.active_player[0].item())
    g.credits[act].fill_(1.5)  # below the high bet of 2: CALL and the raises are no longer valid (game.py:370-376)
    want = g.valid_actions().clone()
    assert int(want[0].item()) == 0b1000001
    flags, _ = g.step(torch.full((64,), 2, dtype=torch.int32, device=g.device))  # CALL
    assert bool(((flags & nat.F_INVALID) != 0).all().item())
    g.refresh_valid()
    g.step_random(0, record_actions=True)
    assert set(g.actions_taken.cpu().numpy().tolist()) <= {0, 6}
    # scalar facade before reset(): the reference accepts FOLD on the all-zero table (game.py:648)
    h = Game(num_players=3)
    assert h.get_valid_actions()[0].tolist() == [1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    h.step(0)
    raw = GameBatch(32, 3, 100, 2, 1, seed=2)  # never reset: cached mask is empty
    raw.step_random(0, record_actions=True)
    assert set(raw.actions_taken.cpu().numpy().tolist()) == {0}


def test_output_buffers_are_checked(torch_cuda):
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    g = GameBatch(100, 2, 100, 2, 1)
    g.reset()
    a = torch.zeros(100, dtype=torch.int32, device=g.device)
    with pytest.raises(ValueError):
        g.step(a, flags=torch.zeros(50, dtype=torch.uint8, device=g.device))
    with pytest.raises(ValueError):
        g.step(a, valid=torch.zeros(100, dtype=torch.uint8))  # host tensor
    with pytest.raises(ValueError):
        g.step(a[:10])
    f1, _ = g.step(torch.full((100,), 255, dtype=torch.uint8))  # host uint8 keeps its encoding: 255 = skip
    assert bool((f1 == 32).all().item())


def test_packed_step_equals_plain_step(torch_cuda, ora):
    """pkb_step_packed (two moves per byte in, one byte per env out) == pkb_step on the same actions: outputs, state,
    skips; an odd env count exercises the last half-filled byte."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    from pokerl_b200 import _native as nat
    N, P, T = 20001, 3, 40
    a_batch = GameBatch(N, P, 60, 2, 1, seed=8, auto_reset=True)
    b_batch = GameBatch(N, P, 60, 2, 1, seed=8, auto_reset=True)
    a_batch.reset(); b_batch.reset()
    gen = torch.Generator(device=a_batch.device).manual_seed(1)
    for t in range(T):
        acts = a_batch.sample_actions(t).clone()
        skip = torch.rand(N, device=acts.device, generator=gen) < 0.1
        acts = torch.where(skip, torch.full_like(acts, nat.ACTION_SKIP), acts)
        flags, valid = a_batch.step(acts)
        packed = b_batch.step_packed(GameBatch.pack_actions(acts))
        f2, v2 = GameBatch.unpack_outputs(packed)
        assert torch.equal(f2, flags & 7), t
        assert torch.equal(v2, valid), t
        assert bool(((flags & nat.F_SKIPPED) != 0).eq(skip).all().item()), t
    ea, eb = a_batch.export(), b_batch.export()
    for k in F64_KEYS:
        assert np.array_equal(np.ascontiguousarray(ea[k]).view(np.uint64), np.ascontiguousarray(eb[k]).view(np.uint64)), k
    for k in INT_KEYS:
        assert np.array_equal(ea[k], eb[k]), k
    sa, sb = a_batch.stats(), b_batch.stats()
    assert {k: v for k, v in sa.items() if k != "payoff_sum"} == {k: v for k, v in sb.items() if k != "payoff_sum"}
    assert np.allclose(sa["payoff_sum"], sb["payoff_sum"], rtol=1e-9)  # float atomics: the order of the adds is not fixed
    bad = GameBatch.pack_actions(torch.full((N,), 9, dtype=torch.int32, device=a_batch.device))  # 7..14: rejected
    b_batch.step_packed(bad)
    assert b_batch.stats()["invalid"] == N


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_soak_random_tables_pipelined_against_the_oracle(torch_cuda, ora, seed):
    """Random table sizes, unequal stacks, blinds and action mixes, 300 fused random-agent steps on 40,000 envs in one
    tight loop (launches pipelined), with the handle's other entry points called in between (they must order
    themselves against the overlapping launches); final state bit for bit against the oracle, poisoned envs aside."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    from pokerl_b200 import _native as nat
    rng = np.random.default_rng(100 + seed)
    P = int(rng.integers(2, 11))
    start = [int(x) for x in rng.integers(5, 300, P)]
    bb = int(rng.choice([2, 4, 10]))
    sb = int(rng.integers(1, bb + 1))
    weights = None if seed % 2 == 0 else [int(x) for x in rng.integers(1, 50, 7)]
    N, T = 40000, 300
    g = GameBatch(N, P, start, bb, sb, seed=seed, env_id_base=7 * seed, auto_reset=True)
    o = ora.OracleBatch(N, P, start, bb, sb, seed=seed, env_id_base=7 * seed, auto_reset=True)
    g.reset(); o.reset()
    poisoned = torch.zeros(N, dtype=torch.bool, device=g.device)
    for t in range(T):
        flags, _ = g.step_random(t, weights)
        poisoned |= (flags & nat.F_POISONED) != 0
        if t % 37 == 5:
            g.valid_actions()
        if t % 53 == 7:
            g.refresh_valid()
        if t % 101 == 50:
            g.stats()
    o.run_random(0, T, weights)
    keep = ~poisoned.cpu().numpy()
    ge, oe = g.export(), o.export()
    for k in F64_KEYS:
        a, b = np.ascontiguousarray(ge[k]).view(np.uint64)[..., keep], np.ascontiguousarray(oe[k]).view(np.uint64)[..., keep]
        assert np.array_equal(a, b), (P, start, bb, sb, weights, k)
    for k in INT_KEYS:
        assert np.array_equal(np.asarray(ge[k])[..., keep], np.asarray(oe[k])[..., keep]), (P, start, bb, sb, weights, k)
    assert g.stats()["steps"] == N * T and (~keep).sum() < N // 100


def test_save_and_load_state_resume_exactly(torch_cuda, ora):
    """GameBatch.save_state / load_state: after putting a snapshot back, the envs replay the same future (decks and
    agent draws are functions of the counters in the state), whatever happened in between."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    g = GameBatch(30000, 4, 60, 2, 1, seed=6, auto_reset=True)
    g.reset()
    for t in range(20):
        g.step_random(t)
    snap = g.save_state()
    for t in range(20, 50):
        g.step_random(t)
    want = g.export()
    for t in range(50, 70):  # wander off
        g.step_random(t)
    g.load_state(snap)
    for t in range(20, 50):
        g.step_random(t)
    got = g.export()
    for k in F64_KEYS:
        assert np.array_equal(np.ascontiguousarray(got[k]).view(np.uint64), np.ascontiguousarray(want[k]).view(np.uint64)), k
    for k in INT_KEYS:
        assert np.array_equal(got[k], want[k]), k


def test_step_launches_replayed_from_a_cuda_graph(torch_cuda, ora):
    """A loop of pkb_step_random launches captured in a CUDA graph (programmatic edges between them) and replayed: the
    launch index is derived on the device and the first launch of a replay waits for whatever precedes it, so every
    replay continues exactly where the stream left the envs."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N, P, K = 1 << 17, 2, 10
    g = GameBatch(N, P, 100, 2, 1, seed=12, auto_reset=True)
    o = ora.OracleBatch(N, P, 100, 2, 1, seed=12, auto_reset=True)
    g.reset(); o.reset()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        side.wait_stream(torch.cuda.current_stream())
        g.refresh_valid()  # move the handle to the capture stream before capturing (a stream switch needs an event)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for t in range(K):
            g.step_random(t)
    # the capture itself executes nothing
    states_equal(g, o, "after capture")
    for rep in range(3):
        graph.replay()
        if rep == 1:
            torch.cuda.synchronize()
            g.valid_actions()  # a live call between two replays
        o.run_random(0, K)
    torch.cuda.synchronize()
    states_equal(g, o, "after three replays")
    g.step_random(99)  # and a live launch after the graph
    o.run_random(99, 1)
    states_equal(g, o, "live launch after the replays")


@pytest.mark.parametrize("P,K", [(2, 4), (2, 7), (4, 3), (9, 4)])
def test_repeated_steps_in_one_launch_equal_single_steps(torch_cuda, ora, P, K):
    """pkb_step_random_n: K device-sampled steps per tile visit (the tile stays in shared memory between them) leave
    exactly the state, statistics and payoffs of K single pkb_step_random calls -- and of the oracle; the outputs are
    the last step's.  Ragged batch size, unequal stacks and weights for nine players."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N, rounds = 3000, 8
    start = 100 if P != 9 else [20 * (i + 1) for i in range(9)]
    w = None if P != 9 else [10, 30, 30, 5, 5, 5, 45]
    a = GameBatch(N, P, start, 2, 1, seed=5, auto_reset=True)
    b = GameBatch(N, P, start, 2, 1, seed=5, auto_reset=True)
    o = ora.OracleBatch(N, P, start, 2, 1, seed=5, auto_reset=True)
    a.reset(); b.reset(); o.reset()
    for r in range(rounds):
        for k in range(K):
            fa, va = a.step_random(r * K + k, w, record_actions=True)
        acts_a = a.actions_taken.clone()
        fb, vb = b.step_random(r * K, w, record_actions=True, repeat=K)
        assert torch.equal(fa, fb) and torch.equal(va, vb) and torch.equal(acts_a, b.actions_taken), r
    o.run_random(0, rounds * K, w)
    ea, eb, eo = a.export(), b.export(), o.export()
    for k in ea:
        assert np.array_equal(np.ascontiguousarray(ea[k]).view(np.uint8), np.ascontiguousarray(eb[k]).view(np.uint8)), k
    for k in ("credits", "bets", "pending", "payoffs", "minraise"):
        assert np.array_equal(np.ascontiguousarray(eb[k]).view(np.uint64), np.ascontiguousarray(eo[k]).view(np.uint64)), k
    for k in ("states", "turn", "hand", "dealer", "sb", "bb", "active", "episode"):
        assert np.array_equal(eb[k], eo[k]), k
    sa, sb = a.stats(), b.stats()
    assert all(sa[k] == sb[k] for k in ("steps", "hands", "showdowns", "games", "invalid", "poisoned")) and sb["steps"] == N * rounds * K
    # the payoff sums are f64 accumulations whose order differs (per-thread tallies over K steps, then atomics): equal to rounding
    assert np.allclose(sa["payoff_sum"], sb["payoff_sum"], rtol=1e-12, atol=1e-9)
