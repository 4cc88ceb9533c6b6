"""GPU (-m gpu): the CUDA Game path (K1 step, K2 end_hand/showdown, K3 deal/reset, K5 sampler)
through the C ABI, bit-exact against (a) step traces recorded from the live reference and (b) the
CPU oracle run in lockstep on the same seeds."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALLIN_HEAVY = [100, 300, 300, 50, 50, 50, 450]
PASSIVE = [20, 400, 400, 60, 60, 60, 20]
F64_KEYS = ("credits", "bets", "pending", "payoffs", "minraise")
INT_KEYS = ("states", "turn", "hand", "dealer", "sb", "bb", "active", "episode")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def assert_state_equal(gpu: dict, cpu: dict, skip=None, ctx=""):
    """Bit-exact comparison of exported states; `skip` = bool[N] envs excluded (poisoned ones)."""
    keep = slice(None) if skip is None else ~skip
    for k in F64_KEYS:
        a = np.ascontiguousarray(gpu[k]).view(np.uint64)[..., keep]
        b = np.ascontiguousarray(cpu[k]).view(np.uint64)[..., keep]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[0]
            raise AssertionError(f"{ctx}: {k} differs at {bad.tolist()}: gpu={gpu[k][..., keep][tuple(bad)]!r} "
                                 f"cpu={cpu[k][..., keep][tuple(bad)]!r}")
    for k in INT_KEYS:
        a, b = np.asarray(gpu[k])[..., keep], np.asarray(cpu[k])[..., keep]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[0]
            raise AssertionError(f"{ctx}: {k} differs at {bad.tolist()}: gpu={a[tuple(bad)]} cpu={b[tuple(bad)]}")


def test_reference_traces_bit_exact(torch_cuda, golden_dir):
    """Replay the action sequences recorded from the live reference Game (tests/golden/make_golden.py)
    through pkb_step and compare every field after every step (float bits, states, indices, flags,
    valid masks)."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    for cfg in json.load(open(os.path.join(golden_dir, "traces.json"))):
        tr = np.load(os.path.join(golden_dir, f"trace_{cfg['name']}.npz"))
        P = cfg["num_players"]
        g = GameBatch(1, P, cfg["start_credits"], cfg["big_blind"], cfg["small_blind"], seed=cfg["seed"], env_id_base=0)
        g.reset()
        resets = set(int(x) for x in tr["resets"])
        for t in range(len(tr["actions"])):
            if t in resets and t != 0:
                g.reset()
            flags, valid = g.step(torch.tensor([int(tr["actions"][t])], dtype=torch.int32, device="cuda"))
            s = g.export()
            ctx = (cfg["name"], t)
            assert int(flags.item()) == int(tr["flags"][t]), ctx
            assert int(valid.item()) == int(tr["valid"][t]), ctx
            for k in ("credits", "bets", "pending", "payoffs"):
                assert np.array_equal(np.ascontiguousarray(s[k][:, 0]).view(np.uint64), tr[k][t]), (ctx, k)
            assert np.array_equal(s["states"][:, 0], tr["states"][t]), ctx
            assert s["minraise"].view(np.uint64)[0] == tr["minraise"][t], ctx
            meta = [int(s[k][0]) for k in ("turn", "hand", "dealer", "sb", "bb", "active")]
            assert meta == tr["meta"][t].tolist(), ctx


CONFIGS = [
    # P, start_credits, bb, sb, weights, N, T
    (2, 100, 2, 1, None, 8192, 160),
    (2, 7, 2, 1, None, 4096, 80),
    (3, 3, 2, 1, None, 4096, 80),
    (4, 100, 2, 1, None, 8192, 160),
    (4, 1000, 40, 20, None, 4096, 120),
    (5, 37, 2, 1, PASSIVE, 4096, 120),
    (6, 100, 2, 1, PASSIVE, 4096, 160),
    (7, [5, 10, 20, 40, 80, 160, 320], 2, 1, None, 4096, 120),
    (8, 100, 2, 1, PASSIVE, 4096, 160),
    (9, [20, 40, 60, 80, 100, 120, 140, 160, 180], 2, 1, ALLIN_HEAVY, 8192, 200),
    (9, 100, 2, 1, PASSIVE, 4096, 160),
    (10, 50, 2, 1, None, 4096, 120),
    (9, 3, 40, 20, None, 2048, 80),  # stacks below the blinds: overbets, negative stacks, resurrection
    (6, 1, 3, 3, None, 2048, 80),
]


@pytest.mark.parametrize("P,start,bb,sb,weights,N,T", CONFIGS)
def test_lockstep_vs_oracle(torch_cuda, ora, P, start, bb, sb, weights, N, T):
    """N envs x T steps: device sampler + pkb_step vs the oracle, every step: actions, flags and valid
    masks; every 8th step and at the end: the whole state, bit for bit.  auto_reset on."""
    from pokerl_b200 import GameBatch
    seed, base = 1234 + P, 5000
    g = GameBatch(N, P, start, bb, sb, seed=seed, env_id_base=base, auto_reset=True)
    o = ora.OracleBatch(N, P, start, bb, sb, seed=seed, env_id_base=base, auto_reset=True)
    g.reset()
    o.reset()
    assert_state_equal(g.export(), o.export(), ctx="after reset")
    assert np.array_equal(g.valid_actions().cpu().numpy(), o.valid())
    poisoned = np.zeros(N, dtype=bool)
    for t in range(T):
        a_gpu = g.sample_actions(t, weights)
        a_cpu = o.sample(t, weights)
        assert np.array_equal(a_gpu.cpu().numpy()[~poisoned], a_cpu[~poisoned]), f"sampler step {t}"
        flags, valid = g.step(a_gpu)
        f_cpu, v_cpu = o.step(a_cpu)
        f_gpu = flags.cpu().numpy()
        poisoned |= (f_cpu & ora.F_POISONED) != 0
        assert np.array_equal(f_gpu[~poisoned], f_cpu[~poisoned]), f"flags step {t}"
        assert np.array_equal((f_gpu & ora.F_POISONED) != 0, (f_cpu & ora.F_POISONED) != 0), f"poison step {t}"
        assert np.array_equal(valid.cpu().numpy()[~poisoned], v_cpu[~poisoned]), f"valid step {t}"
        if t % 8 == 7 or t == T - 1:
            assert_state_equal(g.export(), o.export(), skip=poisoned, ctx=f"step {t}")
    st = g.stats()
    assert st["steps"] + st["invalid"] == N * T and st["invalid"] == 0
    assert st["hands"] > 0 and st["showdowns"] > 0


def test_fused_random_step_equals_sampler_plus_step(torch_cuda, ora):
    """pkb_step_random (sampler fused into the step kernel) == pkb_sample_actions + pkb_step == oracle."""
    from pokerl_b200 import GameBatch
    N, P, T = 16384, 2, 100
    a = GameBatch(N, P, 100, 2, 1, seed=3, auto_reset=True)
    b = GameBatch(N, P, 100, 2, 1, seed=3, auto_reset=True)
    o = ora.OracleBatch(N, P, 100, 2, 1, seed=3, auto_reset=True)
    for x in (a, b, o):
        x.reset()
    for t in range(T):
        fa, va = a.step_random(t, record_actions=True)
        acts = b.sample_actions(t)
        fb, vb = b.step(acts)
        assert (a.actions_taken == acts).all() and (fa == fb).all() and (va == vb).all(), t
    o.run_random(0, T)
    assert_state_equal(a.export(), o.export(), ctx="fused")
    assert_state_equal(b.export(), o.export(), ctx="unfused")


def test_invalid_actions_leave_env_untouched(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N = 1024
    g = GameBatch(N, 4, 100, 2, 1, seed=8)
    g.reset()
    before = g.export()
    valid0 = g.valid_actions().cpu().numpy()
    acts = np.where(np.arange(N) % 2 == 0, 1, 2).astype(np.int32)  # CHECK is invalid facing the blinds, CALL is fine
    acts[5] = 9
    acts[7] = -1
    flags, valid = g.step(torch.from_numpy(acts).cuda())
    f = flags.cpu().numpy()
    bad = (np.arange(N) % 2 == 0) | np.isin(np.arange(N), [5, 7])
    assert ((f & ora.F_INVALID) != 0).tolist() == bad.tolist()
    after = g.export()
    for k in F64_KEYS + INT_KEYS:
        assert np.array_equal(np.asarray(after[k])[..., bad], np.asarray(before[k])[..., bad]), k
    assert np.array_equal(valid.cpu().numpy()[bad], valid0[bad])
    assert g.stats()["invalid"] == int(bad.sum())


def test_masked_reset_and_env_sharding(torch_cuda, ora):
    """Results depend on the GLOBAL env id only: two half-size shards == one full batch."""
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    N, P, T = 4096, 4, 60
    full = GameBatch(N, P, 100, 2, 1, seed=21, env_id_base=0, auto_reset=True)
    lo = GameBatch(N // 2, P, 100, 2, 1, seed=21, env_id_base=0, auto_reset=True)
    hi = GameBatch(N // 2, P, 100, 2, 1, seed=21, env_id_base=N // 2, auto_reset=True)
    for x in (full, lo, hi):
        x.reset()
    for t in range(T):
        for x in (full, lo, hi):
            x.step_random(t)
    e, el, eh = full.export(), lo.export(), hi.export()
    for k in F64_KEYS + INT_KEYS:
        assert np.array_equal(np.asarray(e[k])[..., :N // 2], np.asarray(el[k])), k
        assert np.array_equal(np.asarray(e[k])[..., N // 2:], np.asarray(eh[k])), k
    # masked reset: only the selected envs change
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
    mask[::3] = 1
    full.reset(mask)
    e2 = full.export()
    sel = mask.cpu().numpy().astype(bool)
    assert (e2["hand"][sel] == 1).all() and (e2["episode"][sel] == e["episode"][sel] + 1).all()
    for k in F64_KEYS + INT_KEYS:
        assert np.array_equal(np.asarray(e2[k])[..., ~sel], np.asarray(e[k])[..., ~sel]), k


def test_deal_replay_and_features(torch_cuda, ora):
    """pkb_cards == host pkb_deal == oracle deal for the envs' (episode, hand); hand features == oracle eval
    of hole + visible board on 2/5/6/7 cards (q_learning.py:29-33)."""
    from pokerl_b200 import GameBatch
    from pokerl_b200 import _native as nat
    N, P = 2048, 4
    g = GameBatch(N, P, 100, 2, 1, seed=99, env_id_base=10, auto_reset=True)
    g.reset()
    for t in range(37):
        g.step_random(t, PASSIVE)
    cards = g.cards().cpu().numpy()
    e = g.export()
    feats = g.hand_features().cpu().numpy().view(np.uint32)
    feats0 = g.hand_features(0).cpu().numpy().view(np.uint32)
    turns = set()
    for i in range(0, N, 7):
        want = ora.deal(99, 10 + i, int(e["episode"][i]), int(e["hand"][i]), P)
        assert np.array_equal(cards[:, i], want)
        assert nat.deal(99, 10 + i, int(e["episode"][i]), int(e["hand"][i]), P) == want.tolist()
        turn = int(e["turn"][i])
        turns.add(turn)
        board = want[: (0 if turn == 0 else turn + 2)]
        for who, f in ((int(e["active"][i]), feats[i]), (0, feats0[i])):
            hole = want[5 + 2 * who: 7 + 2 * who]
            assert f == ora.eval_cards(np.concatenate([hole, board])), (i, turn, who)
    assert len(turns) >= 3


def test_full_size_heads_up_sampled_parity(torch_cuda, ora):
    """BASELINE config 3 at full size (2^20 envs, heads-up): run T fused random steps, then compare two
    contiguous env windows against the oracle bit-exactly, plus conservation-style properties on all envs."""
    from pokerl_b200 import GameBatch
    N, P, T = 1 << 20, 2, 96
    g = GameBatch(N, P, 100, 2, 1, seed=0, auto_reset=True)
    g.reset()
    for t in range(T):
        g.step_random(t)
    e = g.export()
    for base in (0, 777_216):
        W = 4096
        o = ora.OracleBatch(W, P, 100, 2, 1, seed=0, env_id_base=base, auto_reset=True)
        o.reset()
        o.run_random(0, T)
        win = {k: np.asarray(v)[..., base:base + W] for k, v in e.items()}
        assert_state_equal(win, o.export(), ctx=f"window {base}")
    st = g.stats()
    assert st["steps"] == N * T and st["poisoned"] == 0 and st["invalid"] == 0
    assert (e["states"] <= 4).all() and (e["turn"] <= 3).all() and (e["active"] < P).all()
    # uniform random play at 100 credits / blinds 2-1 never overbets: no money is created
    total = e["credits"].sum(axis=0) + e["bets"].sum(axis=0)
    assert (total <= 200.0 + 1e-9).all()


def test_full_size_nine_player_side_pot_heavy_sampled_parity(torch_cuda, ora):
    """BASELINE config 4 at its full size (2^22 envs, nine players, unequal stacks 20..180, action mix that is 45 %
    all-in so that multi-layer side pots are common): T fused weighted-random steps, then env windows at both ends and
    in the middle compared bit-exactly with the oracle (n >= 8 np.sum order, side-pot layers), plus whole-batch
    properties."""
    from pokerl_b200 import GameBatch
    N, P, T = 1 << 22, 9, 40
    start = [20 * (i + 1) for i in range(P)]
    weights = [10, 30, 30, 5, 5, 5, 45]  # FOLD .10, CHECK/CALL .30, RAISE_x .05 each, ALL_IN .45
    g = GameBatch(N, P, start, 2, 1, seed=1, auto_reset=True)
    g.reset()
    for t in range(T):
        g.step_random(t, weights)
    e = g.export()
    for base in (0, 2_000_000, N - 1024):
        W = 1024
        o = ora.OracleBatch(W, P, start, 2, 1, seed=1, env_id_base=base, auto_reset=True)
        o.reset()
        o.run_random(0, T, weights)
        win = {k: np.asarray(v)[..., base:base + W] for k, v in e.items()}
        assert_state_equal(win, o.export(), ctx=f"window {base}")
    st = g.stats()
    assert st["steps"] == N * T and st["invalid"] == 0 and st["showdowns"] > N
    assert (e["states"] <= 4).all() and (e["turn"] <= 3).all() and (e["active"] < P).all()
    # chips are only ever destroyed (the reference's skipped-street quirk), never created
    total = e["credits"].sum(axis=0) + e["bets"].sum(axis=0)
    assert (total <= sum(start) + 1e-9).all()


def test_scalar_facade_random_game(torch_cuda, ora):
    """examples/random_game.py shape: Game(num_players=4, start_credits=1000, big_blind=40, small_blind=20) played to
    the end through the reference-style scalar API, mirrored step by step by the oracle."""
    import pokerl_b200
    rng = np.random.default_rng(4)
    game = pokerl_b200.Game(num_players=4, start_credits=1000, big_blind=40, small_blind=20, seed=77)
    o = ora.OracleGame(4, 1000, 40, 20, seed=77, env_id=0)
    game.reset()
    o.reset()
    done, steps = False, 0
    while not done and steps < 5000:
        s = game.active_state
        idx = list(s.valid_action_indices)
        assert sum(1 << a for a in idx) == o.valid_mask()
        a = int(rng.choice(idx))
        done, hand_over, round_over = game.step(a)
        f = o.step(a)
        assert (done, hand_over, round_over) == (bool(f & 1), bool(f & 2), bool(f & 4))
        snap = o.snapshot()
        assert np.array_equal(game.credits, np.array(snap["credits"]))
        assert np.array_equal(game.payoffs, np.array(snap["payoffs"]))
        assert game.active_player == snap["active"] and game.turn == snap["turn"] and game.hand == snap["hand"]
        steps += 1
    assert done and game.game_over
    with pytest.raises(NotImplementedError):
        game.step(2.0)
    before = game.credits.copy()
    with pytest.raises((ValueError, IndexError)):
        game.step(7)
    assert np.array_equal(game.credits, before)


@pytest.mark.gpu
def test_step_u8_equals_step_i32():
    """pkb_step_u8 (one byte per action, 255 = skip, anything else out of range = rejected) steps exactly like pkb_step."""
    import torch
    from pokerl_b200 import GameBatch
    N, P = 3000, 3
    a = GameBatch(N, P, 60, 2, 1, seed=5, auto_reset=True)
    b = GameBatch(N, P, 60, 2, 1, seed=5, auto_reset=True)
    a.reset(); b.reset()
    g = torch.Generator(device="cpu").manual_seed(1)
    for t in range(40):
        acts = a.sample_actions(t).clone()
        r = torch.rand(N, generator=g)
        acts[(r < 0.1).to(acts.device)] = -2147483648          # skip
        acts[((r >= 0.1) & (r < 0.15)).to(acts.device)] = 9     # out of range -> invalid
        u8 = acts.clone()
        u8[acts == -2147483648] = 255
        fa, va = a.step(acts)
        fb, vb = b.step(u8.to(torch.uint8))
        assert torch.equal(fa, fb) and torch.equal(va, vb)
    ea, eb = a.export(), b.export()
    for k in ea:
        assert np.array_equal(np.asarray(ea[k]).view(np.uint8), np.asarray(eb[k]).view(np.uint8)), k
    sa, sb = a.stats(), b.stats()
    for k in sa:  # the counters exactly; the payoff sums are float atomics accumulated in launch order
        if k == "payoff_sum":
            assert np.allclose(sa[k], sb[k], rtol=0, atol=1e-6)
        else:
            assert sa[k] == sb[k], k


@pytest.mark.gpu
def test_seat_masks_decode_to_player_states():
    """The zero-copy seat-mask words (pkb_state_view.seat_masks: ACTIVE | CALLED<<16 | ALL_IN<<32 | BROKEN<<48) decode to
    exactly what pkb_player_states returns (the reference's one PlayerState value per seat, enums.py:130-134)."""
    from pokerl_b200 import GameBatch
    N, P = 5000, 6
    g = GameBatch(N, P, [3, 9, 30, 50, 80, 120], 4, 2, seed=8, auto_reset=False)
    g.reset()
    for t in range(60):
        g.step_random(t, [10, 30, 30, 5, 5, 5, 45])
    m = g.seat_masks.cpu().numpy().astype(np.uint64)
    st = g.player_states.cpu().numpy()
    seen = set()
    for p in range(P):
        dec = sum(((m >> np.uint64(16 * k + p)) & np.uint64(1)).astype(np.uint8) * v for k, v in enumerate((1, 2, 3, 4)))
        assert np.array_equal(dec, st[p]), p
        seen |= set(np.unique(st[p]).tolist())
    assert seen == {0, 1, 2, 3, 4}  # every PlayerState occurs in this run
    assert not (m >> np.uint64(16 + P) & np.uint64((1 << (16 - P)) - 1)).any()  # no bits beyond the table's seats
