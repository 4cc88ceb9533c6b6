"""CPU: the oracle (oracle/pokerl_oracle.c) against everything that pins it.

Pins: the reference's own judger known-answer vectors (tests/kat.py), golden sums over all
C(52,5)/C(52,6) hands computed from the reference's judger.py (C(52,7) is
checked by tests/golden/check_exhaustive7.py and in the GPU suite), vectors and step traces recorded
from the LIVE reference by tests/golden/make_golden.py, and the Random123 Philox vectors.
"""
import json
import os

import numpy as np
import pytest

from tests import kat


def test_eval_kats(ora):
    for cards, (cat, kickers) in kat.EVAL_KATS:
        assert ora.unpack(ora.eval_cards(kat.hand_values(cards))) == (cat, kickers), cards


def test_compare_kats(ora):
    for hands, onehot in kat.COMPARE_KATS:
        packed = [ora.eval_cards(kat.hand_values(h)) for h in hands]
        assert ora.compare_rankings(packed) == onehot, hands


def test_eval_small_cases(ora):
    assert ora.unpack(ora.eval_cards([])) == (10, [])                       # judger.py:30
    assert ora.unpack(ora.eval_cards([kat.card_value("AS")])) == (9, [13])  # judger.py:31
    assert ora.unpack(ora.eval_cards(kat.hand_values("7S 7H"))) == (8, [6])
    assert ora.unpack(ora.eval_cards(kat.hand_values("7S KH"))) == (9, [12, 6])


def test_eval_reference_vectors(ora, golden_dir):
    d = np.load(os.path.join(golden_dir, "eval_vectors.npz"))
    for k in range(8):
        sel = d["counts"] == k
        got = ora.eval_cards_batch(d["cards"][sel][:, :k]) if k else np.array(
            [ora.eval_cards([]) for _ in range(sel.sum())], dtype=np.uint32)
        assert np.array_equal(got, d["packed"][sel]), k


def test_compare_reference_vectors(ora, golden_dir):
    rows = json.load(open(os.path.join(golden_dir, "compare_vectors.json")))
    for r in rows:
        assert ora.compare_rankings(r["packed"]) == r["onehot"]


@pytest.mark.parametrize("k", [5, 6])
def test_exhaustive_golden(ora, k):
    assert ora.eval_exhaustive(k) == kat.EXHAUSTIVE_GOLDEN[k]


def test_exhaustive7_sampled_range(ora):
    # full C(52,7) takes ~7 s x cores; here: enumeration order/unranking on two windows
    from math import comb
    a = ora.eval_enumerate(7, 0, 100000)
    b = ora.eval_enumerate(7, comb(52, 7) - 100000, 100000)
    masks = []
    m = (1 << 7) - 1
    for _ in range(1000):
        masks.append(m)
        c = m & -m
        r = m + c
        m = (((r ^ m) >> 2) // c) | r
    assert np.array_equal(a[:1000], ora.eval_masks_batch(np.array(masks, dtype=np.uint64)))
    top = sum(1 << i for i in range(45, 52))
    assert b[-1] == ora.eval_masks_batch(np.array([top], dtype=np.uint64))[0]


def test_philox_kats(ora):
    for ctr, key, out in kat.PHILOX_KATS:
        assert ora.philox(ctr, key) == out


def test_deal_is_a_partial_permutation(ora):
    for P in (2, 4, 9, 10):
        for hand in range(1, 40):
            c = ora.deal(7, 123, 1, hand, P)
            assert len(set(c.tolist())) == 5 + 2 * P
            assert all((v & 0xF) < 13 and (v >> 4) < 4 for v in c.tolist())


def test_np_sum_order(ora):
    rng = np.random.default_rng(3)
    for n in range(1, 24):
        for _ in range(300):
            a = rng.random(n) * rng.choice([1.0, 1e3, 1e-3], n)
            assert ora.np_sum(a) == float(np.sum(a)), n


def _replay(ora, cfg, tr):
    P = cfg["num_players"]
    g = ora.OracleGame(P, cfg["start_credits"], cfg["big_blind"], cfg["small_blind"], seed=cfg["seed"], env_id=0)
    resets = set(int(x) for x in tr["resets"])
    g.reset()
    for t in range(len(tr["actions"])):
        if t in resets and t != 0:
            g.reset()
        f = g.step(int(tr["actions"][t]))
        assert f == int(tr["flags"][t]), (cfg["name"], t, f)
        s = g.snapshot()
        for k in ("credits", "bets", "pending", "payoffs"):
            assert np.array_equal(np.array(s[k]).view(np.uint64), tr[k][t]), (cfg["name"], t, k)
        assert np.array_equal(np.array(s["states"], dtype=np.uint8), tr["states"][t]), (cfg["name"], t)
        assert np.float64(s["minraise"]).view(np.uint64) == tr["minraise"][t]
        assert [s["turn"], s["hand"], s["dealer"], s["sb"], s["bb"], s["active"]] == tr["meta"][t].tolist()
        assert g.valid_mask() == int(tr["valid"][t]), (cfg["name"], t)


def test_reference_traces(ora, golden_dir):
    """Bit-exact replay of step traces recorded from the live reference Game (Philox-dealt decks)."""
    for cfg in json.load(open(os.path.join(golden_dir, "traces.json"))):
        tr = np.load(os.path.join(golden_dir, f"trace_{cfg['name']}.npz"))
        _replay(ora, cfg, tr)


def test_invalid_action_leaves_state_untouched(ora):
    g = ora.OracleGame(4, 100, 2, 1, seed=5)
    g.reset()
    before = g.snapshot()
    assert g.step(1) == ora.F_INVALID  # CHECK facing the blinds is invalid (game.py:375)
    assert g.step(9) == ora.F_INVALID
    assert g.snapshot() == before


def test_batch_matches_scalar(ora):
    b = ora.OracleBatch(64, 4, 100, 2, 1, seed=9, env_id_base=1000, auto_reset=True)
    b.reset()
    singles = []
    for i in range(64):
        g = ora.OracleGame(4, 100, 2, 1, seed=9, env_id=1000 + i)
        g.reset()
        singles.append(g)
    for t in range(200):
        a = b.sample(t)
        flags, valid = b.step(a)
        for i, g in enumerate(singles):
            assert ora.sample_action(9, 1000 + i, t, g.valid_mask()) == a[i]
            f = g.step(int(a[i]))
            assert f == flags[i]
            if f & ora.F_GAME_OVER:
                g.reset()
            assert g.valid_mask() == valid[i]
    ex = b.export()
    for i, g in enumerate(singles):
        s = g.snapshot()
        assert np.array_equal(ex["credits"][:, i], np.array(s["credits"]))
        assert ex["hand"][i] == s["hand"] and ex["active"][i] == s["active"]
