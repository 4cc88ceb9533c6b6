"""Generate golden fixtures from the LIVE reference (sneppy/pokerl at /root/reference) and
differentially validate the C oracle against it.

Runs only in the build container (the GPU box has no /root/reference); its outputs
(tests/golden/*.npz, *.json) are committed and are what the tests read.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py            # write fixtures
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py --validate 300000   # long differential run

Deck injection: the reference calls random.shuffle(self.deck) once per setup_hand
(game.py:424); we replace the `random` name inside pokerl.game with a shim whose
shuffle() overwrites the deck with the next queued deal.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import pokerl  # noqa: E402  (the reference)
import pokerl.game as ref_game  # noqa: E402
from pokerl.cards import Card  # noqa: E402
from pokerl.judger import compare_rankings, eval_hand, get_kickers_value  # noqa: E402

from oracle import pyoracle as ora  # noqa: E402


class StableArgsortNumpy:
    """Proxy for the `np` name inside pokerl.game that pins np.argsort's tie order.

    game.py:495 calls np.argsort(bets) with the default kind ("quicksort"), whose order among EQUAL
    bets is unspecified and differs between numpy builds: numpy 1.18.4 (the reference's pinned
    version, requirements.txt:4) insertion-sorts n <= 16 elements, i.e. keeps ascending seat order;
    numpy >= 2.0 on AVX-512/AVX2 hosts dispatches to x86-simd-sort, which permutes ties.  The order
    matters only when the last two showdown seats have equal bets and a folded seat has bet more
    (game.py:500-505 gives the leftover to whichever of the two is visited second).  The oracle and
    the CUDA path follow the pinned numpy (ascending seat order); fixtures are generated with the
    same order so that they do not depend on the CPU this script happens to run on.
    """

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def argsort(self, a, *args, **kw):
        kw.setdefault("kind", "stable")
        return self._real.argsort(a, *args, **kw)


NATIVE_ARGSORT = False


class NonTerminating(Exception):
    """The reference's `while` at game.py:607 has no termination guarantee."""


class DeckShim:
    """Stands in for the `random` module inside pokerl.game: deals hand `game.hand` of the current episode."""

    def __init__(self, seed, num_players):
        self.seed, self.P = seed, num_players
        self.game = None
        self.episode = 0
        self.in_step = 0

    def shuffle(self, deck):
        self.in_step += 1
        if self.in_step > 40:  # > ORA_NEXT_TURN_CAP / 4 whole hands inside one step: the reference never returns
            raise NonTerminating()
        first = [int(x) for x in ora.deal(self.seed, 0, self.episode, self.game.hand, self.P)]
        rest = [v for v in ALL_VALUES if v not in set(first)]
        deck[:] = [Card(v) for v in first + rest]


ALL_VALUES = [(s << 4) | r for r in range(13) for s in range(4)]


def f64bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64).copy()


def ref_snapshot(g):
    return dict(
        states=np.array(g.player_states, dtype=np.uint8), credits=f64bits(g.credits), bets=f64bits(g.bets),
        pending=f64bits(g.pending_bets), payoffs=f64bits(g.payoffs),
        minraise=f64bits([g.minimum_raise_value])[0], turn=g.turn, hand=g.hand, dealer=int(g.dealer_idx),
        sb=int(g.small_blind_idx), bb=int(g.big_blind_idx), active=int(g.active_player))


def ora_snapshot(o):
    s = o.snapshot()
    return dict(
        states=np.array(s["states"], dtype=np.uint8), credits=f64bits(s["credits"]), bets=f64bits(s["bets"]),
        pending=f64bits(s["pending"]), payoffs=f64bits(s["payoffs"]), minraise=f64bits([s["minraise"]])[0],
        turn=s["turn"], hand=s["hand"], dealer=s["dealer"], sb=s["sb"], bb=s["bb"], active=s["active"])


def snap_equal(a, b):
    for k in a:
        if isinstance(a[k], np.ndarray):
            if not np.array_equal(a[k], b[k]):
                return k
        elif int(a[k]) != int(b[k]):
            return k
    return None


def valid_mask_of(g):
    onehot, _ = g.get_valid_actions()
    return int(sum(1 << i for i, v in enumerate(onehot) if v))


def run_trace(P, start, bb, sb, seed, steps, weights, check_oracle=True, record=True):
    """Play `steps` reference steps (auto-resetting on game over) with Philox-dealt injected decks.

    Returns the recorded trace; raises on the first oracle/reference divergence.
    """
    shim = DeckShim(seed, P)
    ref_game.random = shim
    ref_game.np = np if NATIVE_ARGSORT else StableArgsortNumpy(np)
    rng = np.random.default_rng(seed)
    g = pokerl.Game(num_players=P, start_credits=start, big_blind=bb, small_blind=sb)
    shim.game = g
    o = ora.OracleGame(P, start, bb, sb, seed=seed, env_id=0)

    rec = dict(actions=[], flags=[], valid=[], states=[], credits=[], bets=[], pending=[], payoffs=[], minraise=[],
               meta=[], resets=[])
    # reset
    shim.episode += 1
    g.reset()
    o.reset()
    assert snap_equal(ref_snapshot(g), ora_snapshot(o)) is None, "reset mismatch"
    rec["resets"].append(0)
    w = np.asarray(weights, dtype=np.float64)
    for t in range(steps):
        vm = valid_mask_of(g)
        if check_oracle:
            assert vm == o.valid_mask(), (t, vm, o.valid_mask())
        acts = [a for a in range(7) if vm >> a & 1]
        pw = w[acts] / w[acts].sum()
        a = int(rng.choice(acts, p=pw))
        shim.in_step = 0
        try:
            done, hand_over, round_over = g.step(a)
        except NonTerminating:
            # the oracle (and the CUDA path) cap the loop and flag the env instead of hanging
            assert o.step(a) & ora.F_POISONED, "reference does not terminate but oracle did not flag it"
            STATS["nonterminating"] += 1
            break
        flags = (1 if done else 0) | (2 if hand_over else 0) | (4 if round_over else 0)
        if check_oracle:
            fo = o.step(a)
            assert fo == flags, ("flags", t, fo, flags)
            bad = snap_equal(ref_snapshot(g), ora_snapshot(o))
            assert bad is None, ("state", t, bad, ref_snapshot(g), ora_snapshot(o))
        if record:
            s = ref_snapshot(g)
            rec["actions"].append(a)
            rec["flags"].append(flags)
            rec["valid"].append(valid_mask_of(g))
            for k in ("states", "credits", "bets", "pending", "payoffs"):
                rec[k].append(s[k])
            rec["minraise"].append(s["minraise"])
            rec["meta"].append([s["turn"], s["hand"], s["dealer"], s["sb"], s["bb"], s["active"]])
        if done:
            shim.episode += 1
            g.reset()
            o.reset()
            if check_oracle:
                assert snap_equal(ref_snapshot(g), ora_snapshot(o)) is None, "reset mismatch"
            if record:
                rec["resets"].append(t + 1)
    if not record:
        return None
    out = {k: np.array(v) for k, v in rec.items()}
    out["actions"] = out["actions"].astype(np.int8)
    out["flags"] = out["flags"].astype(np.uint8)
    out["valid"] = out["valid"].astype(np.uint8)
    out["meta"] = out["meta"].astype(np.int32)
    out["resets"] = out["resets"].astype(np.int32)
    return out


STATS = {"nonterminating": 0}
UNIFORM = [1, 1, 1, 1, 1, 1, 1]
ALLIN_HEAVY = [100, 300, 300, 50, 50, 50, 450]  # SURVEY section 8(d) config 4 action mix
PASSIVE = [20, 400, 400, 60, 60, 60, 20]

TRACE_CONFIGS = [
    # name, P, start_credits, bb, sb, seed, steps, weights
    ("p2_100", 2, 100, 2, 1, 11, 1500, UNIFORM),
    ("p2_7", 2, 7, 2, 1, 12, 600, UNIFORM),
    ("p3_3", 3, 3, 2, 1, 13, 600, UNIFORM),
    ("p4_100", 4, 100, 2, 1, 14, 1500, UNIFORM),
    ("p4_1000_b40", 4, 1000, 40, 20, 15, 1200, UNIFORM),
    ("p4_37_passive", 4, 37, 2, 1, 16, 1200, PASSIVE),
    ("p6_100_passive", 6, 100, 2, 1, 17, 1200, PASSIVE),
    ("p8_100", 8, 100, 2, 1, 21, 800, PASSIVE),
    ("p9_ladder_allin", 9, [20, 40, 60, 80, 100, 120, 140, 160, 180], 2, 1, 18, 1500, ALLIN_HEAVY),
    ("p9_100", 9, 100, 2, 1, 19, 1200, PASSIVE),
    ("p10_50", 10, 50, 2, 1, 20, 800, UNIFORM),
]


def make_eval_vectors(rng, n_per_k=600):
    """Random hands (with and without duplicate cards) through the reference's eval_hand."""
    cards, counts, packed = [], [], []
    for k in range(0, 8):
        for i in range(n_per_k if k >= 3 else 60):
            if i % 5 == 4 and k >= 2:  # allow duplicates, like the reference's own tests do
                vals = rng.choice(ALL_VALUES, size=k, replace=True)
            else:
                vals = rng.choice(ALL_VALUES, size=k, replace=False)
            cat, kick = eval_hand([Card(int(v)) for v in vals])
            row = np.zeros(7, dtype=np.uint8)
            row[:k] = vals
            cards.append(row)
            counts.append(k)
            packed.append((cat << 20) | get_kickers_value(kick))
    return np.array(cards), np.array(counts, dtype=np.uint8), np.array(packed, dtype=np.uint32)


def make_compare_vectors(rng, n=400):
    """Random ranking lists through the reference's compare_rankings (incl. the line-148 behaviour)."""
    rows = []
    for _ in range(n):
        P = int(rng.integers(2, 11))
        packed, rankings = [], []
        for _ in range(P):
            if rng.random() < 0.25:
                rankings.append((10, []))
                packed.append(10 << 20)
            else:
                cat = int(rng.choice([8, 8, 9, 9, 7, 5, 4]))
                kick = [int(x) for x in rng.integers(1, 4, size=int(rng.integers(1, 4)))]
                rankings.append((cat, kick))
                packed.append((cat << 20) | get_kickers_value(kick))
        onehot, _ = compare_rankings(rankings)
        rows.append(dict(packed=packed, onehot=onehot))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--validate", type=int, default=0, help="extra differential steps (not recorded)")
    ap.add_argument("--native-argsort", action="store_true",
                    help="leave np.argsort's tie order to this host's numpy and COUNT diverging configs")
    args = ap.parse_args()
    global NATIVE_ARGSORT
    NATIVE_ARGSORT = args.native_argsort

    import logging
    logging.getLogger().setLevel(logging.CRITICAL)

    if args.validate:
        rng = np.random.default_rng(12345)
        total = 0
        rnd = 0
        while total < args.validate:
            P = int(rng.choice([2, 2, 3, 4, 4, 5, 6, 7, 8, 9, 9, 10]))
            kind = int(rng.integers(0, 5))
            if kind == 0:
                start = int(rng.choice([100, 37, 7, 3, 1000, 2, 1]))
            elif kind == 1:
                start = [int(x) for x in rng.integers(1, 200, size=P)]
            elif kind == 2:
                start = np.array([20.0 * (i + 1) for i in range(P)])
            else:
                start = int(rng.choice([100, 50, 20]))
            bb, sb = [(2, 1), (2, 1), (40, 20), (10, 5), (3, 3)][int(rng.integers(0, 5))]
            weights = [UNIFORM, ALLIN_HEAVY, PASSIVE][int(rng.integers(0, 3))]
            steps = 2000
            try:
                run_trace(P, start, bb, sb, int(rng.integers(1, 1 << 30)), steps, weights, record=False)
            except AssertionError as e:
                if not NATIVE_ARGSORT:
                    raise
                STATS["diverged_configs"] = STATS.get("diverged_configs", 0) + 1
                print("diverged (native argsort tie order):", P, start, bb, sb, e.args[0][:3], flush=True)
            total += steps
            rnd += 1
            if rnd % 10 == 0:
                print(f"validated {total} steps over {rnd} configs", flush=True)
        print(f"ORACLE == REFERENCE bit-exact over {total} steps / {rnd} random configs; "
              f"non-terminating reference steps flagged: {STATS['nonterminating']}; "
              f"configs diverged: {STATS.get('diverged_configs', 0)}")
        return

    rng = np.random.default_rng(2026)
    cards, counts, packed = make_eval_vectors(rng)
    mine = np.array([ora.eval_cards(c[:k]) for c, k in zip(cards, counts)], dtype=np.uint32)
    assert np.array_equal(mine, packed), "oracle eval != reference eval_hand"
    np.savez_compressed(os.path.join(HERE, "eval_vectors.npz"), cards=cards, counts=counts, packed=packed)
    print("eval_vectors:", len(packed))

    rows = make_compare_vectors(rng)
    for r in rows:
        assert ora.compare_rankings(r["packed"]) == r["onehot"], r
    with open(os.path.join(HERE, "compare_vectors.json"), "w") as f:
        json.dump(rows, f)
    print("compare_vectors:", len(rows))

    manifest = []
    for name, P, start, bb, sb, seed, steps, weights in TRACE_CONFIGS:
        tr = run_trace(P, start, bb, sb, seed, steps, weights)
        np.savez_compressed(os.path.join(HERE, f"trace_{name}.npz"), **tr)
        manifest.append(dict(name=name, num_players=P, start_credits=start, big_blind=bb, small_blind=sb, seed=seed,
                             steps=steps, weights=weights))
        print("trace", name, "steps", steps, "resets", len(tr["resets"]), "hands", int(tr["meta"][:, 1].sum() and 0) or "")
    with open(os.path.join(HERE, "traces.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
