"""Game.StateView wire format (reference game.py:208-240): pickles written by the live reference
(tests/golden/make_stateview_pickle.py) load into this package's StateView field for field, and the pickle this
package writes loads back -- and, where a copy of the reference is importable (baseline/_ref or /root/reference),
into the reference's own class."""
import importlib
import json
import os
import pickle
import sys

import numpy as np
import pytest

from pokerl_b200 import wire
from pokerl_b200.cards import Card
from pokerl_b200.game import Game

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    blobs = pickle.load(open(os.path.join(GOLDEN, "stateview_ref.pkl"), "rb"))
    meta = json.load(open(os.path.join(GOLDEN, "stateview_ref.json")))
    return blobs, meta["views"]


def check_view(sv, want):
    assert int(sv.player) == want["player"] and sv.num_players == want["num_players"] and sv.turn == want["turn"]
    assert np.asarray(sv.valid_actions, dtype=np.float64).tolist() == want["valid_actions"]
    assert [c.value for c in sv.player_cards] == want["player_cards"]
    assert [c.value for c in sv.community_cards] == want["community_cards"]
    for k in ("credits", "bets", "pending_bets"):
        assert np.asarray(getattr(sv, k)).tolist() == want[k], k
    assert float(sv.minimum_raise_value) == want["minimum_raise_value"]
    assert float(sv.pot) == want["pot"] and float(sv.high_bet) == want["high_bet"] and float(sv.credit) == want["credit"]
    assert list(sv.valid_action_indices) == [a for a, v in enumerate(want["valid_actions"]) if v]
    assert [c.value for c in sv.player_hand] == want["player_cards"] + want["community_cards"]


def test_reference_pickles_load_into_this_package():
    blobs, views = fixtures()
    assert len(blobs) == 3
    for blob, want in zip(blobs, views):
        sv = wire.loads(blob)
        assert type(sv) is Game.StateView and all(type(c) is Card for c in sv.player_cards + sv.community_cards)
        check_view(sv, want)
        str(sv)  # the reference's report (game.py:184-206) formats from the loaded fields


def test_round_trip_through_our_pickle():
    blobs, views = fixtures()
    for blob, want in zip(blobs, views):
        sv = wire.loads(wire.dumps(wire.loads(blob)))
        check_view(sv, want)
        check_view(pickle.loads(pickle.dumps(wire.loads(blob))), want)  # plain pickle of our own class


def _reference_module():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "pokerl")):
            sys.path.insert(0, cand)
            try:
                return importlib.import_module("pokerl")
            finally:
                sys.path.remove(cand)
    return None


def test_the_reference_loads_what_this_package_writes():
    ref = _reference_module()
    if ref is None:
        pytest.skip("no copy of the reference importable here (baseline/_ref or /root/reference)")
    blobs, views = fixtures()
    for blob, want in zip(blobs, views):
        theirs = pickle.loads(wire.dumps(wire.loads(blob)))  # plain pickle.loads, as pokerl/network/client.py would
        assert type(theirs).__module__ == "pokerl.game" and type(theirs).__qualname__ == "Game.StateView"
        assert all(type(c).__module__ == "pokerl.cards" for c in theirs.player_cards)
        check_view(theirs, want)


def test_batched_views_have_no_wire_format():
    sv = Game.StateView.__new__(Game.StateView)
    sv._batched = True
    with pytest.raises(ValueError):
        wire.dumps(sv)
