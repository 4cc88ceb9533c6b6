"""Wire-format fixture for Game.StateView (reference game.py:208-240): a pickle written by the LIVE reference.

Build container only (needs /root/reference).  Plays a short deterministic 4-player game (decks injected through
make_golden.DeckShim, so the cards are the product's Philox deal of seed 21 / env 0) and pickles `game.active_state`
at three points -- pre-flop, after the flop and on a later hand -- exactly as pokerl/network/server.py would send it.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_stateview_pickle.py   # -> stateview_ref.pkl, stateview_ref.json
"""
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

SEED, P = 21, 4
ACTIONS = [2, 2, 2, 1, 1, 1, 2, 2, 2, 2]  # calls / checks: reaches the flop, later streets


def describe(sv):
    return dict(player=int(sv.player), valid_actions=[float(x) for x in sv.valid_actions], num_players=int(sv.num_players),
                turn=int(sv.turn), player_cards=[int(c.value) for c in sv.player_cards],
                community_cards=[int(c.value) for c in sv.community_cards],
                credits=[float(x) for x in sv.credits], bets=[float(x) for x in sv.bets],
                pending_bets=[float(x) for x in sv.pending_bets], minimum_raise_value=float(sv.minimum_raise_value),
                pot=float(sv.pot), high_bet=float(sv.high_bet), credit=float(sv.credit))


if __name__ == "__main__":
    shim = mg.DeckShim(SEED, P)
    mg.ref_game.random = shim
    mg.ref_game.np = mg.StableArgsortNumpy(np)
    g = mg.pokerl.Game(num_players=P, start_credits=100, big_blind=2, small_blind=1)
    shim.game = g
    shim.episode += 1
    g.reset()
    blobs, meta = [], []
    for t, a in enumerate(ACTIONS):
        if t in (0, 4, 9):
            sv = g.active_state
            blobs.append(pickle.dumps(sv, protocol=4))
            meta.append(dict(step=t, **describe(sv)))
        shim.in_step = 0
        vm = mg.valid_mask_of(g)
        g.step(a if vm >> a & 1 else 0)
    pickle.dump(blobs, open(os.path.join(HERE, "stateview_ref.pkl"), "wb"), protocol=4)
    json.dump(dict(seed=SEED, num_players=P, actions=ACTIONS, views=meta), open(os.path.join(HERE, "stateview_ref.json"), "w"), indent=1)
    print("wrote", len(blobs), "reference StateView pickles:", [len(b) for b in blobs], "bytes")
