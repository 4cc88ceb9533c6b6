"""pokerl_b200: a B200-native batched No-Limit Texas Hold'em environment, drop-in for the hot
path of sneppy/pokerl (`Game.step`, dealing, `judger.eval_hand`).

Same top-level names as the reference package (pokerl/__init__.py:1-4).  All game logic runs in
hand-written CUDA kernels (pokerl_b200/csrc, built for sm_100a) behind the C ABI declared in
include/pokerl_b200.h; there is no CPU fallback.
"""
from .game import Game
from .judger import eval_hand
from .cards import Card, create_default_deck
from .enums import PokerMoves, HandRanking
from .batch import GameBatch
from . import agents, envs  # noqa: E402  (pokerl.envs / pokerl.agents)

__version__ = "0.1.0"
__all__ = ["Game", "GameBatch", "eval_hand", "Card", "create_default_deck", "PokerMoves", "HandRanking"]
