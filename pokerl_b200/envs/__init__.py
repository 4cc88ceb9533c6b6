"""Same names as the reference package (pokerl/envs/__init__.py:1-2)."""
from .env import PokerEnv
from .game_env import PokerGameEnv

__all__ = ["PokerEnv", "PokerGameEnv"]
