"""Same names as the reference package (pokerl/agents/__init__.py:1-3), minus the interactive
CLIAgent, which is out of scope."""
from .agent import PokerAgent
from .random import RandomAgent

__all__ = ["PokerAgent", "RandomAgent"]
