"""`PokerEnv`: the reference's gym-like base class (pokerl/envs/env.py:5-23)."""
from typing import Tuple

from ..game import Game


class PokerEnv:
    """Base class of the poker environments: `reset()` returns the agent's state, `step(action)`
    returns `(next state, reward, done, hand_over)`."""

    def reset(self) -> Game.StateView:
        raise NotImplementedError

    def step(self, action) -> Tuple[Game.StateView, float, bool, bool]:
        raise NotImplementedError
