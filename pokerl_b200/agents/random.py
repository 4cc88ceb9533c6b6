"""`RandomAgent`: uniform choice among the valid actions (pokerl/agents/random.py:4-17).

The reference draws from numpy's global generator (`np.random.choice(len(vu), p=vu / np.sum(vu))`).
Here the draw is made on the device: `torch.multinomial` over the valid-action matrix when the agent is
called like any other callable, and -- when a `PokerGameEnv` sees that all its opponents are RandomAgents
-- the counter-based sampler of the step kernel (`pkb_sample_actions`, Philox keyed by (seed, env id,
round)), which needs no Python per table.  `weights` (7 non-negative integers) turns the uniform choice
into a weighted one over the valid moves (the side-pot-heavy mix of BASELINE.json configs[3]).
"""
import numpy as np
import torch

from .agent import PokerAgent


class RandomAgent(PokerAgent):
    def __init__(self, weights=None):
        self.weights = None if weights is None else tuple(int(w) for w in weights)

    def __call__(self, state):
        vu = state.valid_actions
        if isinstance(vu, torch.Tensor):  # batched view: one draw per table
            p = vu if self.weights is None else vu * torch.tensor(self.weights, dtype=vu.dtype, device=vu.device)
            return torch.multinomial(p, 1)[:, 0].to(torch.int32)
        p = np.asarray(vu, dtype=np.float64) * (1.0 if self.weights is None else np.asarray(self.weights, dtype=np.float64))
        return int(np.random.choice(len(p), p=p / np.sum(p)))
