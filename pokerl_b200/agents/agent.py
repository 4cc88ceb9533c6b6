"""`PokerAgent`: the reference's agent interface (pokerl/agents/agent.py:3-14)."""


class PokerAgent:
    """Minimum requirement: `__call__(state)` predicts the next action -- an int for a scalar
    `StateView`, an int tensor [N] for a batched one."""

    def __call__(self, state):
        raise NotImplementedError
