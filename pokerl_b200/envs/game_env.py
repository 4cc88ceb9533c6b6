"""`PokerGameEnv`: the reference's RL environment (pokerl/envs/game_env.py:6-54) over the B200 `Game`.

The agent sits in seat 0; the other seats are played by callables `StateView -> action`.

* scalar mode (no `num_envs`): the reference's loops verbatim; `step` returns
  `(StateView, reward: float, done: bool, hand_over: bool)` -- the 4-tuple the reference actually
  returns (game_env.py:39, 54), not the 3-tuple of its annotation.
* batched mode (`num_envs=N`): N tables advance in lockstep.  The reference's per-table loops
  "step the opponents until ..." (game_env.py:24-27, 41-44, 49-52) become masked loops: in every
  round each opponent callable is asked for a whole batch of actions, the action of the seat that is
  actually active is selected per table, tables that must not move get `PKB_ACTION_SKIP`, and ONE
  `pkb_step` launch advances all of them.  `step` takes an int tensor [N] and returns
  `(StateView, reward f64[N], done bool[N], hand_over bool[N])`.
  Opponents may be `pokerl_b200.agents.RandomAgent` instances: those are sampled on the device by
  `pkb_sample_actions` (no Python per table, no host round trip).  Tables whose episode ended are
  re-armed with `reset(env_mask=done)`.
  When EVERY opponent is a RandomAgent (with the same weights) the masked loops are not needed at all:
  `step` and `reset` are ONE launch each (`pkb_env_step` / `pkb_env_reset`), the kernel plays the
  opponents round by round inside the tile until the reference's loops would return.  No host
  synchronisation, no per-round torch ops.
"""
from typing import List, Tuple

import torch

from .. import _native as nat
from ..agents.random import RandomAgent
from ..enums import PlayerState
from .env import Game, PokerEnv


class PokerGameEnv(PokerEnv):
    """Simulates entire poker games; an episode is considered done after each hand (game_env.py:6-11)."""

    def __init__(self, agents: list = None, **game_config):
        self.game = Game(**game_config)
        self.agents = [None, *(agents or [])]  # game_env.py:17
        self.player_agent = 0
        if len(self.agents) != self.game.num_players:
            raise ValueError(f"need {self.game.num_players - 1} opponent agents, got {len(self.agents) - 1}")
        self._round = 0  # counter of batched opponent rounds: the step index of the device-side sampler
        opponents = self.agents[1:]
        # every opponent the device-side random agent with one weight vector: one fused launch per step / reset
        self._fused = (self.game.batched and len(opponents) > 0 and all(isinstance(ag, RandomAgent) for ag in opponents)
                       and len({ag.weights for ag in opponents}) == 1 and game_config.get('fused', True))

    # ---- scalar mode: the reference's control flow, line by line -------------------------------
    def _reset_scalar(self) -> Game.StateView:
        self.game.reset()
        while self.game.active_player != self.player_agent:  # game_env.py:24-27
            action = self.agents[self.game.active_player](self.game.active_state)
            done, *_ = self.game.step(int(action))
            if done:
                self.game.reset()
        return self.game.active_state

    def _step_scalar(self, action: int):
        reward = .0
        done, hand, _ = self.game.step(action)
        if done or self.game.player_states[self.player_agent] == PlayerState.BROKEN:  # game_env.py:37-39
            return self.game.active_state, self.game.payoffs[self.player_agent], True, True
        while not hand and self.game.active_player != self.player_agent:  # game_env.py:41-44
            action = self.agents[self.game.active_player](self.game.active_state)
            done, hand, _ = self.game.step(int(action))
        if hand:  # game_env.py:47
            reward = self.game.payoffs[self.player_agent]
        while not done and self.game.active_player != self.player_agent:  # game_env.py:49-52
            action = self.agents[self.game.active_player](self.game.active_state)
            done, *_ = self.game.step(int(action))
        return self.game.active_state, reward, done, hand

    # ---- batched mode --------------------------------------------------------------------------
    def _opponent_actions(self, mask: torch.Tensor) -> torch.Tensor:
        """Actions of the active seats for the tables in `mask`; PKB_ACTION_SKIP elsewhere."""
        b = self.game.batch
        opponents = self.agents[1:]
        if all(isinstance(ag, RandomAgent) for ag in opponents) and len({ag.weights for ag in opponents}) == 1:
            actions = b.sample_actions(self._round, opponents[0].weights).clone()
        else:
            state = self.game.active_state  # one batched view; every callable sees all tables
            active = b.active_player
            actions = torch.zeros(self.game.num_envs, dtype=torch.int32, device=self.game.device)
            for seat in range(1, self.game.num_players):
                ag = self.agents[seat]
                if isinstance(ag, RandomAgent):
                    a = b.sample_actions(self._round, ag.weights, out=torch.empty_like(actions))
                else:
                    a = torch.as_tensor(ag(state), device=self.game.device).to(torch.int32)
                actions = torch.where(active == seat, a, actions)
        self._round += 1
        return torch.where(mask, actions, torch.full_like(actions, nat.ACTION_SKIP))

    def _step_masked(self, actions: torch.Tensor):
        """One launch; returns (game_over, hand_over) bool tensors, False where the table was skipped.

        A table that reaches a state in which the reference would spin forever (game.py:607) or assert (game.py:473) is
        flagged PKB_F_POISONED by the kernel; here it counts as a finished game, so the caller's reset re-arms it."""
        done, hand, _ = self.game.step(actions, check=False)
        flags = self.game.batch.flags
        if self.game.check_actions and bool(((flags & nat.F_INVALID) != 0).any().item()):
            env = int(torch.nonzero(flags & nat.F_INVALID)[0].item())
            raise ValueError('Env %d player %d invalid move: `%d`' % (env, int(self.game.batch.active_player[env].item()),
                                                                      int(actions[env].item())))
        dead = (flags & nat.F_POISONED) != 0
        return done | dead, hand | dead

    def _fused_round_base(self):
        base = self._round
        self._round += nat.ENV_ROUND_CAP  # every call owns its own range of sampler step indices
        return base

    def _reset_fused(self, env_mask=None):
        self.game.batch.env_reset(env_mask, self._fused_round_base(), self.agents[1].weights)
        return self.game.active_state

    def _step_fused(self, action):
        g, b = self.game, self.game.batch
        action = torch.as_tensor(action, device=g.device).to(torch.int32)
        reward, done, _ = b.env_step(action, self._fused_round_base(), self.agents[1].weights)
        if g.check_actions and bool(((done & nat.F_INVALID) != 0).any().item()):
            env = int(torch.nonzero(done & nat.F_INVALID)[0].item())
            raise ValueError('Env %d player 0 invalid move: `%d` (its state is unchanged)' % (env, int(action[env].item())))
        return g.active_state, reward.clone(), (done & 1) != 0, (done & 2) != 0

    def _reset_batched(self, env_mask=None) -> Game.StateView:
        if self._fused:
            return self._reset_fused(env_mask)
        b = self.game.batch
        if env_mask is not None:
            env_mask = torch.as_tensor(env_mask, device=self.game.device).bool()
        self.game.reset(env_mask)
        todo = torch.ones(self.game.num_envs, dtype=torch.bool, device=self.game.device) if env_mask is None else env_mask
        while True:  # game_env.py:24-27 on every table still waiting for seat 0
            mask = todo & (b.active_player != self.player_agent)
            if not bool(mask.any().item()):
                break
            done, _ = self._step_masked(self._opponent_actions(mask))
            over = mask & done
            if bool(over.any().item()):
                self.game.reset(over)
        return self.game.active_state

    def _step_batched(self, action):
        if self._fused:
            return self._step_fused(action)
        g, b = self.game, self.game.batch
        action = torch.as_tensor(action, device=g.device).to(torch.int32)
        done, hand = self._step_masked(action)
        done, hand = done.clone(), hand.clone()
        seat0 = self.player_agent
        early = done | (b.player_states[seat0] == PlayerState.BROKEN)  # game_env.py:37-39
        reward = torch.zeros(g.num_envs, dtype=torch.float64, device=g.device)
        reward = torch.where(early, b.payoffs[seat0], reward)
        while True:  # game_env.py:41-44
            mask = ~early & ~hand & (b.active_player != seat0)
            if not bool(mask.any().item()):
                break
            d, h = self._step_masked(self._opponent_actions(mask))
            done = torch.where(mask, d, done)
            hand = torch.where(mask, h, hand)
        reward = torch.where(~early & hand, b.payoffs[seat0], reward)  # game_env.py:47
        while True:  # game_env.py:49-52
            mask = ~early & ~done & (b.active_player != seat0)
            if not bool(mask.any().item()):
                break
            d, _ = self._step_masked(self._opponent_actions(mask))
            done = torch.where(mask, d, done)
        return g.active_state, reward, done | early, hand | early

    # ---- public API (game_env.py:20, 31) -------------------------------------------------------
    def reset(self, env_mask=None) -> Game.StateView:
        """Start new games (all tables, or those selected by `env_mask` in batched mode) and play the
        opponents until it is the agent's turn."""
        if self.game.batched:
            return self._reset_batched(env_mask)
        return self._reset_scalar()

    def step(self, action) -> Tuple[Game.StateView, float, bool, bool]:
        if self.game.batched:
            return self._step_batched(action)
        return self._step_scalar(action)
