"""`Game` / `Game.StateView`: the reference's API (pokerl/game.py:9-700) over N lockstep tables on a B200.

Two modes, selected by the constructor:

* scalar mode (`num_envs` not given): behaves like one reference `Game`.  `step(int)` returns three
  Python bools, attributes are numpy arrays / ints / lists of `Card`, `active_state` is a
  `StateView` with a float64[7] `valid_actions`.  Invalid action -> `ValueError` with the
  reference's message and the state untouched (game.py:649-651); a non-int action ->
  `NotImplementedError` (game.py:646, 700).  Every call round-trips to the GPU: this mode exists
  for drop-in compatibility and parity tests, not for speed.
* batched mode (`num_envs=N`): `step(actions)` takes an int tensor [N] and returns three bool
  tensors [N]; attributes are device tensors with the env axis first ([N] or [N, P]);
  `active_state` is a batched `StateView` of tensors.

Decks are not stored: hand `hand` of an env is Philox(seed, env_id, episode, hand) (see
csrc/pkb_philox.cuh); `deck` / `community_cards` / `get_cards_of` regenerate it on the device.
"""
import logging
from typing import Generator, List, Tuple

import numpy as np
import torch

from . import _native as nat
from .batch import GameBatch
from .cards import Card
from .enums import PlayerState, PokerMoves


def np_sum_rows(t: torch.Tensor) -> torch.Tensor:
    """Row sums of a [N, P] float64 tensor in numpy's add.reduce order (what `np.sum(self.bets)` does, game.py:283):
    fewer than eight terms left to right; from eight on, eight running lanes combined pairwise, then the tail."""
    P = t.shape[1]
    if P < 8:
        r = torch.zeros_like(t[:, 0])
        for p in range(P):
            r = r + t[:, p]
        return r
    lanes = [t[:, j].clone() for j in range(8)]
    full = P - P % 8
    for k in range(8, full, 8):
        for j in range(8):
            lanes[j] = lanes[j] + t[:, k + j]
    r = ((lanes[0] + lanes[1]) + (lanes[2] + lanes[3])) + ((lanes[4] + lanes[5]) + (lanes[6] + lanes[7]))
    for p in range(full, P):
        r = r + t[:, p]
    return r


class Game:
    """A Texas Hold'em game (scalar mode) or N of them in lockstep (batched mode)."""

    class StateView:
        """The game as seen by one player (game.py:39-240).

        Scalar mode: same field types as the reference (but `credits`/`bets`/`pending_bets` are host
        copies, not aliases of device memory).  Batched mode: tensors with the env axis first;
        `player_cards` is uint8 [N, 2] and `community_cards` uint8 [N, 5] card values with
        `num_community` [N] telling how many are visible; the fields are computed lazily, on first access.
        """

        def __init__(self, game: 'Game', player=None):
            assert game is not None and isinstance(game, Game), 'Invalid game'
            self._batched = game.batched
            self.num_players = game.num_players
            if not game.batched:
                self.player = player or game.active_player  # game.py:122 (seat 0 is falsy)
                self.valid_actions, _ = game.get_valid_actions(self.player)
                self.turn = game.turn
                self.player_cards = game.get_cards_of(self.player)
                self.community_cards = game.community_cards
                self.credits = game.credits
                self.bets = game.bets
                self.pending_bets = game.pending_bets
                self.minimum_raise_value = game.minimum_raise_value
            else:
                # Batched views are LAZY: every field is a kernel launch or a pass over N envs, and a training loop that
                # steps millions of tables looks at one or two of them.  A field is computed from the game's state the
                # first time it is read and kept from then on (read the fields you need before stepping the game again).
                self._game, self._fixed_player = game, player

        _LAZY = ("player", "valid_mask", "valid_actions", "turn", "player_cards", "community_cards", "num_community",
                 "credits", "bets", "pending_bets", "minimum_raise_value")

        def __getattr__(self, name):
            # only reached for attributes that are not set yet: the lazy fields of a batched view
            if name.startswith("_") or name not in Game.StateView._LAZY or not self.__dict__.get("_batched", False):
                raise AttributeError(name)
            self._materialise(name)
            return self.__dict__[name]

        def _materialise(self, name):
            b, d = self._game.batch, self.__dict__
            if name == "player":
                act = b.active_player
                d["player"] = act if self._fixed_player is None else torch.full_like(act, int(self._fixed_player))
            elif name in ("valid_mask", "valid_actions"):
                mask = b.valid_actions(-1 if self._fixed_player is None else int(self._fixed_player))
                d["valid_mask"] = mask
                d["valid_actions"] = ((mask.unsqueeze(1) >> torch.arange(7, device=mask.device)) & 1).to(torch.float64)
            elif name in ("turn", "num_community"):
                turn = b.turn
                d["turn"] = turn
                d["num_community"] = torch.where(turn == 0, torch.zeros_like(turn), turn + 2)
            elif name in ("player_cards", "community_cards"):
                cards = b.cards()  # [5+2P, N]
                idx = 5 + 2 * self.player.long()
                d["player_cards"] = torch.stack([cards.gather(0, idx.unsqueeze(0))[0],
                                                 cards.gather(0, (idx + 1).unsqueeze(0))[0]], dim=1)
                d["community_cards"] = cards[:5].t().contiguous()
            elif name == "credits":
                d["credits"] = b.credits.t()
            elif name == "bets":
                d["bets"] = b.bets.t()
            elif name == "pending_bets":
                d["pending_bets"] = b.pending_bets.t()
            elif name == "minimum_raise_value":
                d["minimum_raise_value"] = b.minimum_raise_value

        @property
        def valid_action_indices(self):
            """Indices of the valid actions: a generator (scalar) or a list of index tensors (batched)."""
            if self._batched:
                return [torch.nonzero(row, as_tuple=False).flatten() for row in self.valid_actions]
            return (action for action, valid in enumerate(self.valid_actions) if valid)

        @property
        def player_hand(self):
            """Player cards plus the visible community cards (game.py:139-143)."""
            if self._batched:
                return torch.cat([self.player_cards, self.community_cards], dim=1)
            return self.player_cards + self.community_cards

        @property
        def pot(self):
            return np_sum_rows(self.bets) if self._batched else np.sum(self.bets)

        @property
        def high_bet(self):
            return self.pending_bets.max(dim=1).values if self._batched else np.max(self.pending_bets)

        @property
        def credit(self):
            if self._batched:
                return self.credits.gather(1, self.player.long().unsqueeze(1))[:, 0]
            return self.credits[self.player]

        def __repr__(self):
            return repr(dict(player=self.player, valid_actions=self.valid_actions, num_players=self.num_players,
                             turn=self.turn, player_cards=self.player_cards, community_cards=self.community_cards,
                             credits=self.credits, bets=self.bets, minimum_raise_value=self.minimum_raise_value))

        def __str__(self) -> str:
            if self._batched:
                return f'<batched StateView of {self.credits.shape[0]} envs>'
            valid = ', '.join(f'{a}: {PokerMoves.as_string[a]}' for a in self.valid_action_indices)
            return ('==== State =================\n'
                    f'# Player ID_________{self.player:d}\n'
                    f'# Game turn_________{self.turn:d}\n'
                    f'# Player credit_____${self.credit:.2f}\n'
                    f"# Player cards______{', '.join(str(c) for c in self.player_cards)}\n"
                    f"# Community cards___{', '.join(str(c) for c in self.community_cards) or 'n/a'}\n"
                    f'# Pot_______________${self.pot:.2f}\n'
                    f'# High bet__________${self.high_bet:.2f}\n'
                    f'# Minimum raise_____${self.minimum_raise_value:.2f}\n'
                    f'# Valid actions_____{valid}\n')

        # pickled as the reference's 10-tuple (game.py:208-240) so a reference client can load it
        def __getstate__(self) -> tuple:
            return (self.player, self.valid_actions, self.num_players, self.turn, self.player_cards,
                    self.community_cards, self.credits, self.bets, self.pending_bets, self.minimum_raise_value)

        def __setstate__(self, state: tuple):
            (self.player, self.valid_actions, self.num_players, self.turn, self.player_cards, self.community_cards,
             self.credits, self.bets, self.pending_bets, self.minimum_raise_value) = state
            self._batched = isinstance(self.credits, torch.Tensor)

    # ---------------------------------------------------------------------------------------------

    def __init__(self, **config):
        self.num_players: int = config.get('num_players', 4)
        self.start_credits = config.get('start_credits', 100)
        self.big_blind: float = config.get('big_blind', 2)
        self.small_blind: float = config.get('small_blind', 1)
        self.logger: logging.Logger = config.get('logger', logging.getLogger())
        num_envs = config.get('num_envs', None)
        self.batched = num_envs is not None
        self.num_envs = int(num_envs) if self.batched else 1
        self.check_actions = config.get('check_actions', True)
        self.batch = GameBatch(
            self.num_envs, num_players=self.num_players, start_credits=self._credits_arg(self.start_credits),
            big_blind=self.big_blind, small_blind=self.small_blind, dealer=0, seed=config.get('seed', 0),
            env_id_base=config.get('env_id_base', 0), auto_reset=config.get('auto_reset', False),
            device=config.get('device', None))
        self.device = self.batch.device
        self._step_count = 0
        # like the reference before reset(): all ACTIVE, zero money (game.py:253-264)
        self.batch.seat_masks.fill_((1 << self.num_players) - 1)  # the ACTIVE field of the seat-mask word
        self.batch.refresh_valid()  # get_valid_actions of the all-zero table (game.py:339-383), for the device sampler

    @staticmethod
    def _credits_arg(sc):
        # game.py:408: int -> broadcast; list / ndarray -> copied; float -> TypeError; np.int64 -> IndexError
        if isinstance(sc, np.integer):
            raise IndexError('invalid index to scalar variable.')
        return sc

    # ---- state attributes ----------------------------------------------------------------------
    def _vec(self, t):  # [P, N] plane -> numpy (P,) or tensor [N, P]
        return t.t() if self.batched else t[:, 0].cpu().numpy()

    def _int(self, t):
        return t if self.batched else int(t[0].item())

    credits = property(lambda self: self._vec(self.batch.credits))
    bets = property(lambda self: self._vec(self.batch.bets))
    pending_bets = property(lambda self: self._vec(self.batch.pending_bets))
    payoffs = property(lambda self: self._vec(self.batch.payoffs))
    player_states = property(lambda self: self._vec(self.batch.player_states))
    turn = property(lambda self: self._int(self.batch.turn))
    hand = property(lambda self: self._int(self.batch.hand))
    dealer_idx = property(lambda self: self._int(self.batch.dealer_idx))
    small_blind_idx = property(lambda self: self._int(self.batch.small_blind_idx))
    big_blind_idx = property(lambda self: self._int(self.batch.big_blind_idx))
    active_player = property(lambda self: self._int(self.batch.active_player))

    @property
    def minimum_raise_value(self):
        t = self.batch.minimum_raise_value
        return t if self.batched else float(t[0].item())

    @property
    def deck(self):
        """The dealt part of the deck: board (5) then hole cards seat by seat (game.py:278, 385-395)."""
        cards = self.batch.cards()
        if self.batched:
            return cards.t()
        return [Card(int(v)) for v in cards[:, 0].tolist()]

    @property
    def community_cards(self):
        """Zero, three, four or five shared cards depending on the street (game.py:266-278)."""
        if self.batched:
            return self.batch.cards()[:5].t()
        return [] if self.turn == 0 else self.deck[:self.turn + 2]

    @property
    def pot(self):
        return np_sum_rows(self.bets) if self.batched else np.sum(self.bets)

    @property
    def high_bet(self):
        return self.pending_bets.max(dim=1).values if self.batched else np.max(self.pending_bets)

    @property
    def high_bidders(self):
        return self.bets == (self.high_bet.unsqueeze(1) if self.batched else self.high_bet)

    @property
    def pending_credits(self):
        return self.credits - self.pending_bets

    @property
    def blind_idx(self):
        return [self.big_blind_idx, self.small_blind_idx]

    @property
    def blind_value(self):
        return [self.big_blind, self.small_blind]

    @property
    def game_over(self):
        """True when all players but one are broken (game.py:316-320)."""
        over = self.batch.game_over()
        return over.bool() if self.batched else bool(over[0].item())

    @property
    def active_state(self) -> 'Game.StateView':
        return self.StateView(self, None if self.batched else self.active_player)

    # ---- helpers with the reference's names ----------------------------------------------------
    def get_first_playing(self, idx: int):
        """First non-broken seat at or after `idx`, cyclically (game.py:334-337); scalar mode."""
        st = self.player_states
        P = self.num_players
        for k in range(P):
            if st[(idx + k) % P] != PlayerState.BROKEN:
                return (idx + k) % P
        return idx % P

    def get_valid_actions(self, player: int = None):
        """One-hot float64 array of valid actions and a generator of their indices (game.py:339-383)."""
        mask = self.batch.valid_actions(-1 if player is None else int(player))
        if self.batched:
            onehot = ((mask.unsqueeze(1) >> torch.arange(7, device=mask.device)) & 1).to(torch.float64)
            return onehot, mask
        m = int(mask[0].item())
        onehot = np.array([float(m >> a & 1) for a in range(PokerMoves.NUM_MOVES)])
        return onehot, (action for action, valid in enumerate(onehot) if valid)

    def get_cards_of(self, player: int):
        """The two hole cards of `player` (game.py:385-389)."""
        i = 5 + player * 2
        return self.deck[:, i:i + 2] if self.batched else self.deck[i:i + 2]

    def get_hand_for(self, player: int):
        """Board plus hole cards, board first (game.py:391-395)."""
        i = 5 + player * 2
        d = self.deck
        return torch.cat([d[:, :5], d[:, i:i + 2]], dim=1) if self.batched else d[:5] + d[i:i + 2]

    # ---- transitions ---------------------------------------------------------------------------
    def reset(self, env_mask=None, **config):
        """Reset to the initial state and deal the first hand (game.py:397-412).

        Only `dealer` is read from the keyword arguments, like the reference (game.py:403).
        Batched mode accepts a bool/uint8 mask [N] of the envs to reset.
        """
        dealer = int(config.get('dealer', 0))
        if dealer != self.batch.config.dealer:
            self.batch.set_reset_dealer(dealer)
        self.batch.reset(env_mask)

    def step(self, action, check: bool = None):
        """Make the active player of every env perform its action (game.py:621-700).

        Returns (game_over, hand_over, round_over): Python bools in scalar mode, bool tensors [N] in
        batched mode.
        """
        check = self.check_actions if check is None else check
        if not self.batched:
            if not isinstance(action, (int, np.integer)):
                raise NotImplementedError  # game.py:646, 700
            act = torch.tensor([int(action)], dtype=torch.int32, device=self.device)
            flags, _ = self.batch.step(act)
            f = int(flags[0].item())
            if f & nat.F_INVALID:
                # the reference formats as_string[action] first: IndexError for action >= 7 (game.py:650)
                msg = 'Player %d invalid move: `%s`' % (self.active_player, PokerMoves.as_string[action])
                self.logger.error(msg)
                raise ValueError(msg)
            if f & nat.F_POISONED:
                raise RuntimeError('Game.step: the reference would not terminate or would assert here '
                                   '(game.py:473 / game.py:607); env flagged PKB_F_POISONED')
            self._step_count += 1
            return bool(f & nat.F_GAME_OVER), bool(f & nat.F_HAND_OVER), bool(f & nat.F_ROUND_OVER)
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(np.asarray(action), dtype=torch.int32)
        flags, _ = self.batch.step(action)
        if check:
            bad = flags & (nat.F_INVALID | nat.F_POISONED)
            if bool(bad.any().item()):
                env = int(torch.nonzero(bad)[0].item())
                kind = 'invalid move' if int(flags[env].item()) & nat.F_INVALID else 'poisoned (non-terminating) state'
                raise ValueError('Env %d player %d %s: `%s` (%d envs rejected; their state is unchanged)' % (
                    env, int(self.batch.active_player[env].item()), kind, int(action[env].item()),
                    int((bad != 0).sum().item())))
        self._step_count += 1
        return (flags & nat.F_GAME_OVER) != 0, (flags & nat.F_HAND_OVER) != 0, (flags & nat.F_ROUND_OVER) != 0

    def step_random(self, weights=None, record_actions=False):
        """Batched step with the on-device RandomAgent (agents/random.py:16-17); returns the three flag tensors."""
        flags, _ = self.batch.step_random(self._step_count, weights, record_actions)
        self._step_count += 1
        if self.batched:
            return (flags & nat.F_GAME_OVER) != 0, (flags & nat.F_HAND_OVER) != 0, (flags & nat.F_ROUND_OVER) != 0
        f = int(flags[0].item())
        return bool(f & nat.F_GAME_OVER), bool(f & nat.F_HAND_OVER), bool(f & nat.F_ROUND_OVER)
