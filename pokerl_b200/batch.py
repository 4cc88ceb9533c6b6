"""GameBatch: N lockstep reference `Game` tables on one B200, driven through the C ABI.

This is the thin host layer between the pokerl-compatible facade (game.py, envs/) and
libpokerl_b200.so.  PyTorch is used only for device buffers, streams and (multi-GPU)
torch.distributed; all game logic runs in the hand-written CUDA kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat


class _DeviceArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, address, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(address), False),
                                         "version": 2, "strides": None}
        self._owner = owner  # keeps the pkb handle alive as long as the tensor lives


def _credits_vector(num_players, start_credits):
    """start_credits as the reference accepts it (game.py:247, 408): int, list or ndarray."""
    if isinstance(start_credits, (int, np.integer)) and not isinstance(start_credits, bool):
        return [float(start_credits)] * num_players
    if isinstance(start_credits, float):
        # the reference does `100.0[:]` here -> TypeError
        raise TypeError("'float' object is not subscriptable")
    vec = [float(x) for x in start_credits]
    if len(vec) != num_players:
        raise ValueError(f"could not broadcast start_credits of length {len(vec)} to {num_players} players")
    return vec


class _Handle:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if self.ptr:
            try:
                nat.lib().pkb_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = None


class GameBatch:
    """N environments on one device.  All tensors returned are views of, or live on, that device."""

    def __init__(self, num_envs, num_players=4, start_credits=100, big_blind=2, small_blind=1, dealer=0, seed=0,
                 env_id_base=0, auto_reset=False, device=None):
        nat.require_cuda()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.num_envs = int(num_envs)
        self.num_players = int(num_players)
        if not 2 <= self.num_players <= nat.MAX_PLAYERS:
            raise ValueError(f"num_players must be in 2..{nat.MAX_PLAYERS} (got {num_players})")
        cfg = nat.PkbConfig()
        cfg.abi_version = nat.ABI_VERSION
        cfg.device = self.device.index
        cfg.num_envs = self.num_envs
        cfg.env_id_base = int(env_id_base)
        cfg.num_players = self.num_players
        cfg.dealer = int(dealer)
        for p, x in enumerate(_credits_vector(self.num_players, start_credits)):
            cfg.start_credits[p] = x
        cfg.big_blind = float(big_blind)
        cfg.small_blind = float(small_blind)
        cfg.seed = int(seed) & (2 ** 64 - 1)
        cfg.auto_reset = int(bool(auto_reset))
        self.config = cfg
        self.seed, self.env_id_base = cfg.seed, int(env_id_base)
        h = C.c_void_p()
        nat.check(nat.lib().pkb_create(C.byref(cfg), C.byref(h)))
        self._h = _Handle(h)
        self._wrap_state()
        with torch.cuda.device(self.device):
            # one [2, N] buffer (row 0 flags, row 1 valid masks): a caller that wants both on the host copies them at once
            self.outputs = torch.zeros((2, self.num_envs), dtype=torch.uint8, device=self.device)
            self.flags, self.valid = self.outputs[0], self.outputs[1]
            self.actions_taken = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            self._packed = None
            self._env_reward = self._env_done = None

    # ---- zero-copy state ---------------------------------------------------------------------
    def _wrap_state(self):
        sv = nat.PkbStateView()
        nat.check(nat.lib().pkb_state(self._h.ptr, C.byref(sv)))
        P, S, N = self.num_players, int(sv.stride), self.num_envs
        self.stride = S

        def wrap(addr, shape, typestr):
            return torch.as_tensor(_DeviceArray(addr, shape, typestr, self._h), device=self.device)

        # [P, N] views of the seat-major planes (env index fastest in memory)
        self.credits = wrap(sv.credits, (P, S), "<f8")[:, :N]
        self.bets = wrap(sv.bets, (P, S), "<f8")[:, :N]
        self.pending_bets = wrap(sv.pending_bets, (P, S), "<f8")[:, :N]
        # Game.payoffs is stored env-major (one zero-padded record of payoff_pitch doubles per env): a strided [P, N] view
        self.payoffs = wrap(sv.payoffs, (S, int(sv.payoff_pitch)), "<f8")[:N, :P].T
        self.minimum_raise_value = wrap(sv.minimum_raise_value, (S,), "<f8")[:N]
        # player_states live on the device as four 16-bit seat masks per env (ACTIVE | CALLED<<16 | ALL_IN<<32 | BROKEN<<48)
        self.seat_masks = wrap(sv.seat_masks, (S,), "<i8")[:N]
        self._states_buf = None
        self._meta = wrap(sv.meta, (S,), "<i4")[:N]
        self._hand = wrap(sv.hand, (S,), "<i4")[:N]
        self._episode = wrap(sv.episode, (S,), "<i4")[:N]

    @property
    def player_states(self):
        """Game.player_states (game.py:259) for every env: uint8 [P, N], one PlayerState value per seat (a fresh copy)."""
        if self._states_buf is None:
            self._states_buf = torch.zeros((self.num_players, self.stride), dtype=torch.uint8, device=self.device)
        nat.check(nat.lib().pkb_player_states(self._h.ptr, nat.ptr(self._states_buf), self._stream()))
        return self._states_buf[:, :self.num_envs]

    @property
    def turn(self):
        return self._meta & 7

    @property
    def dealer_idx(self):
        return (self._meta >> 3) & 31

    @property
    def small_blind_idx(self):
        return (self._meta >> 8) & 31

    @property
    def big_blind_idx(self):
        return (self._meta >> 13) & 31

    @property
    def active_player(self):
        return (self._meta >> 18) & 31

    @property
    def hand(self):
        return self._hand

    @property
    def episode(self):
        return self._episode

    # ---- calls -------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_reset_dealer(self, dealer: int):
        """The `dealer` keyword of Game.reset (game.py:403) for every later reset."""
        nat.check(nat.lib().pkb_set_reset_dealer(self._h.ptr, int(dealer)))
        self.config.dealer = int(dealer)

    def refresh_valid(self):
        """Recompute the cached valid-move masks after writing the state planes (credits, pending_bets, ...) directly."""
        nat.check(nat.lib().pkb_refresh_valid(self._h.ptr, self._stream()))

    def set_pipelining(self, enabled: bool):
        """Whether consecutive step launches may overlap (default on; results are identical either way)."""
        nat.check(nat.lib().pkb_set_pipelining(self._h.ptr, int(bool(enabled))))

    def set_env_phases(self, min_envs: int):
        """PokerGameEnv.step in phases (stragglers packed into a scratch arena) for handles of at least min_envs envs
        (default 16384; 0: always; negative: never).  Results are identical either way."""
        nat.check(nat.lib().pkb_set_env_phases(self._h.ptr, int(min_envs)))

    def reset(self, env_mask=None):
        """Game.reset() on every env, or on those where env_mask (uint8/bool [N]) is nonzero."""
        if env_mask is not None:
            env_mask = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        nat.check(nat.lib().pkb_reset(self._h.ptr, nat.ptr(env_mask), self._stream()))

    def _out(self, t, default, name):
        """A caller-supplied output buffer goes to the kernel by raw pointer: it has to be what the kernel writes."""
        if t is None:
            return default
        if not (isinstance(t, torch.Tensor) and t.dtype == torch.uint8 and t.device == self.device and t.is_contiguous()
                and t.numel() >= self.num_envs):
            raise ValueError(f"{name} must be a contiguous uint8 tensor of at least {self.num_envs} elements on {self.device}")
        return t

    def step(self, actions, flags=None, valid=None):
        """Game.step(actions[i]) on env i.  actions: int32 [N], or uint8 [N] (one byte per move index, 255 = skip: a
        quarter of the bytes when the actions come from the host); tensors that live elsewhere or are not contiguous
        are brought to this device in their own encoding first.  Returns (flags, valid)."""
        flags, valid = self._out(flags, self.flags, "flags"), self._out(valid, self.valid, "valid")
        actions = torch.as_tensor(actions)
        if actions.numel() < self.num_envs:
            raise ValueError(f"need {self.num_envs} actions, got {actions.numel()}")
        if actions.dtype == torch.uint8:  # the byte encoding keeps its meaning (255 = skip) wherever the tensor lives
            if actions.device != self.device or not actions.is_contiguous():
                actions = actions.to(device=self.device).contiguous()
            nat.check(nat.lib().pkb_step_u8(self._h.ptr, nat.ptr(actions), nat.ptr(flags), nat.ptr(valid), self._stream()))
            return flags, valid
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        nat.check(nat.lib().pkb_step(self._h.ptr, nat.ptr(actions), nat.ptr(flags), nat.ptr(valid), self._stream()))
        return flags, valid

    def step_packed(self, actions_u4, out=None):
        """Game.step in the compact encoding of pkb_step_packed: `actions_u4` uint8 [(N+1)//2] on this device, two moves per
        byte (see pack_actions); returns uint8 [N]: flags bits 0-2 | valid bits of moves 1..5 << 3 (see unpack_outputs)."""
        if out is None:
            if self._packed is None:
                self._packed = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            out = self._packed
        out = self._out(out, None, "out")
        if not (isinstance(actions_u4, torch.Tensor) and actions_u4.dtype == torch.uint8 and actions_u4.device == self.device
                and actions_u4.is_contiguous() and actions_u4.numel() >= (self.num_envs + 1) // 2):
            raise ValueError(f"actions_u4 must be a contiguous uint8 tensor of {(self.num_envs + 1) // 2} bytes on {self.device}")
        nat.check(nat.lib().pkb_step_packed(self._h.ptr, nat.ptr(actions_u4), nat.ptr(out), self._stream()))
        return out

    @staticmethod
    def pack_actions(actions):
        """int moves [N] (-2**31 or 15 = skip) -> uint8 [(N+1)//2], two per byte, env 2i in the low nibble."""
        a = torch.as_tensor(actions)
        a = torch.where((a < 0) | (a > 14), torch.full_like(a, 15), a).to(torch.uint8)
        if a.numel() % 2:
            a = torch.cat([a, torch.full((1,), 15, dtype=torch.uint8, device=a.device)])
        return (a[0::2] | (a[1::2] << 4)).contiguous()

    @staticmethod
    def unpack_outputs(packed):
        """packed uint8 [N] -> (flags uint8 [N] with bits 0-2, valid uint8 [N] seven-bit masks)."""
        return packed & 7, ((packed >> 3) << 1) | 0x41

    def env_step(self, actions, round_base, weights=None, reward=None, done=None, valid=None):
        """PokerGameEnv.step (envs/game_env.py:31-54) for every env in ONE launch: seat 0 plays actions[i] (int32 [N]), the
        other seats are played by the device-side random agent until the reference's loops would return.
        Returns (reward float64 [N], done uint8 [N]: bit 0 done, bit 1 hand_over, flag bits as in step flags, valid)."""
        if self._env_reward is None:
            self._env_reward = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
            self._env_done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        reward = self._env_reward if reward is None else reward
        done = self._out(done, self._env_done, "done")
        valid = self._out(valid, self.valid, "valid")
        actions = torch.as_tensor(actions)
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if actions.numel() < self.num_envs:
            raise ValueError(f"need {self.num_envs} actions, got {actions.numel()}")
        w = None if weights is None else (C.c_uint32 * nat.NUM_MOVES)(*[int(x) for x in weights])
        nat.check(nat.lib().pkb_env_step(self._h.ptr, nat.ptr(actions), int(round_base) & 0xFFFFFFFF, w, nat.ptr(reward),
                                         nat.ptr(done), nat.ptr(valid), self._stream()))
        return reward, done, valid

    def env_reset(self, env_mask=None, round_base=0, weights=None, valid=None):
        """PokerGameEnv.reset (envs/game_env.py:20-29) for the envs in env_mask (uint8 / bool [N]; None: all) in ONE launch."""
        if env_mask is not None:
            env_mask = torch.as_tensor(env_mask).to(device=self.device, dtype=torch.uint8).contiguous()
        valid = self._out(valid, self.valid, "valid")
        w = None if weights is None else (C.c_uint32 * nat.NUM_MOVES)(*[int(x) for x in weights])
        nat.check(nat.lib().pkb_env_reset(self._h.ptr, nat.ptr(env_mask), int(round_base) & 0xFFFFFFFF, w, nat.ptr(valid),
                                          self._stream()))
        return valid

    def step_random(self, step_idx, weights=None, record_actions=False, flags=None, valid=None, repeat=1):
        """One step with the on-device random agent (uniform, or integer `weights` per move).  repeat = K: K consecutive
        steps (sampler indices step_idx .. step_idx + K - 1) in one launch -- every tile of envs is stepped K times in
        shared memory between one load and one store; state, statistics and payoffs are those of K single calls, the
        outputs are the last step's (pkb_step_random_n)."""
        w = None if weights is None else (C.c_uint32 * nat.NUM_MOVES)(*[int(x) for x in weights])
        flags, valid = self._out(flags, self.flags, "flags"), self._out(valid, self.valid, "valid")
        out = self.actions_taken if record_actions else None
        if int(repeat) != 1:
            nat.check(nat.lib().pkb_step_random_n(self._h.ptr, int(step_idx) & 0xFFFFFFFF, int(repeat), w, nat.ptr(out),
                                                  nat.ptr(flags), nat.ptr(valid), self._stream()))
        else:
            nat.check(nat.lib().pkb_step_random(self._h.ptr, int(step_idx) & 0xFFFFFFFF, w, nat.ptr(out), nat.ptr(flags),
                                                nat.ptr(valid), self._stream()))
        return flags, valid

    def sample_actions(self, step_idx, weights=None, out=None):
        w = None if weights is None else (C.c_uint32 * nat.NUM_MOVES)(*[int(x) for x in weights])
        out = self.actions_taken if out is None else out
        nat.check(nat.lib().pkb_sample_actions(self._h.ptr, int(step_idx) & 0xFFFFFFFF, w, nat.ptr(out), self._stream()))
        return out

    def valid_actions(self, player=-1, out=None):
        """Bit mask of valid moves (game.py:339-383) for `player` (-1: each env's active player)."""
        out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device) if out is None else out
        nat.check(nat.lib().pkb_valid_actions(self._h.ptr, int(player), nat.ptr(out), self._stream()))
        return out

    def game_over(self, out=None):
        out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device) if out is None else out
        nat.check(nat.lib().pkb_game_over(self._h.ptr, nat.ptr(out), self._stream()))
        return out

    def cards(self):
        """Current deal, uint8 [5+2P, N] card values: rows 0..4 board, 5+2p/6+2p hole cards of seat p."""
        rows = 5 + 2 * self.num_players
        buf = torch.empty((rows, self.stride), dtype=torch.uint8, device=self.device)
        nat.check(nat.lib().pkb_cards(self._h.ptr, nat.ptr(buf), self._stream()))
        return buf[:, :self.num_envs]

    def hand_features(self, player=-1, out=None):
        """packed eval_hand(player_hand) per env (uint32 stored as int64-safe int32 bits)."""
        out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device) if out is None else out
        nat.check(nat.lib().pkb_hand_features(self._h.ptr, int(player), nat.ptr(out), self._stream()))
        return out

    def stats(self) -> dict:
        s = nat.PkbStats()
        nat.check(nat.lib().pkb_stats(self._h.ptr, C.byref(s), self._stream()))
        return dict(steps=s.steps, hands=s.hands, showdowns=s.showdowns, games=s.games, invalid=s.invalid,
                    poisoned=s.poisoned, payoff_sum=[s.payoff_sum[p] for p in range(self.num_players)])

    def stats_reset(self):
        nat.check(nat.lib().pkb_stats_reset(self._h.ptr, self._stream()))

    # ---- device snapshots ----------------------------------------------------------------------
    _PLANES = ("credits", "bets", "pending_bets", "payoffs", "minimum_raise_value", "seat_masks", "_meta", "_hand", "_episode")

    def save_state(self) -> dict:
        """A copy of the whole env state on the device (every plane of pkb_state); see load_state."""
        return {k: getattr(self, k).clone() for k in self._PLANES}

    def load_state(self, state: dict):
        """Put back a save_state() copy: the envs continue exactly from there (decks and agent draws are functions of the
        counters in the state).  Statistics are not part of the state."""
        for k in self._PLANES:
            getattr(self, k).copy_(state[k])
        self.refresh_valid()  # a library call after the foreign writes: the next step launch waits for them

    # ---- host snapshots (tests, scalar facade) -----------------------------------------------
    def export(self) -> dict:
        """Host copy of the whole state in the oracle's export layout ([P, N] planes, [N] vectors)."""
        meta = self._meta.cpu().numpy().astype(np.int64)
        return dict(
            credits=self.credits.cpu().numpy(), bets=self.bets.cpu().numpy(), pending=self.pending_bets.cpu().numpy(),
            payoffs=self.payoffs.contiguous().cpu().numpy(), states=self.player_states.cpu().numpy(),
            minraise=self.minimum_raise_value.cpu().numpy(), turn=(meta & 7).astype(np.int32),
            dealer=((meta >> 3) & 31).astype(np.int32), sb=((meta >> 8) & 31).astype(np.int32),
            bb=((meta >> 13) & 31).astype(np.int32), active=((meta >> 18) & 31).astype(np.int32),
            hand=self._hand.cpu().numpy().astype(np.int32), episode=self._episode.cpu().numpy().astype(np.uint32))
