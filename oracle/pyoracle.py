"""ctypes bindings for the CPU oracle (oracle/pokerl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker / CPU comparator.  The product
package (pokerl_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpokerl_oracle.so")

MAX_PLAYERS = 23
F_GAME_OVER, F_HAND_OVER, F_ROUND_OVER, F_INVALID, F_POISONED = 1, 2, 4, 64, 128
F_SKIPPED, ACTION_SKIP = 32, -2 ** 31


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources involved)."""
    src = [os.path.join(_HERE, f) for f in ("pokerl_oracle.c", "pokerl_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


class Game(C.Structure):
    _fields_ = [
        ("num_players", C.c_int),
        ("start_credits", C.c_double * MAX_PLAYERS),
        ("big_blind", C.c_double),
        ("small_blind", C.c_double),
        ("cfg_dealer", C.c_int),
        ("states", C.c_uint8 * MAX_PLAYERS),
        ("credits", C.c_double * MAX_PLAYERS),
        ("bets", C.c_double * MAX_PLAYERS),
        ("pending", C.c_double * MAX_PLAYERS),
        ("payoffs", C.c_double * MAX_PLAYERS),
        ("minimum_raise_value", C.c_double),
        ("turn", C.c_int),
        ("hand", C.c_int),
        ("dealer", C.c_int),
        ("small_blind_idx", C.c_int),
        ("big_blind_idx", C.c_int),
        ("active_player", C.c_int),
        ("deck", C.c_uint8 * (5 + 2 * MAX_PLAYERS)),
        ("seed", C.c_uint64),
        ("env_id", C.c_uint32),
        ("episode", C.c_uint32),
        ("deck_queue", C.c_void_p),
        ("deck_stride", C.c_int),
        ("deck_count", C.c_int),
        ("deck_cursor", C.c_int),
        ("next_turn_calls", C.c_int),
        ("end_hand_calls", C.c_int),
        ("showdown_layers", C.c_int),
        ("poisoned", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u32p, u64p, i32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_double))
        L.ora_eval_cards.restype = C.c_uint32
        L.ora_eval_cards.argtypes = [C.c_void_p, C.c_int]
        L.ora_eval_cards_batch.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        L.ora_eval_mask.restype = C.c_uint32
        L.ora_eval_mask.argtypes = [C.c_uint64]
        L.ora_eval_masks_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ora_eval_exhaustive.argtypes = [C.c_int, u64p, u64p, u64p]
        L.ora_eval_enumerate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ora_compare_rankings.restype = C.c_int
        L.ora_compare_rankings.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.ora_deal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.ora_sample_action.restype = C.c_int
        L.ora_sample_action.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint, C.c_void_p]
        L.ora_np_sum.restype = C.c_double
        L.ora_np_sum.argtypes = [C.c_void_p, C.c_int]
        gp = C.POINTER(Game)
        L.ora_game_init.argtypes = [gp, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint32]
        L.ora_game_set_decks.argtypes = [gp, C.c_void_p, C.c_int, C.c_int]
        L.ora_game_reset.argtypes = [gp]
        L.ora_game_valid_mask.restype = C.c_uint
        L.ora_game_valid_mask.argtypes = [gp, C.c_int]
        L.ora_game_is_over.restype = C.c_int
        L.ora_game_is_over.argtypes = [gp]
        L.ora_game_step.restype = C.c_uint
        L.ora_game_step.argtypes = [gp, C.c_int]
        L.ora_batch_create.restype = C.c_void_p
        L.ora_batch_create.argtypes = [C.c_size_t, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64,
                                       C.c_uint32, C.c_int]
        L.ora_batch_destroy.argtypes = [C.c_void_p]
        L.ora_batch_set_threads.argtypes = [C.c_int]
        L.ora_batch_reset.argtypes = [C.c_void_p]
        L.ora_batch_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_batch_valid.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_batch_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ora_batch_run_random.restype = C.c_uint64
        L.ora_batch_run_random.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ora_batch_export.argtypes = [C.c_void_p] + [C.c_void_p] * 13
        L.ora_batch_game.restype = gp
        L.ora_batch_game.argtypes = [C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- judger

def eval_cards(cards) -> int:
    """packed = (category << 20) | kickers for a list of card values (suit<<4 | rank0)."""
    a = np.ascontiguousarray(cards, dtype=np.uint8)
    return int(lib().ora_eval_cards(_ptr(a), len(a)))


def eval_cards_batch(cards: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(cards, dtype=np.uint8)
    n, k = a.shape
    out = np.empty(n, dtype=np.uint32)
    lib().ora_eval_cards_batch(_ptr(a), k, n, _ptr(out))
    return out


def eval_masks_batch(masks: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(masks, dtype=np.uint64)
    out = np.empty(a.shape[0], dtype=np.uint32)
    lib().ora_eval_masks_batch(_ptr(a), a.shape[0], _ptr(out))
    return out


def eval_exhaustive(k: int):
    hist = (C.c_uint64 * 11)()
    a, b = C.c_uint64(), C.c_uint64()
    lib().ora_eval_exhaustive(k, hist, C.byref(a), C.byref(b))
    return [int(x) for x in hist], int(a.value), int(b.value)


def eval_enumerate(k: int, first: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint32)
    lib().ora_eval_enumerate(k, first, count, _ptr(out))
    return out


def unpack(packed: int):
    """packed -> (category, [kickers]) as judger.eval_hand returns them."""
    cat, v = packed >> 20, packed & 0xFFFFF
    ks = []
    while v:
        ks.append(v & 0xF)
        v >>= 4
    return cat, ks[::-1]


def compare_rankings(packed):
    a = np.ascontiguousarray(packed, dtype=np.uint32)
    onehot = np.zeros(len(a), dtype=np.uint8)
    lib().ora_compare_rankings(_ptr(a), len(a), _ptr(onehot))
    return onehot.tolist()


# ------------------------------------------------------- philox / dealing

def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ora_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def deal(seed: int, env_id: int, episode: int, hand: int, num_players: int) -> np.ndarray:
    out = np.empty(5 + 2 * num_players, dtype=np.uint8)
    lib().ora_deal(seed, env_id, episode, hand, num_players, _ptr(out))
    return out


def sample_action(seed, env_id, step_idx, valid_mask, weights=None) -> int:
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
    return int(lib().ora_sample_action(seed, env_id, step_idx, valid_mask, _ptr(w)))


def np_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().ora_np_sum(_ptr(a), len(a)))


# ------------------------------------------------------------------ game

def _credits(num_players, start_credits):
    if np.isscalar(start_credits):
        return np.full(num_players, float(start_credits), dtype=np.float64)
    a = np.ascontiguousarray(start_credits, dtype=np.float64)
    assert a.shape == (num_players,)
    return a


class OracleGame:
    """One scalar game; optional injected decks (u8 [n_decks, >=5+2P])."""

    def __init__(self, num_players=4, start_credits=100, big_blind=2, small_blind=1, dealer=0, seed=0, env_id=0,
                 decks=None):
        self.g = Game()
        sc = _credits(num_players, start_credits)
        lib().ora_game_init(C.byref(self.g), num_players, _ptr(sc), float(big_blind), float(small_blind), dealer, seed,
                            env_id)
        self._decks = None
        if decks is not None:
            self._decks = np.ascontiguousarray(decks, dtype=np.uint8)
            lib().ora_game_set_decks(C.byref(self.g), _ptr(self._decks), self._decks.shape[1], self._decks.shape[0])

    def reset(self):
        lib().ora_game_reset(C.byref(self.g))

    def valid_mask(self, player=-1) -> int:
        return int(lib().ora_game_valid_mask(C.byref(self.g), player))

    def step(self, action: int) -> int:
        return int(lib().ora_game_step(C.byref(self.g), int(action)))

    @property
    def game_over(self) -> bool:
        return bool(lib().ora_game_is_over(C.byref(self.g)))

    def snapshot(self) -> dict:
        g, P = self.g, self.g.num_players
        return dict(
            states=list(g.states[:P]), credits=list(g.credits[:P]), bets=list(g.bets[:P]), pending=list(g.pending[:P]),
            payoffs=list(g.payoffs[:P]), minraise=g.minimum_raise_value, turn=g.turn, hand=g.hand, dealer=g.dealer,
            sb=g.small_blind_idx, bb=g.big_blind_idx, active=g.active_player)


class OracleBatch:
    """N lockstep scalar games, Philox-dealt, keyed by global env id = env_id_base + i."""

    def __init__(self, num_envs, num_players=4, start_credits=100, big_blind=2, small_blind=1, dealer=0, seed=0,
                 env_id_base=0, auto_reset=False, threads=0):
        self.n, self.P = int(num_envs), int(num_players)
        sc = _credits(num_players, start_credits)
        if threads:
            lib().ora_batch_set_threads(threads)
        self.h = lib().ora_batch_create(self.n, self.P, _ptr(sc), float(big_blind), float(small_blind), dealer, seed,
                                        env_id_base, int(auto_reset))
        if not self.h:
            raise ValueError("ora_batch_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_batch_destroy(self.h)
            self.h = None

    def reset(self):
        lib().ora_batch_reset(self.h)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        flags = np.empty(self.n, dtype=np.uint8)
        valid = np.empty(self.n, dtype=np.uint8)
        lib().ora_batch_step(self.h, _ptr(a), _ptr(flags), _ptr(valid))
        return flags, valid

    def valid(self):
        v = np.empty(self.n, dtype=np.uint8)
        lib().ora_batch_valid(self.h, _ptr(v))
        return v

    def sample(self, step_idx, weights=None):
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
        a = np.empty(self.n, dtype=np.int32)
        lib().ora_batch_sample(self.h, step_idx, _ptr(w), _ptr(a))
        return a

    def run_random(self, step0, steps, weights=None) -> int:
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)
        return int(lib().ora_batch_run_random(self.h, step0, steps, _ptr(w)))

    def export(self) -> dict:
        n, P = self.n, self.P
        d = dict(
            credits=np.empty((P, n)), bets=np.empty((P, n)), pending=np.empty((P, n)), payoffs=np.empty((P, n)),
            states=np.empty((P, n), dtype=np.uint8), minraise=np.empty(n), turn=np.empty(n, dtype=np.int32),
            hand=np.empty(n, dtype=np.int32), dealer=np.empty(n, dtype=np.int32), sb=np.empty(n, dtype=np.int32),
            bb=np.empty(n, dtype=np.int32), active=np.empty(n, dtype=np.int32), episode=np.empty(n, dtype=np.uint32))
        lib().ora_batch_export(self.h, *[_ptr(d[k]) for k in (
            "credits", "bets", "pending", "payoffs", "states", "minraise", "turn", "hand", "dealer", "sb", "bb",
            "active", "episode")])
        return d

    def game(self, i) -> Game:
        return lib().ora_batch_game(self.h, i).contents
