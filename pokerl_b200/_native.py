"""ctypes binding of the C ABI in include/pokerl_b200.h (libpokerl_b200.so, built in-tree).

There is no CPU fallback: if the shared library is missing and cannot be built, importing this
module raises; if no CUDA device is present, every compute call raises RuntimeError.
"""
import ctypes as C
import os

from . import _build

ABI_VERSION = 4
MAX_PLAYERS = 10
NUM_MOVES = 7

F_GAME_OVER, F_HAND_OVER, F_ROUND_OVER, F_INVALID, F_POISONED = 1, 2, 4, 64, 128
F_SKIPPED, ACTION_SKIP = 32, -2 ** 31  # PKB_F_SKIPPED, PKB_ACTION_SKIP
NEXT_TURN_CAP = 64
ENV_ROUND_CAP = 1024


class PkbConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("device", C.c_int32),
        ("num_envs", C.c_uint64),
        ("env_id_base", C.c_uint64),
        ("num_players", C.c_int32),
        ("dealer", C.c_int32),
        ("start_credits", C.c_double * MAX_PLAYERS),
        ("big_blind", C.c_double),
        ("small_blind", C.c_double),
        ("seed", C.c_uint64),
        ("auto_reset", C.c_int32),
        ("reserved", C.c_int32),
    ]


class PkbStateView(C.Structure):
    _fields_ = [
        ("num_envs", C.c_uint64),
        ("stride", C.c_uint64),
        ("num_players", C.c_int32),
        ("payoff_pitch", C.c_int32),
        ("credits", C.c_void_p),
        ("bets", C.c_void_p),
        ("pending_bets", C.c_void_p),
        ("payoffs", C.c_void_p),
        ("minimum_raise_value", C.c_void_p),
        ("seat_masks", C.c_void_p),
        ("meta", C.c_void_p),
        ("hand", C.c_void_p),
        ("episode", C.c_void_p),
    ]


class PkbStats(C.Structure):
    _fields_ = [
        ("steps", C.c_uint64),
        ("hands", C.c_uint64),
        ("showdowns", C.c_uint64),
        ("games", C.c_uint64),
        ("invalid", C.c_uint64),
        ("poisoned", C.c_uint64),
        ("payoff_sum", C.c_double * MAX_PLAYERS),
    ]


# every symbol include/pokerl_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _u32, _u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
SYMBOLS = {
    "pkb_last_error": (C.c_char_p, []),
    "pkb_abi_version": (C.c_int, []),
    "pkb_device_count": (C.c_int, []),
    "pkb_create": (C.c_int, [C.POINTER(PkbConfig), C.POINTER(_vp)]),
    "pkb_destroy": (C.c_int, [_vp]),
    "pkb_reset": (C.c_int, [_vp, _vp, _vp]),
    "pkb_set_reset_dealer": (C.c_int, [_vp, _i32]),
    "pkb_refresh_valid": (C.c_int, [_vp, _vp]),
    "pkb_set_pipelining": (C.c_int, [_vp, _i32]),
    "pkb_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pkb_step_u8": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pkb_step_packed": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pkb_env_step": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "pkb_env_reset": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp]),
    "pkb_set_env_phases": (C.c_int, [_vp, C.c_int64]),
    "pkb_step_random": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "pkb_step_random_n": (C.c_int, [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "pkb_sample_actions": (C.c_int, [_vp, _u32, _vp, _vp, _vp]),
    "pkb_valid_actions": (C.c_int, [_vp, _i32, _vp, _vp]),
    "pkb_game_over": (C.c_int, [_vp, _vp, _vp]),
    "pkb_state": (C.c_int, [_vp, C.POINTER(PkbStateView)]),
    "pkb_player_states": (C.c_int, [_vp, _vp, _vp]),
    "pkb_cards": (C.c_int, [_vp, _vp, _vp]),
    "pkb_hand_features": (C.c_int, [_vp, _i32, _vp, _vp]),
    "pkb_stats": (C.c_int, [_vp, C.POINTER(PkbStats), _vp]),
    "pkb_stats_reset": (C.c_int, [_vp, _vp]),
    "pkb_deal": (C.c_int, [_u64, _u32, _u32, _u32, _i32, _vp]),
    "pkb_eval_cards": (C.c_int, [_vp, _i32, _u64, _vp, _vp]),
    "pkb_eval_masks": (C.c_int, [_vp, _u64, _vp, _vp]),
    "pkb_eval_exhaustive": (C.c_int, [_i32, _u64, _u64, _vp, _vp, _vp, _vp]),
    "pkb_random_hands": (C.c_int, [_u64, _u64, _vp, _vp]),
    "pkb_compare_rankings": (C.c_int, [_vp, _i32, _u64, _vp, _vp]),
}

_lib = None


def library_path() -> str:
    return _build.LIB


def lib():
    """The loaded shared library; built on first use when nvcc is available, otherwise an error."""
    global _lib
    if _lib is None:
        path = os.environ.get("PKB_LIB") or _build.LIB  # PKB_LIB: load an experimental build instead
        if not os.environ.get("PKB_LIB"):
            if not os.path.exists(path) or os.environ.get("PKB_AUTO_REBUILD") == "1":
                try:  # build when the library is missing (or, on request, older than its sources)
                    _build.build_native()
                except Exception as e:  # no silent fallback: the CUDA extension IS the product
                    if not os.path.exists(path):
                        raise ImportError(
                            f"pokerl_b200: CUDA extension {path} is missing and could not be built ({e}). "
                            "Run `python -m pokerl_b200._build` on a machine with nvcc.") from e
            if _build.is_stale():
                # Not rebuilt behind the caller's back: several ranks of one job would all do it at once.  Say so loudly.
                import warnings
                warnings.warn(f"pokerl_b200: {path} is older than its sources; run `python -m pokerl_b200._build` "
                              "(or set PKB_AUTO_REBUILD=1)", RuntimeWarning, stacklevel=2)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.pkb_abi_version() != ABI_VERSION:
            raise ImportError(f"pokerl_b200: ABI mismatch (library {L.pkb_abi_version()}, binding {ABI_VERSION})")
        _lib = L
    return _lib


class PkbError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise PkbError(lib().pkb_last_error().decode("utf-8", "replace"))


def require_cuda():
    n = lib().pkb_device_count()
    if n <= 0:
        raise PkbError("pokerl_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback"
                       + (": " + lib().pkb_last_error().decode() if n < 0 else ""))
    return n


def ptr(t) -> int:
    """Device (or pinned host) address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def deal(seed: int, env_id: int, episode: int, hand: int, num_players: int):
    """Host replay of one deal -> list of 5+2P card values (board, then hole cards seat by seat)."""
    buf = (C.c_uint8 * (5 + 2 * num_players))()
    check(lib().pkb_deal(seed, env_id, episode, hand, num_players, C.cast(buf, C.c_void_p)))
    return list(buf)
