"""Hand ranking on the GPU behind the reference's judger API (pokerl/judger.py:7-189).

`eval_hand`, `get_kickers_value`, `compare_rankings` and `compare_hands` keep the reference's
signatures and return types for single hands; the `*_batch` functions are the same operators
over device tensors.  All ranking runs in libpokerl_b200.so (K4); nothing here evaluates
hands on the CPU.
"""
import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _native as nat
from .cards import Card
from .enums import HandRanking


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _cuda_device(device=None):
    nat.require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def unpack_ranking(packed: int) -> Tuple[int, List[int]]:
    """packed (category << 20 | kicker nibbles) -> (category, [kickers]) as eval_hand returns them.

    Kicker ranks are >= 1, so the number of kickers is the number of nibbles up to the top one.
    """
    packed = int(packed) & 0xFFFFFFFF
    cat, v = packed >> 20, packed & 0xFFFFF
    kickers = []
    while v:
        kickers.append(v & 0xF)
        v >>= 4
    return cat, kickers[::-1]


def get_kickers_value(kickers) -> int:
    """Nibble-pack kickers, first one most significant (judger.py:101-109)."""
    value = 0
    for k in kickers:
        value = (value << 4) | int(k)
    return value


def pack_ranking(ranking: Tuple[int, List[int]]) -> int:
    return (int(ranking[0]) << 20) | get_kickers_value(ranking[1])


# ---- batched operators ------------------------------------------------------------------------

def eval_cards_batch(cards: torch.Tensor) -> torch.Tensor:
    """eval_hand on a uint8 [n, k] tensor of card values (k <= 7, duplicates allowed) -> int32 [n] packed."""
    dev = _cuda_device(cards.device if cards.is_cuda else None)
    cards = cards.to(device=dev, dtype=torch.uint8).contiguous()
    n, k = cards.shape
    out = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().pkb_eval_cards(nat.ptr(cards) if k else None, k, n, nat.ptr(out), _stream(dev)))
    return out


def eval_masks_batch(masks: torch.Tensor) -> torch.Tensor:
    """eval_hand on int64 [n] 52-bit card-id masks (<= 7 bits set) -> int32 [n] packed."""
    dev = _cuda_device(masks.device if masks.is_cuda else None)
    masks = masks.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty(masks.shape[0], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().pkb_eval_masks(nat.ptr(masks), masks.shape[0], nat.ptr(out), _stream(dev)))
    return out


def eval_exhaustive(k: int, first: int = 0, count: int = None, out: torch.Tensor = None, device=None):
    """All C(52,k) hands enumerated on the device -> (hist[11], sumA, sumB); see include/pokerl_b200.h."""
    from math import comb
    dev = _cuda_device(device)
    count = comb(52, k) - first if count is None else count
    hist = (C.c_uint64 * 11)()
    sums = (C.c_uint64 * 2)()
    with torch.cuda.device(dev):
        nat.check(nat.lib().pkb_eval_exhaustive(k, first, count, C.cast(hist, C.c_void_p), C.cast(sums, C.c_void_p),
                                                nat.ptr(out), _stream(dev)))
    return [int(x) for x in hist], int(sums[0]), int(sums[1])


def random_hands(seed: int, n: int, device=None) -> torch.Tensor:
    """n random 7-card hands as int64 id masks, generated on the device from Philox(seed, index)."""
    dev = _cuda_device(device)
    masks = torch.empty(n, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().pkb_random_hands(seed, n, nat.ptr(masks), _stream(dev)))
    return masks


def compare_rankings_batch(packed: torch.Tensor) -> torch.Tensor:
    """compare_rankings on int32 [n, p] packed rankings -> uint8 [n, p] one-hot winners."""
    dev = _cuda_device(packed.device if packed.is_cuda else None)
    packed = packed.to(device=dev, dtype=torch.int32).contiguous()
    n, p = packed.shape
    onehot = torch.empty((n, p), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().pkb_compare_rankings(nat.ptr(packed), p, n, nat.ptr(onehot), _stream(dev)))
    return onehot


# ---- the reference's scalar API -----------------------------------------------------------------

def eval_hand(hand: List[Card]) -> Tuple[int, List[int]]:
    """Evaluate up to seven cards -> (HandRanking, kickers) (judger.py:7-99), on the GPU."""
    if len(hand) > 7:
        raise ValueError("eval_hand: at most seven cards")
    values = torch.tensor([[c.value for c in hand]], dtype=torch.uint8).reshape(1, len(hand))
    return unpack_ranking(eval_cards_batch(values).item())


def compare_rankings(rankings: List[Tuple[int, List[int]]]) -> Tuple[List[int], List[int]]:
    """Winners among rankings -> (onehot, winner indices) (judger.py:111-158, including line 148)."""
    packed = torch.tensor([[pack_ranking(r) for r in rankings]], dtype=torch.int32)
    onehot = compare_rankings_batch(packed)[0].tolist()
    return onehot, [i for i, w in enumerate(onehot) if w]


def compare_hands(hands: List[List[Card]]):
    """eval_hand on each hand, then compare_rankings (judger.py:160-189)."""
    rankings = [eval_hand(h) for h in hands]
    return compare_rankings(rankings) + (rankings,)


__all__ = ["eval_hand", "get_kickers_value", "compare_rankings", "compare_hands", "eval_cards_batch",
           "eval_masks_batch", "eval_exhaustive", "random_hands", "compare_rankings_batch", "unpack_ranking",
           "pack_ranking", "HandRanking"]
