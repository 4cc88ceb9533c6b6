"""CPU: the product's evaluator / Philox headers (compiled for the host by tests/hostcheck) against the
oracle, the C-ABI library's exports, and the host-side facade logic that needs no GPU."""
import ctypes as C
import json
import os
import re
from math import comb

import numpy as np
import pytest

from tests import kat
from tests.hostcheck import load as load_hostcheck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc():
    return load_hostcheck()


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_product_evaluator_exhaustive(hc, ora, k):
    """Every k-card hand: product bitmask evaluator == oracle, element by element."""
    n = comb(52, k)
    hist = np.zeros(11, dtype=np.uint64)
    sums = np.zeros(2, dtype=np.uint64)
    out = np.empty(n, dtype=np.uint32)
    hc.hc_eval_exhaustive(k, hist.ctypes.data, sums.ctypes.data, out.ctypes.data, n)
    assert np.array_equal(out, ora.eval_enumerate(k, 0, n))
    if k in kat.EXHAUSTIVE_GOLDEN:
        h, a, b = kat.EXHAUSTIVE_GOLDEN[k]
        assert hist.tolist() == h and int(sums[0]) == a and int(sums[1]) == b


def test_product_evaluator_7_sampled(hc, ora):
    rng = np.random.default_rng(11)
    ids = np.argsort(rng.random((400000, 52)), axis=1)[:, :7]
    masks = (np.uint64(1) << ids.astype(np.uint64)).sum(axis=1).astype(np.uint64)
    out = np.empty(len(masks), dtype=np.uint32)
    hc.hc_eval_masks(masks.ctypes.data, len(masks), out.ctypes.data)
    assert np.array_equal(out, ora.eval_masks_batch(masks))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7])
def test_table_evaluator_equals_bitmask_evaluator(hc, ora, k):
    """The table-driven, branch-free form used by the batch kernels returns exactly what the bitmask evaluator
    returns: all C(52,k) hands for k <= 6, the first 8M seven-card hands (the GPU test covers all of them)."""
    import math
    count = math.comb(52, k) if k <= 6 else 8_000_000
    assert hc.hc_eval_tbl_mismatches(k, count) == 0
    assert hc.hc_eval_mask_tbl(0) == int(ora.eval_masks_batch(np.zeros(1, dtype=np.uint64))[0])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7])
def test_select_evaluator_equals_bitmask_evaluator(hc, k):
    """The select-only form the step kernels inline (pkb_eval.cuh eval_mask52_sel) returns exactly what the bitmask
    evaluator returns: all C(52,k) hands for k <= 6, the first 8M seven-card hands."""
    import math
    count = math.comb(52, k) if k <= 6 else 8_000_000
    hc.hc_eval_sel_mismatches.restype = C.c_uint64
    hc.hc_eval_sel_mismatches.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    assert hc.hc_eval_sel_mismatches(k, 0, count) == 0
    hc.hc_eval_mask_sel.restype = C.c_uint32
    hc.hc_eval_mask_sel.argtypes = [C.c_uint64]
    assert hc.hc_eval_mask_sel(0) == 10 << 20


def test_perfect_hash_evaluator_equals_bitmask_evaluator(hc):
    """The seven-card perfect-hash evaluator of the batch kernels (pkb_eval.cuh eval7_hash3: rank keys with distinct
    multiset sums, displacement-hashed dense value table, flush table, the judger's special cases as selects), tables
    built by the product's own host code: equal to the bitmask evaluator on ALL C(52,7) hands.  A negative result means
    the tables could not be built (colliding keys / no multiplier for a bucket)."""
    hc.hc_eval_hash3_mismatches.restype = C.c_longlong
    hc.hc_eval_hash3_mismatches.argtypes = [C.c_uint64]
    assert hc.hc_eval_hash3_mismatches(133_784_560) == 0


def test_streaming_hash_evaluator_equals_bitmask_evaluator_or_says_rare(hc):
    """The streaming form of the perfect hash (pkb_eval.cuh eval7_fast: key tables that also count cards per suit,
    16-bit first level per 1024 key sums, rank-wheel mark in the value table): on ALL C(52,7) hands it either marks the
    hand rare (a suit with five cards, or with 2,3,4,5 and no six: the batch kernel ranks those with the select evaluator) or returns
    exactly the bitmask evaluator's value; masks with any other number of cards are always marked rare."""
    hc.hc_eval_fast_mismatches.restype = C.c_longlong
    hc.hc_eval_fast_mismatches.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64]
    rare = C.c_uint64(0)
    assert hc.hc_eval_fast_mismatches(133_784_560, C.byref(rare), 200_000) == 0
    # flushes (incl. straight flushes): 4,089,228 + ... ; what matters here: rare hands are a small minority
    assert 0 < rare.value < 0.035 * 133_784_560, rare.value


def test_all_at_once_deal_equals_sequential_dealer(hc):
    """deal_ids (interleaved Philox blocks + sorted draw chain, P <= 4) deals exactly the cards of the sequential Dealer."""
    rng = np.random.default_rng(3)
    hc.hc_deal_ids.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    for P in (1, 2, 3, 4):
        for _ in range(20000):
            seed = int(rng.integers(0, 2**63)); env = int(rng.integers(0, 2**32)); ep = int(rng.integers(0, 2**20)); hand = int(rng.integers(0, 2**16))
            a = np.zeros(5 + 2 * P, dtype=np.uint8); b = np.zeros(5 + 2 * P, dtype=np.uint8)
            hc.hc_deal(seed, env, ep, hand, P, a.ctypes.data)
            assert hc.hc_deal_ids(seed, env, ep, hand, P, b.ctypes.data) == 0
            assert np.array_equal(a, b), (P, seed, env, ep, hand, a, b)


def test_product_card_list_paths(hc, ora, golden_dir):
    d = np.load(os.path.join(golden_dir, "eval_vectors.npz"))
    for c, k, p in zip(d["cards"], d["counts"], d["packed"]):
        c = np.ascontiguousarray(c)
        assert hc.hc_eval_cards(c.ctypes.data, int(k)) == p
        assert hc.hc_eval_cards_literal(c.ctypes.data, int(k)) == p
    for cards, (cat, kickers) in kat.EVAL_KATS:
        c = np.array(kat.hand_values(cards), dtype=np.uint8)
        assert ora.unpack(hc.hc_eval_cards(c.ctypes.data, len(c))) == (cat, kickers)
    rng = np.random.default_rng(5)
    for _ in range(50000):  # random lists WITH duplicates: literal path vs oracle
        k = int(rng.integers(0, 8))
        ids = rng.integers(0, 52, size=max(k, 1))
        c = np.ascontiguousarray((((ids // 13) << 4) | (ids % 13)).astype(np.uint8))
        assert hc.hc_eval_cards(c.ctypes.data, k) == ora.eval_cards(c[:k])


def test_product_compare(hc, golden_dir):
    for r in json.load(open(os.path.join(golden_dir, "compare_vectors.json"))):
        p = np.array(r["packed"], dtype=np.uint32)
        o = np.zeros(len(p), dtype=np.uint8)
        hc.hc_compare(p.ctypes.data, len(p), o.ctypes.data)
        assert o.tolist() == r["onehot"]


def test_product_deal_and_sampler(hc, ora):
    rng = np.random.default_rng(0)
    w = np.array([100, 300, 300, 50, 50, 50, 450], dtype=np.uint32)
    for _ in range(2000):
        seed, env = int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 32))
        ep, hand, P = int(rng.integers(0, 1000)), int(rng.integers(0, 1000)), int(rng.integers(2, 11))
        out = np.zeros(5 + 2 * P, dtype=np.uint8)
        hc.hc_deal(seed, env, ep, hand, P, out.ctypes.data)
        assert np.array_equal(out, ora.deal(seed, env, ep, hand, P))
        vm = int(rng.integers(0, 128)) | 0x41
        assert hc.hc_sample(seed, env, hand, vm, None) == ora.sample_action(seed, env, hand, vm)
        assert hc.hc_sample(seed, env, hand, vm, w.ctypes.data) == ora.sample_action(seed, env, hand, vm, w)


# ---- the C-ABI library -------------------------------------------------------------------------

def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pokerl_b200.h")).read()
    return sorted(set(re.findall(r"\b(pkb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pokerl_b200 import _native as nat
    L = C.CDLL(nat._build.build_native())
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/pokerl_b200.h but not exported"
    assert set(declared) == set(nat.SYMBOLS), set(declared) ^ set(nat.SYMBOLS)
    assert nat.lib().pkb_abi_version() == nat.ABI_VERSION


def test_host_deal_entry_point_matches_oracle(ora):
    from pokerl_b200 import _native as nat
    for P in (2, 9):
        for hand in (1, 2, 77):
            assert nat.deal(42, 7, 1, hand, P) == ora.deal(42, 7, 1, hand, P).tolist()


def test_abi_struct_layout_matches_header():
    """ctypes mirrors of the header structs: sizes as a C compiler lays them out."""
    import subprocess
    import tempfile
    from pokerl_b200 import _native as nat
    src = ('#include <stdio.h>\n#include "pokerl_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pkb_config), '
           'sizeof(pkb_state_view), sizeof(pkb_stats_t));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")],
                       check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(nat.PkbConfig), C.sizeof(nat.PkbStateView), C.sizeof(nat.PkbStats)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pokerl_b200 import _native as nat
    import pokerl_b200
    with pytest.raises(nat.PkbError):
        pokerl_b200.Game(num_players=2)
    with pytest.raises(nat.PkbError):
        pokerl_b200.eval_hand([pokerl_b200.Card("AS")])


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pokerl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "pokerl_oracle" not in text, f


def test_cards_and_enums_mirror_reference():
    from pokerl_b200 import Card, create_default_deck, PokerMoves, HandRanking
    from pokerl_b200.enums import PlayerState, CardRank
    assert Card("AS").value == Card("1S").value == Card((13, 0)).value == Card((0, 0)).value == 0
    assert Card("TH").value == (1 << 4) | 9 and Card("TH").rank == 9 and Card("AS").rank == 13
    assert Card("KC").id == 12 + 39 and Card.from_id(51) == Card("KC")
    deck = create_default_deck()
    assert [c.value for c in deck[:5]] == [0x00, 0x10, 0x20, 0x30, 0x01] and len({c.id for c in deck}) == 52
    assert repr(Card("AS")) == chr(0x1f0a1) and repr(Card("QS")) == chr(0x1f0a2 + 11)
    assert (PokerMoves.FOLD, PokerMoves.ALL_IN, PokerMoves.NUM_MOVES) == (0, 6, 7)
    assert (HandRanking.STRAIGHT_FLUSH, HandRanking.NONE) == (1, 10)
    assert (PlayerState.FOLDED, PlayerState.BROKEN) == (0, 4) and CardRank.ACE == 13
    with pytest.raises(AssertionError):
        Card(0x4f)


def test_env_and_agent_api_mirror_reference():
    """pokerl.envs / pokerl.agents names (pokerl/envs/__init__.py:1-2, agents/__init__.py:1-3) exist, the random
    agent keeps the reference's scalar behaviour, and the env refuses to run without a GPU (no CPU fallback)."""
    import numpy as np
    import pokerl_b200
    from pokerl_b200 import _native as nat
    from pokerl_b200.agents import PokerAgent, RandomAgent
    from pokerl_b200.envs import PokerEnv, PokerGameEnv
    assert issubclass(PokerGameEnv, PokerEnv) and issubclass(RandomAgent, PokerAgent)
    assert pokerl_b200.envs.PokerGameEnv is PokerGameEnv

    class S:
        valid_actions = np.array([1., 0., 1., 0., 0., 0., 1.])
    picks = {RandomAgent()(S()) for _ in range(200)}
    assert picks == {0, 2, 6}
    assert {RandomAgent(weights=[0, 5, 5, 5, 5, 5, 1])(S()) for _ in range(200)} == {2, 6}
    for cls in (PokerEnv, PokerAgent):
        with pytest.raises(NotImplementedError):
            (cls().reset() if cls is PokerEnv else cls()(None))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(nat.PkbError):
            PokerGameEnv(agents=[RandomAgent()] * 3, num_players=4)


def test_closed_form_winner_scan_equals_sequential_scan(hc):
    """winners_closed (the side-pot loop's compare_rankings, judger.py:133-158) == the sequential WinnerScan: every
    assignment of a small alphabet of rankings (NONE included) to 2..5 seats, and random draws for 6..10 seats."""
    import itertools
    hc.hc_winners_closed.restype = C.c_uint32
    hc.hc_winners_closed.argtypes = [C.c_void_p, C.c_int]
    hc.hc_winners_scan.restype = C.c_uint32
    hc.hc_winners_scan.argtypes = [C.c_void_p, C.c_int]
    alphabet = [10 << 20, 9 << 20 | 5, 9 << 20 | 7, 9 << 20 | 9, 8 << 20 | 3, 8 << 20 | 7, 1 << 20 | 13]
    for n in range(2, 6):
        for combo in itertools.product(alphabet, repeat=n):
            a = np.array(combo, dtype=np.uint32)
            assert hc.hc_winners_closed(a.ctypes.data, n) == hc.hc_winners_scan(a.ctypes.data, n), combo
    rng = np.random.default_rng(5)
    for n in range(6, 11):
        for _ in range(20000):
            a = (rng.integers(7, 11, n).astype(np.uint32) << 20) | rng.integers(0, 4, n).astype(np.uint32)
            a[a >> 20 == 10] = 10 << 20
            assert hc.hc_winners_closed(a.ctypes.data, n) == hc.hc_winners_scan(a.ctypes.data, n), a


def test_branch_free_select_bit(hc):
    hc.hc_select_bit_mismatches.restype = C.c_uint32
    hc.hc_select_bit_mismatches.argtypes = [C.c_uint64, C.c_uint64]
    assert hc.hc_select_bit_mismatches(3, 20000) == 0
