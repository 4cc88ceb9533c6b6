"""GPU (-m gpu): the CUDA judger (K4) through the C ABI against the oracle and the golden vectors."""
import json
import os
from math import comb

import numpy as np
import pytest

from tests import kat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def hc():
    from tests.hostcheck import load
    return load()


def test_native_library_loaded(torch_cuda):
    from pokerl_b200 import _native as nat
    assert nat.require_cuda() >= 1
    maps = open("/proc/self/maps").read()
    assert "libpokerl_b200.so" in maps, "the in-tree CUDA extension is not the code that is running"


def test_eval_kats_through_facade(torch_cuda):
    from pokerl_b200 import Card, eval_hand
    from pokerl_b200.judger import compare_hands
    for cards, expect in kat.EVAL_KATS:  # reference tests/pokerl/test_judger.py:14-78
        assert eval_hand([Card(s) for s in cards.split()]) == expect, cards
    for hands, onehot in kat.COMPARE_KATS:  # test_judger.py:83-117
        out = compare_hands([[Card(s) for s in h.split()] for h in hands])
        assert out[0] == onehot, hands
    assert eval_hand([]) == (10, [])
    assert eval_hand([Card("AS")]) == (9, [13])


def test_eval_reference_vectors(torch_cuda, golden_dir):
    torch = torch_cuda
    from pokerl_b200.judger import eval_cards_batch
    d = np.load(os.path.join(golden_dir, "eval_vectors.npz"))
    for k in range(8):
        sel = d["counts"] == k
        cards = torch.from_numpy(np.ascontiguousarray(d["cards"][sel][:, :k])).cuda()
        got = eval_cards_batch(cards).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, d["packed"][sel]), k


def test_compare_reference_vectors(torch_cuda, golden_dir):
    torch = torch_cuda
    from pokerl_b200.judger import compare_rankings_batch
    for r in json.load(open(os.path.join(golden_dir, "compare_vectors.json"))):
        p = torch.tensor([r["packed"]], dtype=torch.int32).cuda()
        assert compare_rankings_batch(p)[0].tolist() == r["onehot"]


def test_eval_card_lists_with_duplicates_vs_oracle(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200.judger import eval_cards_batch
    rng = np.random.default_rng(17)
    for k in range(1, 8):
        ids = rng.integers(0, 52, size=(20000, k))
        cards = (((ids // 13) << 4) | (ids % 13)).astype(np.uint8)
        got = eval_cards_batch(torch.from_numpy(cards).cuda()).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, ora.eval_cards_batch(cards)), k


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7])
def test_exhaustive_all_hands_bit_exact(torch_cuda, ora, k):
    """Every C(52,k) hand, enumerated on the device: element-wise equal to the oracle, and (k >= 5)
    equal to the golden histogram / checksums computed from the reference's judger.py."""
    torch = torch_cuda
    from pokerl_b200.judger import eval_exhaustive
    n = comb(52, k)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    hist, a, b = eval_exhaustive(k, out=out)
    if k in kat.EXHAUSTIVE_GOLDEN:
        assert (hist, a, b) == kat.EXHAUSTIVE_GOLDEN[k]
    assert sum(hist) == n
    got = out.cpu().numpy().view(np.uint32)
    step = 1 << 24
    for lo in range(0, n, step):
        cnt = min(step, n - lo)
        assert np.array_equal(got[lo:lo + cnt], ora.eval_enumerate(k, lo, cnt)), (k, lo)


def test_exhaustive_subrange(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200.judger import eval_exhaustive
    first, count = 77_777_777, 100_003
    out = torch.empty(count, dtype=torch.int32, device="cuda")
    hist, a, b = eval_exhaustive(7, first=first, count=count, out=out)
    ref = ora.eval_enumerate(7, first, count)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref)
    assert a == int(ref.astype(np.uint64).sum()) and sum(hist) == count


def test_random_hands_and_mask_eval(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200.judger import eval_masks_batch, random_hands
    masks = random_hands(0, 1 << 20)
    m = masks.cpu().numpy().view(np.uint64)
    assert (np.bitwise_count(m) == 7).all() and (m < (1 << 52)).all()
    assert len(np.unique(m)) > 0.99 * len(m)
    got = eval_masks_batch(masks).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, ora.eval_masks_batch(m))
    again = random_hands(0, 1 << 20)
    assert torch.equal(masks, again)  # counter-based: same seed, same hands


def test_mask_eval_on_every_seven_card_hand(torch_cuda, hc):
    """The streaming batch evaluator (eval_masks_fast_kernel: shared-memory perfect hash, rare hands queued per warp)
    on ALL C(52,7) masks, in the enumeration order of the exhaustive kernel: element-wise equal to that kernel's
    output (which test_exhaustive_all_hands_bit_exact pins to the oracle and the golden checksums)."""
    import ctypes as C
    torch = torch_cuda
    from pokerl_b200.judger import eval_exhaustive, eval_masks_batch
    n = comb(52, 7)
    want = torch.empty(n, dtype=torch.int32, device="cuda")
    eval_exhaustive(7, out=want)
    hc.hc_enumerate_masks.restype = C.c_uint64
    hc.hc_enumerate_masks.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    step, start = 1 << 25, 0
    buf = np.empty(step, dtype=np.uint64)
    for lo in range(0, n, step):
        cnt = min(step, n - lo)
        start = hc.hc_enumerate_masks(7, start, cnt, buf.ctypes.data)
        got = eval_masks_batch(torch.from_numpy(buf[:cnt].view(np.int64)).cuda())
        assert torch.equal(got, want[lo:lo + cnt]), lo


def test_mask_eval_mixed_sizes_ragged_and_unaligned(torch_cuda, ora):
    """Masks of 0..7 cards mixed in one batch (everything but seven cards takes the queue of rare hands), batch sizes
    around the streaming kernel's trip size and threshold, and arrays that start at odd elements (the unaligned path)."""
    torch = torch_cuda
    from pokerl_b200 import _native as nat
    rng = np.random.default_rng(5)
    n = (1 << 18) + 77
    k = rng.integers(0, 8, size=n)
    k[: n // 2] = 7
    rng.shuffle(k)
    order = np.argsort(rng.random((n, 52)), axis=1)
    bits = (np.arange(52)[None, :] < k[:, None])
    ids = np.where(bits, order, 63)
    m = np.zeros(n, dtype=np.uint64)
    for j in range(7):
        sel = ids[:, j] < 52
        m[sel] |= np.uint64(1) << ids[sel, j].astype(np.uint64)
    assert (np.bitwise_count(m) == k).all()
    want = ora.eval_masks_batch(m)
    dm = torch.from_numpy(m.view(np.int64)).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    for lo, cnt in ((0, n), (0, 1 << 17), (0, (1 << 17) + 127), (0, (1 << 17) - 1), (1, (1 << 17) + 5), (3, n - 3), (2, 1 << 17), (0, 1000)):
        out = torch.full((cnt + 3,), -1, dtype=torch.int32, device="cuda")
        for off in (0, 1):
            masks = dm[lo:lo + cnt]
            nat.check(nat.lib().pkb_eval_masks(masks.data_ptr(), cnt, out[off:].data_ptr(), stream))
            got = out[off:off + cnt].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want[lo:lo + cnt]), (lo, cnt, off)
            assert (out[off + cnt:] == -1).all()


def test_empty_and_ragged_inputs(torch_cuda):
    torch = torch_cuda
    from pokerl_b200.judger import eval_cards_batch, eval_masks_batch
    assert eval_cards_batch(torch.empty((0, 7), dtype=torch.uint8, device="cuda")).numel() == 0
    assert eval_masks_batch(torch.empty(0, dtype=torch.int64, device="cuda")).numel() == 0
    out = eval_cards_batch(torch.empty((5, 0), dtype=torch.uint8, device="cuda"))
    assert out.tolist() == [10 << 20] * 5
