"""The pattern builder's hand-written primitives (csrc/sortscan.cuh) against numpy: the stable LSD radix sort
must reproduce np.argsort(kind='stable') exactly -- the order np.bincount later sums in (ref fem/space.py:113-131) --
and the scans np.cumsum, at sizes around the tile boundaries and for skewed key distributions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sort(keys, end_bit):
    from fes_b200 import _lib
    dev = _lib.require_cuda()
    k = torch.as_tensor(keys.astype(np.int64)).to(dev)
    v = torch.arange(len(keys), dtype=torch.int64).to(torch.int32).to(dev)
    _lib.check(_lib.lib().fes_sort_pairs_u64(_lib.ptr(k), _lib.ptr(v), len(keys), end_bit, _lib.stream()))
    return k.cpu().numpy().astype(np.uint64), v.cpu().numpy().astype(np.int64) & 0xffffffff


@pytest.mark.parametrize('n', [1, 2, 31, 255, 256, 257, 4095, 4096, 4097, 100003, 3_000_001])
@pytest.mark.parametrize('dist', ['uniform', 'few', 'sorted', 'reversed'])
def test_radix_sort_is_numpys_stable_argsort(n, dist):
    rng = np.random.default_rng(n)
    bits = 42
    if dist == 'uniform':
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64)
    elif dist == 'few':                                   # long runs of equal keys: stability is all that orders them
        keys = rng.integers(0, 7, n, dtype=np.uint64) * np.uint64(0x10101010101)
    elif dist == 'sorted':
        keys = np.sort(rng.integers(0, 1 << bits, n, dtype=np.uint64))
    else:
        keys = np.sort(rng.integers(0, 1 << bits, n, dtype=np.uint64))[::-1].copy()
    got_k, got_v = _sort(keys, bits + 3)                  # a bit count that is not a multiple of 8
    order = np.argsort(keys, kind='stable')
    assert np.array_equal(got_v, order)
    assert np.array_equal(got_k, keys[order])


def test_radix_sort_ignores_bits_above_end_bit():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 1 << 40, 50000, dtype=np.uint64)
    got_k, got_v = _sort(keys, 16)
    order = np.argsort(keys & np.uint64(0xffff), kind='stable')
    assert np.array_equal(got_v, order)


@pytest.mark.parametrize('n', [1, 2047, 2048, 2049, 2048 * 2048 + 5, 9_000_000])
def test_exclusive_scan_is_cumsum(n):
    from fes_b200 import _lib
    dev = _lib.require_cuda()
    x = np.random.default_rng(n).integers(0, 50, n).astype(np.int32)
    d = torch.as_tensor(x).to(dev)
    out = torch.empty_like(d)
    _lib.check(_lib.lib().fes_exclusive_scan_i32(_lib.ptr(d), _lib.ptr(out), n, _lib.stream()))
    ref = np.concatenate([[0], np.cumsum(x[:-1], dtype=np.int64)]).astype(np.int32)
    assert np.array_equal(out.cpu().numpy(), ref)
    _lib.check(_lib.lib().fes_exclusive_scan_i32(_lib.ptr(d), _lib.ptr(d), n, _lib.stream()))   # in place
    assert np.array_equal(d.cpu().numpy(), ref)
