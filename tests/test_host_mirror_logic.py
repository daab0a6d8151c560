"""Host-side logic of the Python mirror that needs no GPU: element widening by value (HashableChar semantics,
/root/reference/src/details/common.rs:29-37) and the byte-balanced shard ranges."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))
from rapidfuzz_b200._scorer import widen_elems, _as_query
from rapidfuzz_b200 import sharding


def test_widen_elems_keeps_values_apart():
    assert widen_elems(np.array([1, 2], np.uint8)).dtype == np.uint8
    assert widen_elems(np.array([1, 70000], np.uint32)).dtype == np.uint32
    u16 = widen_elems(np.array([65535, 7], np.uint16))
    i16 = widen_elems(np.array([-1, 7], np.int16))
    assert u16.dtype == i16.dtype == np.uint32
    assert u16[0] == 65535 and i16[0] == 0xFFFFFFFF and u16[1] == i16[1] == 7      # same value <=> same symbol
    assert widen_elems(np.array([-1], np.int8))[0] != widen_elems(np.array([255], np.uint8))[0]
    assert np.array_equal(widen_elems(np.array([5, -3], np.int64)), np.array([5, 0xFFFFFFFD], np.uint32))
    assert widen_elems(np.array([True, False])).tolist() == [1, 0]
    assert widen_elems(np.zeros(0, np.int16)).dtype == np.uint32
    with pytest.raises(NotImplementedError):
        widen_elems(np.array([1 << 32], np.uint64))
    with pytest.raises(NotImplementedError):
        widen_elems(np.array([-1, 1 << 31], np.int64))                              # ambiguous as 32-bit symbols
    with pytest.raises(TypeError):
        widen_elems(np.array([1.5]))


def test_as_query_forms():
    assert _as_query("abc").dtype == np.uint8 and _as_query("abc").tolist() == [97, 98, 99]
    assert _as_query("aé").dtype == np.uint8                                        # latin-1 stays bytes
    assert _as_query("aЖ").dtype == np.uint32 and _as_query("aЖ").tolist() == [97, 0x416]
    assert _as_query(b"\x00\xff").tolist() == [0, 255]
    assert _as_query(np.array([300], np.uint16)).tolist() == [300]                  # no silent truncation to a byte


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 200, 10_000)
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    for world in (1, 2, 3, 8):
        bounds = [sharding.shard_range(off, world, r) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == len(lens)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        sizes = [int(off[hi] - off[lo]) for lo, hi in bounds]
        assert max(sizes) - min(sizes) <= 2 * 200                                   # balanced by bytes, not by count
    # empty corpus: ranges split by count
    assert [sharding.shard_range(np.zeros(5, np.uint64), 2, r) for r in range(2)] == [(0, 2), (2, 4)]
