"""CPU-side checks of the packing / corpus-file layer (csrc/rf_io.cpp): the step before the scoring path."""
import os

import numpy as np
import pytest

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi


def test_pack_strings_matches_python_join():
    rng = np.random.default_rng(0)
    strings = [bytes(rng.integers(0, 256, int(l), dtype=np.uint8)) for l in rng.integers(0, 40, 5000)]
    strings += [b"", b"x", "héllo"]
    chars, offsets = rf.pack_strings(strings, nthreads=4)
    bs = [s.encode("latin-1") if isinstance(s, str) else s for s in strings]
    assert bytes(chars) == b"".join(bs)
    assert np.array_equal(offsets, np.concatenate([[0], np.cumsum([len(b) for b in bs])]).astype(np.uint64))
    c0, o0 = rf.pack_strings([])
    assert len(c0) == 0 and o0.tolist() == [0]


@pytest.mark.parametrize("n", [0, 1, 1000])
def test_corpus_file_round_trip(tmp_path, n):
    q = synth.synth_query(1, 32)
    chars, offsets = synth.synth_corpus(1, q, n, 0, 64, 16)
    path = str(tmp_path / "c.rfc")
    rf.write_corpus_file(path, chars, offsets)
    assert os.path.getsize(path) % 1 == 0 and os.path.getsize(path) >= 64
    with rf.CorpusFile(path) as f:
        assert f.n == n and f.total == len(chars)
        assert f.offsets.dtype == np.uint32            # small corpora store 32-bit offsets
        assert np.array_equal(f.offsets.astype(np.uint64), offsets)
        assert np.array_equal(np.asarray(f.chars), chars)
        assert f.offsets.ctypes.data % 64 == 0 and (len(chars) == 0 or f.chars.ctypes.data % 64 == 0)


def test_corpus_file_rejects_garbage(tmp_path):
    p = tmp_path / "bad.rfc"
    p.write_bytes(b"not a corpus file" * 10)
    with pytest.raises(rf.RfError) as ei:
        rf.CorpusFile(str(p))
    assert ei.value.status == _ffi.RF_ERR_INVALID_ARG
    with pytest.raises(rf.RfError):
        rf.CorpusFile(str(tmp_path / "missing.rfc"))
    # truncated file
    q = synth.synth_query(1, 8)
    chars, offsets = synth.synth_corpus(1, q, 100, 1, 20, 4)
    good = tmp_path / "good.rfc"
    rf.write_corpus_file(str(good), chars, offsets)
    (tmp_path / "cut.rfc").write_bytes(good.read_bytes()[:-5])
    with pytest.raises(rf.RfError):
        rf.CorpusFile(str(tmp_path / "cut.rfc"))
    # non-monotone offsets are refused at write time
    bad = offsets.copy()
    bad[5] = bad[6] + 1
    with pytest.raises(rf.RfError):
        rf.write_corpus_file(str(tmp_path / "x.rfc"), chars, bad)
    # ... and a file whose index was corrupted afterwards is refused at open time (rf_io.cpp:157)
    raw = bytearray(good.read_bytes())
    width = 4 if len(chars) < 2**32 - 16 else 8
    pos = 64 + 5 * width                      # offsets[5], behind the 64-byte header
    raw[pos:pos + width] = (int(offsets[7]) + 3).to_bytes(width, "little")
    (tmp_path / "corrupt.rfc").write_bytes(bytes(raw))
    with pytest.raises(rf.RfError) as ei:
        rf.CorpusFile(str(tmp_path / "corrupt.rfc"))
    assert ei.value.status == _ffi.RF_ERR_INVALID_ARG


def test_pack6_round_trip_on_the_host():
    """rf_pack6_u8: 4 characters -> 3 bytes, dictionary in ascending symbol order; more than 64 symbols are refused."""
    rng = np.random.default_rng(4)
    alpha = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    for total in (0, 1, 3, 4, 5, 63, 64, 65, 100_003):
        chars = alpha[rng.integers(0, len(alpha), total)]
        packed, d = rf.pack6(chars, nthreads=3)
        assert len(packed) == (total + 3) // 4 * 3 + 64
        used = np.unique(chars)
        assert np.array_equal(d[: len(used)], used) and not d[len(used):].any()
        bits = np.unpackbits(packed[: (total + 3) // 4 * 3], bitorder="little")
        codes = bits[: total * 6].reshape(total, 6).dot(1 << np.arange(6)).astype(np.uint8) if total else np.zeros(0, np.uint8)
        assert np.array_equal(d[codes], chars)
    with pytest.raises(rf.RfError) as ei:
        rf.pack6(np.arange(65, dtype=np.uint8))
    assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
    rf.pack6(np.arange(64, dtype=np.uint8) + 100)    # exactly 64 symbols fit
