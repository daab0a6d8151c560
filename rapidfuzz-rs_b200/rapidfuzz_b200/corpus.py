"""Packed candidate corpus resident in one GPU's HBM (CSR: chars u8[total] + offsets[n+1]).

New on this side: the reference consumes one candidate iterator per call
(levenshtein.rs:1750-1762); here the candidates are uploaded once and scored by whole-corpus kernels."""
import ctypes as C
import os

import numpy as np

from . import _ffi


class Corpus:
    def __init__(self, chars, offsets, device=0):
        """chars: uint8 array/bytes of all candidates back to back; offsets: n+1 CSR starts (u32 or u64)."""
        chars = np.ascontiguousarray(np.frombuffer(chars, dtype=np.uint8) if isinstance(chars, (bytes, bytearray)) else chars,
                                     dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets)
        if offsets.dtype not in (np.uint32, np.uint64):
            offsets = offsets.astype(np.uint64)
        n = len(offsets) - 1
        if n < 0:
            raise ValueError("offsets needs n+1 entries")
        h = C.c_void_p()
        fn = _ffi.lib().rf_corpus_create_u8 if offsets.dtype == np.uint64 else _ffi.lib().rf_corpus_create_u8_off32
        _ffi.check(fn(chars.ctypes.data, offsets.ctypes.data, n, device, C.byref(h)))
        self._h = h
        self.device = device

    @classmethod
    def from_u32(cls, elems, offsets, device=0):
        """rf_corpus_create_u32: candidates with u32 elements (e.g. Unicode code points); needs comparators created
        from u32 / non-latin-1 queries."""
        elems = np.ascontiguousarray(elems, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_create_u32(elems.ctypes.data, offsets.ctypes.data, len(offsets) - 1, device, C.byref(h)))
        self._h = h
        self.device = device
        self.wide = True     # needs comparators created from u32 queries (BatchComparatorBase widens a byte query itself)
        return self

    @classmethod
    def from_elems(cls, elems, offsets, device=0):
        """Candidates with integer elements of any width (u8 ... i32, 64-bit when the values fit): widened to the C ABI's
        u8 / u32 by VALUE (see _scorer.widen_elems), so signed and unsigned sequences compare like in the reference."""
        from ._scorer import widen_elems
        e = widen_elems(elems)
        return cls(e, offsets, device) if e.dtype == np.uint8 else cls.from_u32(e, offsets, device)

    @classmethod
    def from_typed(cls, elems, offsets, device=0):
        """rf_corpus_create_elems: integer elements of any width handed to the C ABI AS THEY ARE (u8 ... u64, i8 ... i64);
        the library widens them by value (what from_elems does on the Python side)."""
        elems = np.ascontiguousarray(elems)
        if elems.dtype.name not in _ffi.ELEM_TYPES:
            raise TypeError("unsupported element type %s" % elems.dtype)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_create_elems(elems.ctypes.data, _ffi.ELEM_TYPES[elems.dtype.name], offsets.ctypes.data,
                                                    len(offsets) - 1, device, C.byref(h)))
        self._h = h
        self.device = device
        self.wide = elems.dtype != np.uint8
        return self

    @classmethod
    def from_unicode(cls, strings, device=0):
        """Python str candidates as sequences of code points (what Rust's `.chars()` yields)."""
        cps = [np.fromiter((ord(ch) for ch in s), dtype=np.uint32, count=len(s)) for s in strings]
        offsets = np.zeros(len(cps) + 1, dtype=np.uint64)
        if cps:
            offsets[1:] = np.cumsum([len(c) for c in cps])
        elems = np.concatenate(cps + [np.zeros(0, np.uint32)]).astype(np.uint32)
        return cls.from_u32(elems, offsets, device)

    @classmethod
    def from_strings(cls, strings, device=0):
        bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in strings]
        offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            offsets[1:] = np.cumsum([len(b) for b in bs])
        return cls(np.frombuffer(b"".join(bs), dtype=np.uint8), offsets, device)

    @classmethod
    def from_device(cls, chars_ptr, offsets_ptr, n, total_chars, device=0, stream=0):
        """Adopt (copy) buffers already on `device`: raw device pointers, e.g. torch tensor .data_ptr();
        offsets are u64."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_create_device_u8(chars_ptr, offsets_ptr, n, total_chars, device, stream, C.byref(h)))
        self._h = h
        self.device = device
        return self

    @classmethod
    def from_file(cls, path, device=0):
        """rf_corpus_create_from_file: map a corpus file (write_corpus_file) and upload it."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_create_from_file(os.fsencode(path), device, C.byref(h)))
        self._h = h
        self.device = device
        return self

    def __len__(self):
        return int(_ffi.lib().rf_corpus_size(self._h))

    @property
    def total_chars(self):
        return int(_ffi.lib().rf_corpus_total_chars(self._h))

    def release_csr(self):
        """rf_corpus_release_csr: frees the CSR copy (45 % of the footprint); only what the interleaved layout serves works afterwards."""
        _ffi.check(_ffi.lib().rf_corpus_release_csr(self._h))

    @property
    def has_csr(self):
        return bool(_ffi.lib().rf_corpus_has_csr(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _ffi.lib().rf_corpus_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_strings(strings, nthreads=0):
    """rf_pack_u8: list of bytes/str -> (chars u8, offsets u64), copied in parallel by the C library."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in strings]
    n = len(bs)
    ptrs = (C.c_char_p * max(n, 1))(*bs) if n else (C.c_char_p * 1)()
    lens = np.array([len(b) for b in bs], dtype=np.uint64)
    offsets = np.empty(n + 1, dtype=np.uint64)
    l = _ffi.lib()
    _ffi.check(l.rf_pack_u8(C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, offsets.ctypes.data, None, nthreads))
    chars = np.empty(int(offsets[n]), dtype=np.uint8)
    _ffi.check(l.rf_pack_u8(C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, offsets.ctypes.data, chars.ctypes.data, nthreads))
    return chars, offsets


def pack6(chars, nthreads=0, pinned=False):
    """rf_pack6_u8: concatenated byte candidates with at most 64 distinct symbols -> (packed u8, dict u8[64]); 4 characters in
    3 bytes, for the *_packed6 streaming entry points."""
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    l = _ffi.lib()
    size = int(l.rf_pack6_size(len(chars)))
    if pinned:
        import torch
        packed_t = torch.empty(size, dtype=torch.uint8).pin_memory()
        packed = packed_t.numpy()
    else:
        packed = np.empty(size, dtype=np.uint8)
    d = np.zeros(64, dtype=np.uint8)
    _ffi.check(l.rf_pack6_u8(chars.ctypes.data, len(chars), packed.ctypes.data, d.ctypes.data, nthreads))
    return packed, d


def write_corpus_file(path, chars, offsets):
    """rf_corpus_file_write: CSR corpus -> a file that maps back without parsing (see include/rfgpu.h)."""
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    _ffi.check(_ffi.lib().rf_corpus_file_write(os.fsencode(path), chars.ctypes.data, offsets.ctypes.data, len(offsets) - 1))


class CorpusFile:
    """rf_corpus_file_open: read-only mmap of a corpus file; .chars / .offsets are numpy views into the mapping
    (valid until close()), usable with Corpus(...) and BatchComparator.stream(...)."""

    def __init__(self, path):
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_file_open(os.fsencode(path), C.byref(h)))
        self._h = h
        l = _ffi.lib()
        self.n = int(l.rf_corpus_file_size(h))
        self.total = int(l.rf_corpus_file_total_chars(h))
        w = int(l.rf_corpus_file_offset_width(h))
        odt = np.uint32 if w == 4 else np.uint64
        self.offsets = np.ctypeslib.as_array(C.cast(l.rf_corpus_file_offsets(h), C.POINTER(C.c_uint32 if w == 4 else C.c_uint64)),
                                             shape=(self.n + 1,)).view(odt)
        if self.total:
            self.chars = np.ctypeslib.as_array(C.cast(l.rf_corpus_file_chars(h), C.POINTER(C.c_uint8)), shape=(self.total,))
        else:
            self.chars = np.zeros(0, dtype=np.uint8)

    def close(self):
        if getattr(self, "_h", None):
            self.offsets = self.chars = None
            _ffi.lib().rf_corpus_file_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def cdist_topk_u32(q_elems, q_offsets, corpus, k=10, score_cutoff=None):
    """rf_cdist_topk_u32: u32-element queries (CSR) against a corpus made by Corpus.from_u32 / from_unicode."""
    q_elems = np.ascontiguousarray(q_elems, dtype=np.uint32)
    q_off = np.ascontiguousarray(q_offsets, dtype=np.uint64)
    nq = len(q_off) - 1
    a = _ffi.RfArgs()
    _ffi.lib().rf_args_default(C.byref(a))
    if score_cutoff is not None:
        a.has_cutoff = 1
        a.cutoff_u = int(score_cutoff)
    idx = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.uint32)
    _ffi.check(_ffi.lib().rf_cdist_topk_u32(q_elems.ctypes.data, q_off.ctypes.data, nq, corpus._h, C.byref(a), k,
                                           idx.ctypes.data, dist.ctypes.data))
    return idx, dist


def cdist_topk(queries, corpus, k=10, score_cutoff=None, metric="levenshtein"):
    """Many-vs-many: for every query the k best candidates of `corpus` by (distance, index); metric = levenshtein (default),
    osa, indel or lcs_seq (rf_cdist_topk_metric_u8).
    New on this side (the reference has no cdist).  queries: list of bytes/str, or (chars u8, offsets u64).
    Returns (idx [nq,k] uint32, dist [nq,k] uint32); rows with fewer than k hits are padded with 0xFFFFFFFF."""
    import ctypes as C
    if isinstance(queries, tuple):
        q_chars = np.ascontiguousarray(queries[0], dtype=np.uint8)
        q_off = np.ascontiguousarray(queries[1], dtype=np.uint64)
    else:
        bs = [q.encode("latin-1") if isinstance(q, str) else bytes(q) for q in queries]
        q_off = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            q_off[1:] = np.cumsum([len(b) for b in bs])
        q_chars = np.frombuffer(b"".join(bs), dtype=np.uint8)
    nq = len(q_off) - 1
    a = _ffi.RfArgs()
    _ffi.lib().rf_args_default(C.byref(a))
    if score_cutoff is not None:
        a.has_cutoff = 1
        a.cutoff_u = int(score_cutoff)
    idx = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.uint32)
    if metric == "levenshtein":
        _ffi.check(_ffi.lib().rf_cdist_topk_u8(q_chars.ctypes.data, q_off.ctypes.data, nq, corpus._h, C.byref(a), k,
                                              idx.ctypes.data, dist.ctypes.data))
    else:
        _ffi.check(_ffi.lib().rf_cdist_topk_metric_u8(_ffi.METRICS[metric], q_chars.ctypes.data, q_off.ctypes.data, nq, corpus._h,
                                                     C.byref(a), k, idx.ctypes.data, dist.ctypes.data))
    return idx, dist
