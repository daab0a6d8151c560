"""Shared machinery of the per-metric BatchComparator mirrors."""
import ctypes as C

import numpy as np

from . import _ffi
from .corpus import Corpus


class Args:
    """Mirror of the reference's per-metric `Args` builder (levenshtein.rs:86-126, jaro_winkler.rs:25-62,
    lcs_seq.rs / indel.rs / osa.rs / jaro.rs / fuzz.rs equivalents): Args().score_cutoff(x).score_hint(y)
    [.weights(ins, del, sub)] [.prefix_weight(w)].  With a score_cutoff the result carries `None`
    (masked) entries, exactly where the reference returns Option::None."""

    def __init__(self):
        self._cutoff = None
        self._hint = None
        self._weights = (1, 1, 1)
        self._prefix_weight = 0.1
        self._quirks = False
        self._pad = False

    def _copy(self):
        a = type(self)()
        a.__dict__.update(self.__dict__)
        return a

    def score_cutoff(self, v):
        a = self._copy()
        a._cutoff = v
        return a

    def score_hint(self, v):
        a = self._copy()
        a._hint = v
        return a

    def weights(self, insertion_cost=1, deletion_cost=1, substitution_cost=1):
        a = self._copy()
        a._weights = (insertion_cost, deletion_cost, substitution_cost)
        return a

    def prefix_weight(self, w):
        a = self._copy()
        a._prefix_weight = w
        return a

    def pad(self, pad=True):
        """hamming::Args::pad (hamming.rs:112-118)."""
        a = self._copy()
        a._pad = bool(pad)
        return a

    def reference_quirks(self, on=True):
        a = self._copy()
        a._quirks = on
        return a

    def _c(self, is_float):
        r = _ffi.RfArgs()
        _ffi.lib().rf_args_default(C.byref(r))
        r.insertion_cost, r.deletion_cost, r.substitution_cost = self._weights
        r.prefix_weight = self._prefix_weight
        r.reference_quirks = 1 if self._quirks else 0
        r.pad = 1 if self._pad else 0
        if self._cutoff is not None:
            r.has_cutoff = 1
            if is_float:
                r.cutoff_f = float(self._cutoff)
            else:
                r.cutoff_u = min(int(self._cutoff), 2**64 - 1)
        if self._hint is not None:
            r.has_hint = 1
            if is_float:
                r.hint_f = float(self._hint)
            else:
                r.hint_u = min(int(self._hint), 2**64 - 1)
        return r


def widen_elems(a):
    """Integer elements of any width -> what the C ABI takes: uint8 as is, everything else as uint32 holding the VALUE
    (the reference compares elements numerically -- HashableChar::hash_char, details/common.rs:29-37 -- so a signed -1
    must not meet an unsigned 255).  Negative values keep their two's-complement 32-bit pattern; the one ambiguity of
    that encoding (a negative i32 vs a u32 >= 2^31 with the same bits) and values outside [-2^31, 2^32) are rejected
    per array / not representable: 64-bit elements are accepted when every value fits."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint8 or a.dtype == np.uint32:
        return a
    if a.dtype == np.bool_:
        return a.astype(np.uint8)
    if not np.issubdtype(a.dtype, np.integer):
        raise TypeError("elements must be integers, bytes or str, not %s" % a.dtype)
    if a.size:
        lo, hi = int(a.min()), int(a.max())
        if lo < -(1 << 31) or hi >= (1 << 32):
            raise NotImplementedError("element values outside [-2^31, 2^32) (64-bit symbols) are not supported")
        if lo < 0 and hi >= (1 << 31):
            raise NotImplementedError("negative values and values >= 2^31 in one sequence are ambiguous as 32-bit symbols")
    return (a.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32)


def _as_query(q):
    if isinstance(q, str):
        if any(ord(ch) > 255 for ch in q):   # code points, like Rust's .chars(): u32-element comparator
            return np.fromiter((ord(ch) for ch in q), dtype=np.uint32, count=len(q))
        q = q.encode("latin-1")
    if isinstance(q, (bytes, bytearray)):
        return np.frombuffer(bytes(q), dtype=np.uint8)
    return widen_elems(q)


class BatchComparatorBase:
    """`BatchComparator::new(query)` of one metric module: caches the query and its pattern-match bit
    table on the GPU; every scoring method takes a Corpus (one-vs-many in one launch) or a single
    candidate (bytes/str, returns a scalar like the reference)."""
    METRIC = None

    def __init__(self, query, device=0):
        q = _as_query(query)
        h = C.c_void_p()
        create = _ffi.lib().rf_batch_create_u32 if q.dtype == np.uint32 else _ffi.lib().rf_batch_create_u8
        _ffi.check(create(_ffi.METRICS[self.METRIC], q.ctypes.data, len(q), device, C.byref(h)))
        self._h = h
        self.device = device
        self.query = q

    @classmethod
    def from_typed(cls, query, device=0):
        """rf_batch_create_elems: an integer query of any width handed to the C ABI as it is."""
        q = np.ascontiguousarray(query)
        if q.dtype.name not in _ffi.ELEM_TYPES:
            raise TypeError("unsupported element type %s" % q.dtype)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_batch_create_elems(_ffi.METRICS[cls.METRIC], q.ctypes.data, _ffi.ELEM_TYPES[q.dtype.name], len(q),
                                                   device, C.byref(h)))
        self._h = h
        self.device = device
        self.query = np.zeros(0, np.uint32)   # marks the comparator as wide for _score
        return self

    def stream_elems32(self, kind, elems, offsets, args=None):
        """rf_batch_stream_*_elems32: host-resident candidates with u32 elements (needs a u32 / typed comparator)."""
        args = args if args is not None else Args()
        elems = np.ascontiguousarray(elems, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        out = np.empty(n, dtype=np.float64 if is_f else np.uint32)
        fn = _ffi.lib().rf_batch_stream_f64_elems32 if is_f else _ffi.lib().rf_batch_stream_u32_elems32
        _ffi.check(fn(self._h, elems.ctypes.data, offsets.ctypes.data, n, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "_wide_twin", None) is not None:
            self._wide_twin.close()
            self._wide_twin = None
        if getattr(self, "_h", None):
            _ffi.lib().rf_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _score(self, kind, s2, args):
        args = args if args is not None else Args()
        single = not isinstance(s2, Corpus)
        if single:
            wide = (isinstance(s2, str) and any(ord(ch) > 255 for ch in s2)) or (isinstance(s2, np.ndarray) and s2.dtype == np.uint32)
            if wide and self.query.dtype != np.uint32:   # a byte query still has to meet the candidate's symbols
                other = type(self)(self.query.astype(np.uint32), self.device)
                try:
                    return other._score(kind, s2, args)
                finally:
                    other.close()
            if isinstance(s2, np.ndarray) and s2.dtype == np.uint32:
                corpus = Corpus.from_u32(s2, np.array([0, len(s2)], dtype=np.uint64), self.device)
            else:
                corpus = Corpus.from_unicode([s2], self.device) if wide else Corpus.from_strings([s2], self.device)
        else:
            corpus = s2
            if getattr(corpus, "wide", False) and self.query.dtype != np.uint32:   # byte query, u32-element corpus
                if getattr(self, "_wide_twin", None) is None:
                    self._wide_twin = type(self)(self.query.astype(np.uint32), self.device)
                return self._wide_twin._score(kind, corpus, args)
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        n = len(corpus)
        out = np.empty(n, dtype=np.float64 if is_f else np.uint32)
        fn = _ffi.lib().rf_batch_score_f64 if is_f else _ffi.lib().rf_batch_score_u32
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        has_cutoff = args._cutoff is not None
        if single:
            v = out[0]
            if is_f:
                return None if np.isnan(v) else float(v)
            return None if (has_cutoff and v == _ffi.NONE_U32) else int(v)
        if not has_cutoff:
            return out
        mask = np.isnan(out) if is_f else (out == _ffi.NONE_U32)
        return np.ma.MaskedArray(out, mask=mask)

    def extract(self, kind, corpus, k=5, args=None):
        """rf_batch_extract_*: the k best candidates of `corpus` by (score best-first, index ascending), selected on
        the GPU (Python rapidfuzz's process.extract).  Returns (idx uint32[m], score[m]) with m <= k."""
        args = args if args is not None else Args()
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        idx = np.empty(k, dtype=np.uint32)
        score = np.empty(k, dtype=np.float64 if is_f else np.uint32)
        m = C.c_uint32(0)
        fn = _ffi.lib().rf_batch_extract_f64 if is_f else _ffi.lib().rf_batch_extract_u32
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), k, idx.ctypes.data, score.ctypes.data, C.byref(m)))
        return idx[: m.value], score[: m.value]

    def filter(self, kind, corpus, args, capacity=None):
        """rf_batch_filter_*: every candidate whose score passed args.score_cutoff, in index order, compacted on the
        GPU.  Returns (idx uint32[m], score[m], total_hits); m = min(total_hits, capacity)."""
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        cap = len(corpus) if capacity is None else int(capacity)
        idx = np.empty(cap, dtype=np.uint32)
        score = np.empty(cap, dtype=np.float64 if is_f else np.uint32)
        tot = C.c_uint64(0)
        fn = _ffi.lib().rf_batch_filter_f64 if is_f else _ffi.lib().rf_batch_filter_u32
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), cap, idx.ctypes.data, score.ctypes.data, C.byref(tot)))
        m = min(tot.value, cap)
        return idx[:m], score[:m], tot.value

    def stream(self, kind, chars, offsets, args=None, out=None):
        """rf_batch_stream_*: scores HOST-resident candidates (CSR chars u8 + offsets u32/u64) without keeping a
        corpus on the GPU; chunked H2D / scan / D2H pipeline.  Returns the raw sentinel-carrying array
        (0xFFFFFFFF / NaN == None).  Pinned buffers (e.g. torch .pin_memory()) run at full PCIe speed."""
        args = args if args is not None else Args()
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets)
        if offsets.dtype not in (np.uint32, np.uint64):
            offsets = offsets.astype(np.uint64)
        n = len(offsets) - 1
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        if out is None:
            out = np.empty(n, dtype=np.float64 if is_f else np.uint32)
        assert out.dtype == (np.float64 if is_f else np.uint32) and len(out) >= n and out.flags.c_contiguous
        name = "rf_batch_stream_%s%s" % ("f64" if is_f else "u32", "_off32" if offsets.dtype == np.uint32 else "")
        _ffi.check(getattr(_ffi.lib(), name)(self._h, chars.ctypes.data, offsets.ctypes.data, n, _ffi.KINDS[kind],
                                             C.byref(ca), out.ctypes.data))
        return out

    def stream_len8_packed6(self, kind, packed, dict64, lens, args=None, out=None, u8_results=False):
        """rf_batch_stream_{u32,u8}_len8_packed6: like stream_len8 with the characters 6-bit packed by corpus.pack6."""
        args = args if args is not None else Args()
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        dict64 = np.ascontiguousarray(dict64, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint8)
        n = len(lens)
        ca = args._c(False)
        dt = np.uint8 if u8_results else np.uint32
        if out is None:
            out = np.empty(n, dtype=dt)
        assert out.dtype == dt and len(out) >= n and out.flags.c_contiguous and len(dict64) == 64
        fn = _ffi.lib().rf_batch_stream_u8_len8_packed6 if u8_results else _ffi.lib().rf_batch_stream_u32_len8_packed6
        _ffi.check(fn(self._h, packed.ctypes.data, dict64.ctypes.data, lens.ctypes.data, n, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def stream_len8(self, kind, chars, lens, args=None, out=None, u8_results=False):
        """rf_batch_stream_{u32,u8}_len8: host-resident candidates described by one LENGTH byte each (<= 255 elements),
        integer-valued kinds; with u8_results the scores come back as bytes (0xFF == None, scores must be <= 254)."""
        args = args if args is not None else Args()
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint8)
        n = len(lens)
        ca = args._c(False)
        dt = np.uint8 if u8_results else np.uint32
        if out is None:
            out = np.empty(n, dtype=dt)
        assert out.dtype == dt and len(out) >= n and out.flags.c_contiguous
        fn = _ffi.lib().rf_batch_stream_u8_len8 if u8_results else _ffi.lib().rf_batch_stream_u32_len8
        _ffi.check(fn(self._h, chars.ctypes.data, lens.ctypes.data, n, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def score_u8(self, kind, corpus, args=None):
        """rf_batch_score_u8: integer scores of a resident corpus as bytes (0xFF = None); raises when a score exceeds 254."""
        args = args if args is not None else Args()
        ca = args._c(False)
        out = np.empty(len(corpus), dtype=np.uint8)
        _ffi.check(_ffi.lib().rf_batch_score_u8(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def score_into(self, kind, corpus, out_ptr, args=None, stream=0):
        """Device-pointer variant (rf_batch_score_*_device): results stay on the GPU at `out_ptr`
        (u32[n] or f64[n]); enqueued on `stream` (a cudaStream_t as int), no synchronisation."""
        args = args if args is not None else Args()
        is_f = bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.METRIC], _ffi.KINDS[kind]))
        ca = args._c(is_f)
        fn = _ffi.lib().rf_batch_score_f64_device if is_f else _ffi.lib().rf_batch_score_u32_device
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), out_ptr, stream))

    # --- the reference's method set (e.g. levenshtein.rs:1660-1817)
    def distance(self, s2):
        return self._score("distance", s2, None)

    def distance_with_args(self, s2, args):
        return self._score("distance", s2, args)

    def similarity(self, s2):
        return self._score("similarity", s2, None)

    def similarity_with_args(self, s2, args):
        return self._score("similarity", s2, args)

    def normalized_distance(self, s2):
        return self._score("normalized_distance", s2, None)

    def normalized_distance_with_args(self, s2, args):
        return self._score("normalized_distance", s2, args)

    def normalized_similarity(self, s2):
        return self._score("normalized_similarity", s2, None)

    def normalized_similarity_with_args(self, s2, args):
        return self._score("normalized_similarity", s2, args)


def make_module(metric):
    cls = type("BatchComparator", (BatchComparatorBase,), {"METRIC": metric, "__doc__": BatchComparatorBase.__doc__})

    def _free(kind):
        def f(s1, s2, args=None, device=0):
            b = cls(s1, device)
            try:
                return b._score(kind, s2, args)
            finally:
                b.close()
        f.__name__ = kind
        return f
    return cls, {k: _free(k) for k in _ffi.KINDS}
