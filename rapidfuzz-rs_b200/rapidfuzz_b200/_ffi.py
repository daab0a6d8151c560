"""ctypes view of the C ABI in include/rfgpu.h (librfgpu.so, built in-tree by rapidfuzz-rs_b200/build.py).

There is no CPU fallback: if the shared library is missing it is built; if it cannot be loaded, or no CUDA
device is usable, the compute entry points raise RfError."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("RF_LIB_PATH") or os.path.join(_ROOT, "lib", "librfgpu.so")  # RF_LIB_PATH: dev A/B builds

RF_OK, RF_ERR_INVALID_ARG, RF_ERR_UNSUPPORTED, RF_ERR_CUDA, RF_ERR_OOM, RF_ERR_NCCL = range(6)
METRICS = {"levenshtein": 0, "indel": 1, "lcs_seq": 2, "osa": 3, "jaro": 4, "jaro_winkler": 5, "ratio": 6,
           "hamming": 7, "prefix": 8, "postfix": 9, "damerau_levenshtein": 10}
KINDS = {"distance": 0, "similarity": 1, "normalized_distance": 2, "normalized_similarity": 3}
ELEM_TYPES = {"uint8": 0, "uint16": 1, "uint32": 2, "uint64": 3, "int8": 4, "int16": 5, "int32": 6, "int64": 7}
NONE_U32 = 0xFFFFFFFF
RF_MAX_QUERY_LEN = 4194304


class RfArgs(C.Structure):
    _fields_ = [("has_cutoff", C.c_uint8), ("cutoff_u", C.c_uint64), ("cutoff_f", C.c_double),
                ("has_hint", C.c_uint8), ("hint_u", C.c_uint64), ("hint_f", C.c_double),
                ("insertion_cost", C.c_uint64), ("deletion_cost", C.c_uint64), ("substitution_cost", C.c_uint64),
                ("prefix_weight", C.c_double), ("reference_quirks", C.c_uint8), ("pad", C.c_uint8)]


class RfError(RuntimeError):
    def __init__(self, status, detail):
        super().__init__("rfgpu status %d: %s" % (status, detail))
        self.status = status


# every symbol include/rfgpu.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_PA = C.POINTER(RfArgs)
SYMBOLS = {
    "rf_args_default": (None, [_PA]),
    "rf_status_string": (C.c_char_p, [_int]),
    "rf_last_error": (C.c_char_p, []),
    "rf_device_count": (_int, []),
    "rf_corpus_create_u8": (_int, [_vp, _vp, _u64, _int, C.POINTER(_vp)]),
    "rf_corpus_create_u8_off32": (_int, [_vp, _vp, _u64, _int, C.POINTER(_vp)]),
    "rf_corpus_create_device_u8": (_int, [_vp, _vp, _u64, _u64, _int, _vp, C.POINTER(_vp)]),
    "rf_corpus_create_u32": (_int, [_vp, _vp, _u64, _int, C.POINTER(_vp)]),
    "rf_corpus_create_elems": (_int, [_vp, _int, _vp, _u64, _int, C.POINTER(_vp)]),
    "rf_corpus_destroy": (_int, [_vp]),
    "rf_corpus_release_csr": (_int, [_vp]),
    "rf_corpus_has_csr": (_int, [_vp]),
    "rf_corpus_size": (_u64, [_vp]),
    "rf_corpus_total_chars": (_u64, [_vp]),
    "rf_corpus_device": (_int, [_vp]),
    "rf_batch_create_u8": (_int, [_int, _vp, _u32, _int, C.POINTER(_vp)]),
    "rf_batch_create_u32": (_int, [_int, _vp, _u32, _int, C.POINTER(_vp)]),
    "rf_batch_create_elems": (_int, [_int, _vp, _int, _u32, _int, C.POINTER(_vp)]),
    "rf_batch_destroy": (_int, [_vp]),
    "rf_batch_set_option": (_int, [_vp, C.c_char_p, _int]),
    "rf_batch_score_u32": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_batch_score_u8": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_batch_score_u8_device": (_int, [_vp, _vp, _int, _PA, _vp, _vp]),
    "rf_batch_score_f64": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_batch_score_u32_device": (_int, [_vp, _vp, _int, _PA, _vp, _vp]),
    "rf_batch_score_f64_device": (_int, [_vp, _vp, _int, _PA, _vp, _vp]),
    "rf_result_is_float": (_int, [_int, _int]),
    "rf_batch_distance_u32": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_similarity_u32": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_distance_f64": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_similarity_f64": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_normalized_distance_f64": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_normalized_similarity_f64": (_int, [_vp, _vp, _PA, _vp]),
    "rf_batch_extract_u32": (_int, [_vp, _vp, _int, _PA, _u32, _vp, _vp, _vp]),
    "rf_batch_extract_f64": (_int, [_vp, _vp, _int, _PA, _u32, _vp, _vp, _vp]),
    "rf_batch_filter_u32": (_int, [_vp, _vp, _int, _PA, _u64, _vp, _vp, _vp]),
    "rf_batch_filter_f64": (_int, [_vp, _vp, _int, _PA, _u64, _vp, _vp, _vp]),
    "rf_batch_stream_u32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_u32_off32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_f64": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_f64_off32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_u32_elems32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_f64_elems32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_u32_len8": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_u8_len8": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_pack6_size": (_u64, [_u64]),
    "rf_pack6_u8": (_int, [_vp, _u64, _vp, _vp, _int]),
    "rf_batch_stream_u32_len8_packed6": (_int, [_vp, _vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_batch_stream_u8_len8_packed6": (_int, [_vp, _vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_cdist_topk_u8": (_int, [_vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp]),
    "rf_cdist_topk_u8_device": (_int, [_vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp, _vp]),
    "rf_cdist_topk_metric_u8": (_int, [_int, _vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp]),
    "rf_cdist_topk_metric_u8_device": (_int, [_int, _vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp, _vp]),
    "rf_cdist_topk_u32": (_int, [_vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp]),
    "rf_cdist_topk_u32_device": (_int, [_vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp, _vp]),
    "rf_topk_merge_device": (_int, [_vp, _vp, _u64, _vp, _u32, _u32, _u32, _vp, _vp, _int, _vp]),
    "rf_corpus_create_sharded_u8": (_int, [_vp, _vp, _u64, _vp, _int, C.POINTER(_vp)]),
    "rf_sharded_corpus_destroy": (_int, [_vp]),
    "rf_sharded_corpus_size": (_u64, [_vp]),
    "rf_sharded_corpus_shards": (_int, [_vp]),
    "rf_sharded_corpus_shard_range": (_int, [_vp, _int, C.POINTER(_u64), C.POINTER(_u64)]),
    "rf_sharded_corpus_shard": (_vp, [_vp, _int]),
    "rf_sharded_corpus_uses_nccl": (_int, [_vp]),
    "rf_sharded_batch_create_u8": (_int, [_int, _vp, _u32, _vp, _int, C.POINTER(_vp)]),
    "rf_sharded_batch_create_u32": (_int, [_int, _vp, _u32, _vp, _int, C.POINTER(_vp)]),
    "rf_sharded_batch_destroy": (_int, [_vp]),
    "rf_sharded_score_u32": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_sharded_score_f64": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_sharded_score_u32_allgather_device": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_sharded_score_f64_allgather_device": (_int, [_vp, _vp, _int, _PA, _vp]),
    "rf_sharded_extract_u32": (_int, [_vp, _vp, _int, _PA, _u32, _vp, _vp, _vp]),
    "rf_sharded_extract_f64": (_int, [_vp, _vp, _int, _PA, _u32, _vp, _vp, _vp]),
    "rf_sharded_cdist_topk_u8": (_int, [_vp, _vp, _u32, _vp, _PA, _u32, _vp, _vp]),
    "rf_sharded_stream_u32": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_sharded_stream_f64": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_sharded_stream_u32_len8": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_sharded_stream_u8_len8": (_int, [_vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_sharded_stream_u8_len8_packed6": (_int, [_vp, _vp, _vp, _vp, _u64, _int, _PA, _vp]),
    "rf_comm_unique_id": (_int, [_vp]),
    "rf_comm_create_rank": (_int, [_vp, _int, _int, _int, C.POINTER(_vp)]),
    "rf_comm_destroy": (_int, [_vp]),
    "rf_comm_rank": (_int, [_vp]),
    "rf_comm_size": (_int, [_vp]),
    "rf_batch_score_u32_allgather_device": (_int, [_vp, _vp, _vp, _int, _PA, _vp, _u64, _vp, _vp]),
    "rf_batch_score_f64_allgather_device": (_int, [_vp, _vp, _vp, _int, _PA, _vp, _u64, _vp, _vp]),
    "rf_pack_u8": (_int, [_vp, _vp, _u64, _vp, _vp, _int]),
    "rf_corpus_file_write": (_int, [C.c_char_p, _vp, _vp, _u64]),
    "rf_corpus_file_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "rf_corpus_file_close": (_int, [_vp]),
    "rf_corpus_file_size": (_u64, [_vp]),
    "rf_corpus_file_total_chars": (_u64, [_vp]),
    "rf_corpus_file_offset_width": (_u32, [_vp]),
    "rf_corpus_file_offsets": (_vp, [_vp]),
    "rf_corpus_file_chars": (_vp, [_vp]),
    "rf_corpus_create_from_file": (_int, [C.c_char_p, _int, C.POINTER(_vp)]),
    "rf_kernel_launch_count": (_u64, []),
    "rf_set_option": (_int, [C.c_char_p, _int]),
}

_lib = None


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_rf_build", os.path.join(_ROOT, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def _preload_nccl():
    """librfgpu.so needs libnccl.so.2 (sharded entry points).  PyTorch bundles its own copy under nvidia/nccl/lib; if that
    one is present, map it first (RTLD_GLOBAL) so that this library and a later `import torch` share ONE NCCL instead of
    the system copy shadowing torch's.  Without the bundled copy the loader picks the system libnccl.so.2."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            p = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(p):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _preload_nccl()
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != RF_OK:
        raise RfError(status, lib().rf_last_error().decode("utf-8", "replace"))
