"""rapidfuzz_b200 -- host-side mirror of the rapidfuzz-rs `distance::*::BatchComparator` / `fuzz::*` API
over the B200 CUDA engine (librfgpu.so, C ABI in include/rfgpu.h).  No CPU fallback."""
from . import _ffi
from ._scorer import Args
from .corpus import Corpus, CorpusFile, cdist_topk, cdist_topk_u32, pack6, pack_strings, write_corpus_file
from . import distance, fuzz

RfError = _ffi.RfError
__all__ = ["Args", "Corpus", "CorpusFile", "RfError", "cdist_topk", "cdist_topk_u32", "pack_strings", "pack6", "write_corpus_file", "distance", "fuzz"]
