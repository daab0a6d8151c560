"""rapidfuzz::distance (reference: src/distance.rs:1-10) -- the metric modules on the GPU hot path."""
from . import damerau_levenshtein, hamming, indel, jaro, jaro_winkler, lcs_seq, levenshtein, osa, postfix, prefix  # noqa: F401
