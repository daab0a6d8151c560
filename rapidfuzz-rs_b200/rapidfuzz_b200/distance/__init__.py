"""rapidfuzz::distance (reference: src/distance.rs:1-10) -- the metric modules on the GPU hot path."""
from . import indel, jaro, jaro_winkler, lcs_seq, levenshtein, osa  # noqa: F401
