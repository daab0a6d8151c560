"""Mirror of rapidfuzz::fuzz (reference: src/fuzz.rs): ratio / ratio_with_args / RatioBatchComparator.

`ratio` is the Indel normalized similarity (fuzz.rs:60-85).  The reference's RatioBatchComparator literally
normalises by max(len1,len2) (fuzz.rs:141) although its documentation promises `ratio`;
the documented semantics are the default here, Args().reference_quirks() selects the literal behaviour."""
from ._scorer import Args, BatchComparatorBase


class RatioBatchComparator(BatchComparatorBase):
    METRIC = "ratio"

    def similarity(self, s2):
        return self._score("similarity", s2, None)

    def similarity_with_args(self, s2, args):
        return self._score("similarity", s2, args)


def ratio(s1, s2, args=None, device=0):
    b = RatioBatchComparator(s1, device)
    try:
        return b._score("similarity", s2, args)
    finally:
        b.close()


ratio_with_args = ratio
__all__ = ["Args", "RatioBatchComparator", "ratio", "ratio_with_args"]
