"""Mirror of rapidfuzz::distance::damerau_levenshtein (reference: src/distance/damerau_levenshtein.rs): BatchComparator + Args and the free functions,
executed by the CUDA kernels behind the C ABI."""
from .._scorer import Args, make_module

BatchComparator, _free = make_module("damerau_levenshtein")
distance = _free["distance"]
similarity = _free["similarity"]
normalized_distance = _free["normalized_distance"]
normalized_similarity = _free["normalized_similarity"]
distance_with_args = distance
similarity_with_args = similarity
normalized_distance_with_args = normalized_distance
normalized_similarity_with_args = normalized_similarity
__all__ = ["Args", "BatchComparator", "distance", "similarity", "normalized_distance", "normalized_similarity"]
