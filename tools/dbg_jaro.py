"""Dev helper: Jaro similarity of short queries against mixed-length groups (incl. padding lanes) vs the oracle, with a
per-length breakdown of the mismatches -- found the wrapped-radius bug of the segment bounds (run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rapidfuzz_b200 as rf
from oracle import oracle as orc
from gpu_util import gpu_batch
rng = np.random.default_rng(7)
for qlen in (1, 2, 5, 32):
    q = rng.integers(1, 5, qlen).astype(np.uint8)
    lens = rng.choice([1, 3, 8, 31, 32, 33, 50, 64, 65, 100], 3000)
    chars = rng.integers(1, 5, int(lens.sum())).astype(np.uint8)
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    c = rf.Corpus(chars, off)
    got = gpu_batch("jaro", "similarity", q, c)
    exp = orc.batch("jaro", "similarity", q, chars, off, nthreads=0)
    bad = np.nonzero(got != exp)[0]
    print("qlen", qlen, "mismatches", len(bad), "by len2:", {int(l): int(np.sum(lens[bad] == l)) for l in np.unique(lens[bad])})
    for i in bad[:4]:
        print("   idx", i, "len2", lens[i], "got", got[i], "exp", exp[i])
    c.close()
