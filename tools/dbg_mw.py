import sys, faulthandler, numpy as np
faulthandler.dump_traceback_later(25, exit=True)
sys.path.insert(0,'.'); sys.path.insert(0,'rapidfuzz-rs_b200'); sys.path.insert(0,'tests')
import rapidfuzz_b200 as rf
from gpu_util import make_corpus, gpu_batch
from oracle import oracle as orc
rng = np.random.default_rng(365)
q = (rng.integers(0, 4, 65) + 97).astype(np.uint8)
for n, lens in ((40,[0,1,5,63,64,65,127]), (600,[0, 1, 5, 63, 64, 65, 127, 128, 129, 200, 256, 257, 64, 65, 66, 105])):
    chars, offsets = make_corpus(rng, n, lens, alphabet=4, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for metric, kw in (("levenshtein", {}), ("levenshtein", {"cutoff": 4}), ("indel", {}), ("osa", {})):
        print("run", n, metric, kw, flush=True)
        got = gpu_batch(metric, "distance", q, corpus, **kw)
        exp = orc.batch(metric, "distance", q, chars, offsets, **kw)
        bad = np.nonzero(got != exp)[0]
        print("   ok" if len(bad) == 0 else "   MISMATCH %s" % [(int(i), int(offsets[i+1]-offsets[i]), int(got[i]), int(exp[i])) for i in bad[:12]], flush=True)
