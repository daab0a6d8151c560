#!/usr/bin/env python3
"""Builds tests/golden/golden.json + ocr.npz from the reference's own #[test] blocks.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The 20x20 Jaro and 22x22 Jaro-Winkler score matrices and the OCR byte arrays are parsed out of the
reference sources; the small known-answer asserts are transcribed below, each with its file:line.
Matrices are stored once ({names, cutoffs, scores}); the test expands them exactly as the reference's loop
does (expected = Some(score) iff cutoff <= score; distance tested at cutoff 1-c with expected 1-score).
Record format: {metric, kind, s1, s2, args{cutoff?, hint?, weights?, prefix_weight?}, expected|None, tol, src}.
Strings holding non-ASCII text are stored as code-point lists under s1_cp/s2_cp (u32 element tests).
"""
import json, os, re
import numpy as np

REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))
recs = []
matrices = {}

def add(metric, kind, s1, s2, expected, src, tol=0.0, **args):
    r = {"metric": metric, "kind": kind, "args": args, "expected": expected, "tol": tol, "src": src}
    for k, s in (("s1", s1), ("s2", s2)):
        if all(ord(c) < 128 for c in s):
            r[k] = s
        else:
            r[k + "_cp"] = [ord(c) for c in s]
    recs.append(r)

L = "distance/levenshtein.rs"
W112 = [1, 1, 2]
# empty / simple (:1934-1977)
add("levenshtein", "distance", "", "", 0, L + ":1935")
add("levenshtein", "distance", "aaaa", "", 4, L + ":1936")
for s2, d, ns in (("aaaa", 0, 1.0), ("aaa", 1, 0.75), ("aaab", 1, 0.75), ("bbbb", 4, 0.0)):
    add("levenshtein", "distance", "aaaa", s2, d, L + ":1942-1949")
    add("levenshtein", "normalized_similarity", "aaaa", s2, ns, L + ":1951-1976", tol=1e-4, cutoff=0.0)
add("levenshtein", "distance", "abaa", "baaa", 2, L + ":1947")
add("levenshtein", "normalized_similarity", "abaa", "baaa", 0.5, L + ":1969", tol=1e-4, cutoff=0.0)
# weighted_simple (:1981-2020)
for s1, s2, d, ns in (("aaaa", "aaaa", 0, 1.0), ("aaaa", "aaa", 1, 0.8571), ("abaa", "baaa", 2, 0.75),
                      ("aaaa", "aaab", 2, 0.75), ("aaaa", "bbbb", 8, 0.0)):
    add("levenshtein", "distance", s1, s2, d, L + ":1988-1992", weights=W112)
    add("levenshtein", "normalized_similarity", s1, s2, ns, L + ":1995-2019", tol=1e-4, weights=W112, cutoff=0.0)
# test_mbleven (:2024-2066)
a, b = "South Korea", "North Korea"
add("levenshtein", "distance", a, b, 2, L + ":2029")
for c, e in ((4, 2), (3, 2), (2, 2), (1, None), (0, None)):
    add("levenshtein", "distance", a, b, e, L + ":2030-2034", cutoff=c)
add("levenshtein", "distance", a, b, 4, L + ":2042", weights=W112)
for c, e in ((4, 4), (3, None), (2, None), (1, None)):
    add("levenshtein", "distance", a, b, e, L + ":2043-2046", weights=W112, cutoff=c)
a, b = "aabc", "cccd"
add("levenshtein", "distance", a, b, 4, L + ":2051")
for c, e in ((4, 4), (3, None), (2, None), (1, None), (0, None)):
    add("levenshtein", "distance", a, b, e, L + ":2052-2056", cutoff=c)
add("levenshtein", "distance", a, b, 6, L + ":2058", weights=W112)
for c, e in ((6, 6), (5, None), (4, None), (3, None), (2, None), (1, None), (0, None)):
    add("levenshtein", "distance", a, b, e, L + ":2059-2065", weights=W112, cutoff=c)
# ---- hamming.rs:549-641 (pad / error semantics: args pad=True, expected "error" = Err(DifferentLengthArgs))
H = "distance/hamming.rs"
add("hamming", "distance", "", "", 0, H + ":551")
add("hamming", "distance", "hamming", "hamming", 0, H + ":556")
add("hamming", "distance", "hamming", "hammers", 3, H + ":566")
add("hamming", "distance", "hammers", "hamming", 3, H + ":568-575", pad=True)
add("hamming", "distance", "hammers", "hamming", 3, H + ":576-583", pad=True, cutoff=3)
add("hamming", "distance", "hammers", "hamming", None, H + ":584-591", pad=True, cutoff=2)
add("hamming", "distance", "hammers", "hamming", 3, H + ":592-599", cutoff=3)
add("hamming", "distance", "hammers", "hamming", None, H + ":600-607", cutoff=2)
add("hamming", "distance", "hamming", "h\u9999mm\u00fcng", 2, H + ":612")
add("hamming", "distance", "ham", "hamming", "error", H + ":617-620")
add("hamming", "distance", "ham", "hamming", 4, H + ":622-625", pad=True)
add("hamming", "distance", "ham", "hamming", None, H + ":627-634", pad=True, cutoff=3)
add("hamming", "distance", "Friedrich Nietzs", "Jean-Paul Sartre", 14, H + ":639")
add("hamming", "distance", "hamming", "humming", 1, H + ":198")
# ---- damerau_levenshtein.rs:639-716
D = "distance/damerau_levenshtein.rs"
add("damerau_levenshtein", "distance", "", "", 0, D + ":641")
add("damerau_levenshtein", "distance", "aaaa", "", 4, D + ":642")
for s2, d, ns in (("aaaa", 0, 1.0), ("aaa", 1, 0.75), ("aaab", 1, 0.75), ("bbbb", 4, 0.0)):
    add("damerau_levenshtein", "distance", "aaaa", s2, d, D + ":648-655")
    add("damerau_levenshtein", "normalized_similarity", "aaaa", s2, ns, D + ":658-690", tol=1e-4, cutoff=0.0)
add("damerau_levenshtein", "distance", "abaa", "baaa", 1, D + ":651-654")
add("damerau_levenshtein", "normalized_similarity", "abaa", "baaa", 0.75, D + ":673-681", tol=1e-4, cutoff=0.0)
add("damerau_levenshtein", "distance", "CA", "ABC", 2, D + ":656,34,226,402")
add("damerau_levenshtein", "distance", "\u0418\u0432\u0430\u043d\u043a\u043e", "\u041f\u0435\u0442\u0440\u0443\u043d\u043a\u043e", 5, D + ":695-698")
add("damerau_levenshtein", "distance", "\u0418\u0432a\u043d\u043aoIvan", "\u041f\u0435\u0442\u0440\u0443\u043d\u043a\u043e", 10, D + ":700-703")
# ---- prefix.rs / postfix.rs doc-tests
add("prefix", "similarity", "prefix", "preference", 4, "distance/prefix.rs:122,256")
add("postfix", "similarity", "postfix", "prefix", 3, "distance/postfix.rs:122,256")
# test_banded (:2070-2130)
banded = [
    ("kkkkbbbbfkkkkkkibfkkkafakkfekgkkkkkkkkkkbdbbddddddddddafkkkekkkhkk",
     "khddddddddkkkkdgkdikkccccckcckkkekkkkdddddddddddafkkhckkkkkdckkkcc", 36, [(31, None)]),
    ("ccddcddddddddddddddddddddddddddddddddddddddddddddddddddddaaaaaaaaaaa",
     "aaaaaaaaaaaaaadddddddddbddddddddddddddddddddddddddddddddddbddddddddd", 26, [(31, 26)]),
    ("accccccccccaaaaaaaccccccccccccccccccccccccccccccacccccccccccccccccccccccccccccc"
     "ccccccccccccccccccccaaaaaaaaaaaaacccccccccccccccccccccc",
     "ccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccccc"
     "ccccccccccccccccccccccccccccccccccccbcccb", 24, [(25, 24)]),
    ("miiiiiiiiiiliiiiiiibghiiaaaaaaaaaaaaaaacccfccccedddaaaaaaaaaaaaaaaaaaaaaaaaaaaa"
     "aaaaaaaaaaaaa",
     "aaaaaaajaaaaaaaabghiiaaaaaaaaaaaaaaacccfccccedddaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa"
     "aajjdim", 27, [(27, 27)]),
    ("lllllfllllllllllllllllllllllllllllllllllllllllllllllllglllllilldcaaaaaaaaaaaaaa"
     "aaaaadbbllllllllllhllllllllllllllllllllllllllgl",
     "aaaaaaaaaaaaaadbbllllllllllllllelllllllllllllllllllllllllllllllglllllilldcaaaaa"
     "aaaaaaaaaaaaaadbbllllllllllllllellllllllllllllhlllllllllill", 23, [(27, 23), (28, 23)]),
    ("llccacaaaaaaaaaccccccccccccccccddffaccccaccecccggggclallhcccccljif",
     "bddcbllllllbcccccccccccccccccddffccccccccebcccggggclbllhcccccljifbddcccccc", 27, [(27, 27), (28, 27)]),
]
for s1, s2, d, cuts in banded:
    add("levenshtein", "distance", s1, s2, d, L + ":2070-2130")
    for c, e in cuts:
        add("levenshtein", "distance", s1, s2, e, L + ":2070-2130", cutoff=c)
add("levenshtein", "distance", "a" * 128, "b" * 128, 128, L + ":2133-2137")
add("levenshtein", "distance", "Иванко", "Петрунко", 5, L + ":2164-2169")
add("levenshtein", "distance", "CA", "ABC", 3, L + ":1378")
add("levenshtein", "distance", "kitten", "sitting", 3, "Readme.md:66-105")
add("levenshtein", "distance", "kitten", "sitting", None, "Readme.md:66-105", cutoff=2)

S = "distance/lcs_seq.rs"
add("lcs_seq", "distance", "a", "a", 0, S + ":1141")
add("lcs_seq", "distance", "aaaa", "aaaa", 0, S + ":1142")
add("lcs_seq", "similarity", "aaaa", "aaaa", 4, S + ":1143")
add("lcs_seq", "normalized_distance", "aaaa", "aaaa", 0.0, S + ":1144", tol=1e-4, cutoff=1.0)
add("lcs_seq", "normalized_similarity", "aaaa", "aaaa", 1.0, S + ":1149", tol=1e-4, cutoff=0.0)
add("lcs_seq", "distance", "aaaa", "bbbb", 4, S + ":1157")
add("lcs_seq", "similarity", "aaaa", "bbbb", 0, S + ":1158")
add("lcs_seq", "normalized_distance", "aaaa", "bbbb", 1.0, S + ":1159", tol=1e-4, cutoff=1.0)
add("lcs_seq", "normalized_similarity", "aaaa", "bbbb", 0.0, S + ":1164", tol=1e-4, cutoff=0.0)
a, b = "South Korea", "North Korea"
add("lcs_seq", "similarity", a, b, 9, S + ":1175")
add("lcs_seq", "similarity", a, b, 9, S + ":1176", cutoff=9)
add("lcs_seq", "similarity", a, b, None, S + ":1180", cutoff=10)
add("lcs_seq", "distance", a, b, 2, S + ":1185")
for c, e in ((4, 2), (3, 2), (2, 2), (1, None), (0, None)):
    add("lcs_seq", "distance", a, b, e, S + ":1186-1205", cutoff=c)
a, b = "aabc", "cccd"
add("lcs_seq", "similarity", a, b, 1, S + ":1209")
add("lcs_seq", "similarity", a, b, 1, S + ":1210", cutoff=1)
add("lcs_seq", "similarity", a, b, None, S + ":1214", cutoff=2)
add("lcs_seq", "distance", a, b, 3, S + ":1219")
for c, e in ((4, 3), (3, 3), (2, None), (1, None), (0, None)):
    add("lcs_seq", "distance", a, b, e, S + ":1220-1239", cutoff=c)
add("lcs_seq", "similarity", "001", "220", 1, S + ":1245-1249")
add("lcs_seq", "distance", "Иванко", "Петрунко", 5, S + ":1252-1257")
add("lcs_seq", "distance", "ab", "ac", 1, S + ":1260-1265")
add("lcs_seq", "distance", "lewenstein", "levenshtein", 2, S + ":581")
add("lcs_seq", "similarity", "lewenstein", "levenshtein", 9, S + ":763-764")

I = "distance/indel.rs"
add("indel", "distance", "aaaa", "aaaa", 0, I + ":712")
add("indel", "similarity", "aaaa", "aaaa", 8, I + ":713")
add("indel", "normalized_distance", "aaaa", "aaaa", 0.0, I + ":714", tol=1e-4, cutoff=1.0)
add("indel", "normalized_similarity", "aaaa", "aaaa", 1.0, I + ":719", tol=1e-4, cutoff=0.0)
add("indel", "distance", "aaaa", "bbbb", 8, I + ":728")
add("indel", "similarity", "aaaa", "bbbb", 0, I + ":729")
add("indel", "normalized_distance", "aaaa", "bbbb", 1.0, I + ":730", tol=1e-4, cutoff=1.0)
add("indel", "normalized_similarity", "aaaa", "bbbb", 0.0, I + ":735", tol=1e-4, cutoff=0.0)
a, b = "South Korea", "North Korea"
add("indel", "distance", a, b, 4, I + ":747")
for c, e in ((5, 4), (4, 4), (3, None), (2, None), (1, None), (0, None)):
    add("indel", "distance", a, b, e, I + ":748-771", cutoff=c)
a, b = "aabc", "cccd"
add("indel", "distance", a, b, 6, I + ":775")
for c, e in ((6, 6), (5, None), (4, None), (3, None), (2, None), (1, None), (0, None)):
    add("indel", "distance", a, b, e, I + ":776-803", cutoff=c)
add("indel", "normalized_similarity", "001", "220", 0.3333333, I + ":808-816", tol=1e-4, cutoff=0.0)
s1 = "ddccbccc"
s2 = ("aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa"
      "aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaacca"
      "cccaccaaaaaaaadaaaaaaaaccccaccccccaaaaaaaccccaaacccaccccadddaaaaaaaaaaaaaaaaa"
      "aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaccccccccacccaaaaaacccaaaaaacc"
      "cacccaaaaaacccdccccccaccccccccccccccccccccccccccccccccccccccccccccccccccccccc"
      "ccccccddddddaaaaaaaaaaaaaaaaaaaaaaaaaacacccaaaaaacccddddaaaaaaaaaaaaaaaaaaaaa"
      "aaaaaaaaccccaaaaaaaaaaccccccaadddaaaaaaaaaaaaaaaaaaaaaacaaaaaa")
add("indel", "distance", s1, s2, 508, I + ":819-848")
add("indel", "distance", s1, s2, 508, I + ":819-848", cutoff=508)
add("indel", "distance", s1, s2, None, I + ":819-848", cutoff=507)
add("indel", "distance", s1, s2, 508, I + ":819-848", cutoff=2**64 - 1)
s1 = ("bbbdbbmbbbbbbbbbBbfbbbbbbbbbbbbbbbbbbbrbbbbbrbbbbbdbnbbbjbhbbbbbbbbbhbbb"
      "bbCbobbbxbbbbbkbbbAbxbbwbbbtbcbbbbebbiblbbbbqbbbbbbpbbbbbbubbbkbbDbbbhbkbC"
      "bbgbbrbbbbbbbbbbbkbyvbbsbAbbbbz")
s2 = "jaaagaaqyaaaanrCfwaaxaeahtaaaCzaaaspaaBkvaaaaqDaacndaaeolwiaaauaaaaaaamA"
add("indel", "distance", s1, s2, 231, I + ":849-855")
add("indel", "distance", "Иванко", "Петрунко", 8, I + ":851-857")
add("indel", "distance", "ab", "ac", 2, I + ":859-864")
add("indel", "distance", "lewenstein", "levenshtein", 3, I + ":119")
add("indel", "distance", "lewenstein", "levenshtein", None, I + ":122", cutoff=2)

O = "distance/osa.rs"
add("osa", "distance", "", "", 0, O + ":672")
add("osa", "distance", "aaaa", "", 4, O + ":674")
add("osa", "distance", "aaaa", "", None, O + ":675", cutoff=1)
add("osa", "distance", "CA", "ABC", 3, O + ":677")
add("osa", "distance", "CA", "AC", 1, O + ":678")
filler = "a" * 64
add("osa", "distance", "a" + filler + "CA" + filler + "a", "b" + filler + "AC" + filler + "b", 3, O + ":680-683")
add("osa", "distance", "Иванко", "Петрунко", 5, O + ":686-691")

def parse_matrix(path, start_pat):
    src = open(os.path.join(REF, path)).read()
    i = src.index(start_pat)
    names_blk = re.search(r"let names = \[(.*?)\];", src[i:], re.S).group(1)
    names = re.findall(r'"([^"]*)"', names_blk)
    cut_blk = re.search(r"let score_cutoffs = \[(.*?)\];", src[i:], re.S).group(1)
    cutoffs = [float(x) for x in re.findall(r"[0-9.]+", cut_blk)]
    sc_blk = re.search(r"let scores = \[(.*?)\];", src[i:], re.S).group(1)
    scores = [float(x) for x in re.findall(r"[0-9.]+", sc_blk)]
    assert len(scores) == len(names) ** 2, (len(scores), len(names))
    return names, cutoffs, scores

J = "distance/jaro.rs"
add("jaro", "similarity", "james", "robert", 0.455556, J + ":1081-1086", tol=1e-4, cutoff=0.0)
add("jaro", "distance", "james", "robert", 1.0 - 0.455556, J + ":1087-1091", tol=1e-4, cutoff=1.0)
names, cutoffs, scores = parse_matrix(J, "fn test_flag_chars")
matrices["jaro"] = {"names": names, "cutoffs": cutoffs, "scores": scores, "tol": 1e-4, "src": J + ":1095-1189"}
add("jaro", "distance", "Иванко", "Петрунко", 0.375, J + ":1192-1199", tol=1e-4, cutoff=1.0)

JW = "distance/jaro_winkler.rs"
add("jaro_winkler", "similarity", "james", "robert", 0.455556, JW + ":680-685", tol=1e-4, cutoff=0.0)
add("jaro_winkler", "distance", "james", "robert", 1.0 - 0.455556, JW + ":686-690", tol=1e-4, cutoff=1.0)
names, cutoffs, scores = parse_matrix(JW, "fn test_flag_chars")
matrices["jaro_winkler"] = {"names": names, "cutoffs": cutoffs, "scores": scores, "tol": 1e-4, "src": JW + ":694-798"}
add("jaro_winkler", "distance", "Иванко", "Петрунко", 0.375, JW + ":801-808", tol=1e-4, cutoff=1.0)

F = "fuzz.rs"
S1, S3 = "new york mets", "the wonderful new york mets"
for s in (S1, "test", "mets", ""):
    add("ratio", "similarity", s, s, 1.0, F + ":182-214", tol=1e-4)
add("ratio", "similarity", S1, S3, 0.65, F + ":210", tol=1e-4)
add("ratio", "similarity", "test", "", 0.0, F + ":225", tol=1e-4)
add("ratio", "similarity", "", "test", 0.0, F + ":236", tol=1e-4)
for a, b in (("South Korea", "North Korea"), ("bc", "bca")):
    # issues 206/210 (:247-301): None at score+1e-4, Some(score) at score-1e-4; score = 2*LCS/(len1+len2)
    lcs = {("South Korea", "North Korea"): 9, ("bc", "bca"): 2}[(a, b)]
    score = 2.0 * lcs / (len(a) + len(b))
    add("ratio", "similarity", a, b, None, F + ":247-301", tol=1e-9, cutoff=score + 0.0001)
    add("ratio", "similarity", a, b, score, F + ":247-301", tol=1e-9, cutoff=score - 0.0001)

json.dump({"cases": recs, "matrices": matrices}, open(os.path.join(OUT, "golden.json"), "w"), ensure_ascii=True, indent=0)
print("wrote", len(recs), "records")

# OCR fixture (levenshtein.rs:2139-2161; data at distance/example/ocr.rs:2,5077)
src = open(os.path.join(REF, "distance/example/ocr.rs")).read()
arrs = re.findall(r"static (OCR_EXAMPLE\d)\s*: \[u8; (\d+)\] = \[(.*?)\];", src, re.S)
out = {}
for name, n, body in arrs:
    vals = np.array([int(x) for x in re.findall(r"\d+", body)], dtype=np.uint8)
    assert len(vals) == int(n), (name, len(vals), n)
    out[name] = vals
np.savez_compressed(os.path.join(OUT, "ocr.npz"), **out)
print({k: len(v) for k, v in out.items()})
