"""Pins the CPU oracle (oracle/kmer_oracle.py) against

  * every golden vector under tests/golden/ -- produced by running the UNMODIFIED
    reference (oracle/gen_golden.py), both alphabets, with and without -r;
  * the reference's own known-answer tests (tests/test_seq.py:125-181,
    tests/test_batch.py:80-87 in /root/reference), restated here as data.
"""
import hashlib
import os

import numpy as np
import pytest

import kmer_oracle as ko


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def test_golden_edge_cases_both_tiers(golden):
    assert len(golden["cases"]) >= 100
    for c in golden["cases"]:
        recs = ko.parse_fasta_text(c["fasta_text"])
        for cmd, f_py, f_np in (("count", ko.count_text_py, ko.count_text_np), ("uniq", ko.uniq_text_py, ko.uniq_text_np)):
            want = c[cmd].encode("latin-1")
            tag = (c["name"], c["k"], c["alphabet"], c["rc"], cmd)
            assert f_py(recs, c["k"], c["rc"], alphabet=c["alphabet"]) == want, tag
            assert f_np(recs, c["k"], c["rc"], alphabet=c["alphabet"]) == want, tag


def test_golden_batch_files(golden):
    n = 0
    for c in golden["cases"]:
        if "batch_b10" not in c:
            continue
        recs = ko.parse_fasta_text(c["fasta_text"])
        got = sorted(ko.batch_text_py(recs, c["k"], c["rc"], batch_size=10, alphabet=c["alphabet"]))
        assert got == sorted(x.encode("latin-1") for x in c["batch_b10"])
        n += 1
    assert n >= 4


def test_batch_size_independence(golden):
    c = [c for c in golden["cases"] if c["name"] == "tiny" and c["k"] == 4 and c["alphabet"] == "IUPAC" and not c["rc"]][0]
    recs = ko.parse_fasta_text(c["fasta_text"])
    for b in (1, 3, 7, 1000):
        assert ko.count_text_py(recs, 4, batch_size=b).decode("latin-1") == c["count"]
        assert ko.uniq_text_py(recs, 4, batch_size=b).decode("latin-1") == c["uniq"]


@pytest.mark.parametrize("name", ["syn_100k", "syn_1m"])
def test_golden_synthetic_hashes(golden, name):
    ent = [e for e in golden["synthetic"] if e["name"] == name][0]
    data = ko.synth_fasta_bytes([("chr1 synthetic seed=%d" % ent["seed"], ko.synth_bases(ent["n"], ent["seed"]))])
    assert _sha(data) == ent["fasta_sha256"]
    recs = ko.parse_fasta_text(data.decode())
    out = ko.count_text_np(recs, ent["k"])
    assert (out.count(b"\n"), _sha(out)) == (ent["count_lines"], ent["count_sha256"])
    out = ko.uniq_text_np(recs, ent["k"])
    assert (out.count(b"\n"), _sha(out)) == (ent["uniq_lines"], ent["uniq_sha256"])
    if name == "syn_100k":
        assert _sha(ko.count_text_py(recs, ent["k"])) == ent["count_sha256"]
        assert _sha(ko.uniq_text_py(recs, ent["k"])) == ent["uniq_sha256"]


def test_golden_syn_dup(golden):
    ents = [e for e in golden["synthetic"] if e["name"] == "syn_dup"]
    assert len(ents) == 8
    for ent in ents:
        recs = ko.parse_fasta_text(ent["fasta_text"])
        for cmd, f_py, f_np in (("count", ko.count_text_py, ko.count_text_np), ("uniq", ko.uniq_text_py, ko.uniq_text_np)):
            for f in (f_py, f_np):
                out = f(recs, ent["k"], ent["rc"])
                assert (out.count(b"\n"), _sha(out)) == (ent[cmd + "_lines"], ent[cmd + "_sha256"]), (ent["k"], ent["rc"], cmd)


# ---- the reference's own known-answer tests, restated as data ---------------------------------
def test_ref_kat_kmers_of_ACGAT():  # /root/reference/tests/test_seq.py:125-130
    got = list(ko.kmers_py("ACGAT", 4, "stest"))
    assert got == [("stest", 0, 4, "+", "ACGA"), ("stest", 1, 5, "+", "CGAT")]


def test_ref_kat_batcher_overlap():  # tests/test_seq.py:136-138
    assert list(ko.batcher_py("ACGATCGATCG", 3, 5)) == [("ACGAT", 0), ("ATCGA", 3), ("GATCG", 6)]


def test_ref_kat_batched_coords_and_rc():  # tests/test_seq.py:140-181
    seq = "ACGATCGATCG"
    chunks = list(ko.batcher_py(seq, 4, 5))
    got = [[(s, e, st, km) for _, s, e, st, km in ko.kmers_py(c, 4, "ref", off, rc=True)] for c, off in chunks]
    want = [
        [(0, 4, "+", "ACGA"), (0, 4, "-", "TCGT"), (1, 5, "+", "CGAT"), (1, 5, "-", "ATCG")],
        [(2, 6, "+", "GATC"), (2, 6, "-", "GATC"), (3, 7, "+", "ATCG"), (3, 7, "-", "CGAT")],
        [(4, 8, "+", "TCGA"), (4, 8, "-", "TCGA"), (5, 9, "+", "CGAT"), (5, 9, "-", "ATCG")],
        [(6, 10, "+", "GATC"), (6, 10, "-", "GATC"), (7, 11, "+", "ATCG"), (7, 11, "-", "CGAT")],
    ]
    assert got == want


def test_ref_kat_header_format():  # tests/test_seq.py:11-54,110-112 ; seq.py:103-104
    assert ko.header_py("chr1", 0, 1000, "+") == "chr1:0-1000:+"
    assert ko.header_py("chr1", 5, 9, "-") == "chr1:5-9:-"


def test_ref_kat_sorted_batch():  # tests/test_batch.py:80-87: Python's sorted() on the records
    recs = [("h%d" % i, s) for i, s in enumerate(["TTTT", "ACGT", "ACGA", "GGGG", "ACGT"])]
    assert [r[1] for r in sorted(recs, key=lambda r: r[1])] == ["ACGA", "ACGT", "ACGT", "GGGG", "TTTT"]
    # stability: equal sequences keep emission order
    assert [r[0] for r in sorted(recs, key=lambda r: r[1]) if r[1] == "ACGT"] == ["h1", "h4"]


def test_np_extract_matches_py_random():
    rng = np.random.default_rng(5)
    for trial in range(20):
        n_rec = int(rng.integers(1, 4))
        recs = []
        for r in range(n_rec):
            n = int(rng.integers(0, 200))
            s = "".join(rng.choice(list("ACGTacgtNRYXn-"), p=[.2, .2, .2, .2, .03, .03, .03, .03, .02, .01, .01, .02, .01, .01], size=n))
            recs.append(("r%d x" % r, s))
        for k in (2, 5, 16, 17, 31, 32, 33, 40, 64):
            for rc in (False, True):
                for ab in ("IUPAC", "ACGT"):
                    assert ko.count_text_np(recs, k, rc, ab) == ko.count_text_py(recs, k, rc, alphabet=ab), (trial, k, rc, ab)
                    assert ko.uniq_text_np(recs, k, rc, ab) == ko.uniq_text_py(recs, k, rc, alphabet=ab), (trial, k, rc, ab)


def test_threaded_count_table_equals_the_single_threaded_tier():
    """count_table_np_threads (bench.py's reference arm: chunks with k-1 overlap sorted by a thread each,
    merge + grouping split by key range) gives count_table_np's table for any chunk size and thread count."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        recs = []
        for r in range(int(rng.integers(1, 5))):
            n = int(rng.integers(0, 3000))
            recs.append(("r%d" % r, "".join(rng.choice(list("ACGTacgtN"), p=[.23, .23, .23, .23, .02, .02, .02, .01, .01], size=n))))
        for k in (2, 9, 31, 32, 33, 63):
            for rc in (False, True):
                want = ko.count_table_np(recs, k, rc, "ACGT")
                for threads, bs in ((1, None), (3, None), (4, 64 + k), (7, 1000)):
                    got = ko.count_table_np_threads(recs, k, rc, "ACGT", threads=threads, batch_size=bs)
                    assert len(got[0]) == len(want[0])
                    assert all(np.array_equal(a, b) for a, b in zip(got[0], want[0])), (trial, k, rc, threads, bs)
                    assert np.array_equal(got[1], want[1]), (trial, k, rc, threads, bs)


def test_k_must_exceed_one():  # batcher.py:477-478
    with pytest.raises(AssertionError):
        ko.count_text_py([("a", "ACGT")], 1)
    with pytest.raises(AssertionError):
        ko.count_text_np([("a", "ACGT")], 1)


# ---- large-configuration table hashes (tests/golden/table_hashes.json) -------------------------------
def test_table_hash_digests_follow_the_text_tier():
    """The digests oracle/gen_table_hashes.py commits are computed on packed keys; on a small
    config-4-shaped input (N runs, soft-masked bases, IUPAC symbols, three records) the same
    keys, decoded, must be exactly the text the literal pure-Python tier writes."""
    import gen_table_hashes as gth
    import synth_configs as sc

    recs = sc.cfg4(30_000)
    for k, rc, ab in ((25, False, "IUPAC"), (25, True, "ACGT"), (45, False, "IUPAC")):
        ex = ko.extract_np(recs, k, rc, ab)
        d_count, d_uniq = gth.digests(recs, k, rc, ab, "count"), gth.digests(recs, k, rc, ab, "uniq")
        txt = ko.count_text_py(recs, k, rc, alphabet=ab)
        assert txt == ko.count_text_np(recs, k, rc, ab)
        assert txt.count(b"\n") == d_count["narrow"]["rows"] + d_count["wide"]["rows"]
        assert d_count["narrow"]["total"] + d_count["wide"]["total"] == len(ex["narrow"]["pos"]) + len(ex["wide"]["pos"])
        utxt = ko.uniq_text_py(recs, k, rc, alphabet=ab)
        assert utxt.count(b"\n") == 2 * (d_uniq["narrow"]["rows"] + d_uniq["wide"]["rows"])
        # the count digest is the hash of exactly the table count_np groups
        _, _, det = ko.count_np(recs, k, rc, ab)
        assert d_count["narrow"] == {**sc.count_digest(sc.key_rows(det["narrow"]["keys"]), det["narrow"]["counts"]),
                                     "keys_in": len(ex["narrow"]["pos"])}


def test_committed_table_hashes_reproduce():
    """One committed case is recomputed here (about 20 s): the file is what the generator writes."""
    import json

    import gen_table_hashes as gth

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "table_hashes.json")))
    assert set(gold) == set(gth.CASES)
    name = "cfg3_24mbp_24rec_k31_count"
    make, k, rc, ab, mode = gth.CASES[name]
    got = gth.digests(make(), k, rc, ab, mode)
    for key, val in got.items():
        assert gold[name][key] == val, key


# ---- abundance vectors (VEC_COUNT / VEC_COUNT_MASKED) ------------------------------------------------
def test_vec_count_tiers_match_reference_goldens():
    """tests/golden/golden_vec.json: the reference's own outputs (oracle/gen_golden_vec.py: one abstract
    method neutralised at run time, kmermaid/abundance.py:123 -> :60); both oracle tiers must reproduce
    every file of every case (5 edge FASTA texts x k 3/4/7 x both alphabets x +-rc x both modes, plus a
    duplicated multi-record input at k 11/31/45)."""
    import json

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_vec.json")))
    assert len(gold["cases"]) >= 130
    for c in gold["cases"]:
        assert "error" not in c["files"], c
        recs = ko.parse_fasta_text(c["fasta_text"])
        want = {f: t.encode("latin-1") for f, t in c["files"].items()}
        masked = c["mode"] == "VEC_COUNT_MASKED"
        assert ko.vec_count_py(recs, c["k"], c["rc"], masked, c["alphabet"]) == want, (c["name"], c["k"], c["mode"])
        assert ko.vec_count_np(recs, c["k"], c["rc"], masked, c["alphabet"]) == want, (c["name"], c["k"], c["mode"])
