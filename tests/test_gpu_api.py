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


@pytest.mark.parametrize("rc", [False, True])
@pytest.mark.parametrize("k", [7, 25, 40])
def test_batch_files_reload_gives_the_direct_result(tmp_path, monkeypatch, k, rc):
    """`kmer batch` -> `kmer count/uniq -B` (kmermaid/batcher.py:616-636, with its inverted guard fixed)
    must equal the direct run and the oracle: the batch files are formatted on the GPU, re-imported as
    device keys, and the uniq output keeps the headers the files stored.  Also through a gzipped copy."""
    import gzip

    import kmer_oracle as ko
    import numpy as np

    monkeypatch.setenv("KMG_ALPHABET", "IUPAC")
    rng = np.random.default_rng(500 + k)
    s1 = "".join(rng.choice(list("ACGT"), size=5000))
    s2 = "".join(rng.choice(list("ACGT"), size=3000))
    recs = [("chrA first", s1 + s1[:800].lower() + "NNNNNNNNNNRY" + s1[100:400]), ("chrB", s2 + "N" + s2[:500] + "ACGTN")]
    fa = tmp_path / "in.fa"
    fa.write_bytes(ko.synth_fasta_bytes([(t, s.encode()) for t, s in recs]))
    extra = ["-r"] if rc else []
    outdir = tmp_path / "batches"
    _run_cli(["batch", *extra, str(fa), str(outdir), str(k)])
    files = sorted(os.listdir(outdir))
    assert files
    # the files hold every k-mer record, sorted by sequence, in the reference's record format
    lines = b"".join((outdir / f).read_bytes() for f in files).split(b"\n")[:-1]
    assert len(lines) % 2 == 0 and all(h.startswith(b">") for h in lines[0::2])
    seqs = lines[1::2]
    assert seqs == sorted(seqs) and all(len(x) == k for x in seqs)
    # one batch (fewer records than the batch size): byte for byte the reference's stably sorted batch file
    (one,) = ko.batches_py(recs, k, rc)
    assert list(zip(lines[0::2], seqs)) == [(b">" + h.encode(), q.encode()) for h, q in one]
    gz = tmp_path / "batches_gz"
    gz.mkdir()
    for f in files:
        with gzip.open(gz / (f + ".gz"), "wb") as oh:
            oh.write((outdir / f).read_bytes())
    for cmd, want in (("count", ko.count_text_np(recs, k, rc)), ("uniq", ko.uniq_text_np(recs, k, rc))):
        for src in (outdir, gz):
            out = tmp_path / f"{cmd}_{src.name}.txt"
            _run_cli([cmd, "-B", str(src), str(fa), str(out), str(k)])
            assert out.read_bytes() == want, (cmd, src.name)


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


def test_gpu_fasta_loader_matches_host_loader(golden, tmp_path):
    """kmg_fasta_flatten (three kernels) against kman_b200.fasta (numpy) and the oracle's parser."""
    import gzip

    import kmer_oracle as ko
    import numpy as np
    from kman_b200 import fasta
    from kman_b200.engine import get_engine

    eng = get_engine()
    texts = {c["name"]: c["fasta_text"] for c in golden["cases"]}
    rng = np.random.default_rng(3)
    # long single-line records, many tiny records, headers at tile boundaries, CRLF, preamble, empty records
    big = ">long one\n" + ko.synth_bases(150_000, 1).decode() + "\n>" + "x" * 5000 + " y\n" + "\n".join(
        ko.synth_bases(61, i).decode() for i in range(300)) + "\n"
    texts["long_lines"] = big
    texts["many_records"] = "".join(">r%d d\nAC GT\n\nNNAC\n" % i for i in range(3000))
    texts["crlf_big"] = "junk\r\n;c\r\n" + "".join(">q%d\r\n%s\r\n" % (i, ko.synth_bases(100 + i, i).decode()) for i in range(200))
    texts["empty_records"] = ">a\n>b\nACGT\n>c\n>d\n"
    texts["no_final_newline"] = ">a\nACGT\n>b x y\nGG"
    pad = "".join(rng.choice(list("ACGT\n"), p=[.24, .24, .24, .24, .04], size=4096 * 3 + 7))
    for off in (4080, 4095, 4096, 4097):
        texts["boundary_%d" % off] = ">h\n" + pad[: off - 3] + "\n>next header here\nACGTACGT\n"
    for name, text in texts.items():
        p = tmp_path / (name + ".fa")
        p.write_bytes(text.encode("latin-1"))
        want = fasta.parse_bytes(text.encode("latin-1"))
        d = eng.load_fasta(str(p))
        got = d.flat
        assert got.titles == want.titles, name
        assert got.names == want.names, name
        assert (got.rec_starts == want.rec_starts).all(), name
        assert got.bases.tobytes() == want.bases.tobytes(), name
        assert d.n_bases == want.bases.size
        assert d.bases[: d.n_bases].cpu().numpy().tobytes() == want.bases.tobytes(), name
    # gz, tabs (host fallback), missing / empty files
    p = tmp_path / "t.fa.gz"
    with gzip.open(p, "wt") as fh:
        fh.write(">r\nACGT\nAC\n")
    assert eng.load_fasta(str(p)).flat.bases.tobytes() == b"ACGTAC\n"
    p = tmp_path / "tab.fa"
    p.write_text(">a\tb c\nACG\tT \t\nAC GT\n")
    assert eng.load_fasta(str(p)).flat.bases.tobytes() == fasta.parse_bytes(p.read_bytes()).bases.tobytes()
    with pytest.raises(AssertionError):
        eng.load_fasta(str(tmp_path / "missing.fa"))
    (tmp_path / "nohdr.fa").write_text("no header here\nACGT\n")
    with pytest.raises(AssertionError):
        eng.load_fasta(str(tmp_path / "nohdr.fa"))


def test_abundance_vectors_match_reference_goldens(tmp_path, monkeypatch):
    """`kmer count -m VEC_COUNT / VEC_COUNT_MASKED` through the command line: every REF___STRAND.gz file
    of every golden case (reference outputs, tests/golden/golden_vec.json)."""
    import gzip
    import json

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_vec.json")))
    n = 0
    for c in gold["cases"]:
        monkeypatch.setenv("KMG_ALPHABET", c["alphabet"])
        fa = tmp_path / f"in_{n}.fa"
        fa.write_bytes(c["fasta_text"].encode("latin-1"))
        out = tmp_path / f"vec_{n}.txt"
        _run_cli(["count", "-m", c["mode"], *(["-r"] if c["rc"] else []), str(fa), str(out), str(c["k"])])
        d = tmp_path / f"vec_{n}"
        got = {f: gzip.open(d / f, "rb").read().decode("latin-1") for f in sorted(os.listdir(d))} if d.is_dir() else {}
        assert got == c["files"], (c["name"], c["k"], c["alphabet"], c["rc"], c["mode"])
        n += 1
    assert n >= 130


@pytest.mark.parametrize("k,rc", [(31, False), (21, True), (45, False)])
def test_abundance_vectors_large_input_vs_oracle(k, rc):
    """> 2^20 k-mers with duplicated stretches shared between records, N runs (groups of 10^4 equal wide
    k-mers), soft-masked bases: device vectors against the oracle's numpy tier, both modes."""
    import kmer_oracle as ko
    import numpy as np

    from kman_b200 import abundance, fasta
    from kman_b200.engine import get_engine

    rng = np.random.default_rng(77 + k)
    b = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=1_300_000, dtype=np.uint8)].copy()
    b[600_000:700_000] = b[100_000:200_000]      # shared between records 1 and 2
    b[900_000:950_000] = b[880_000:930_000]      # repeated inside record 3
    b[300_000:312_000] = ord("N")
    b[1_000_000:1_020_000] |= 0x20
    recs = [("r1 a", b[:500_000].tobytes().decode()), ("r2", b[500_000:800_000].tobytes().decode()),
            ("r3", b[800_000:].tobytes().decode())]
    eng = get_engine()
    d = eng.upload(fasta.from_records(recs))
    for masked in (False, True):
        vecs = abundance.device_vectors(eng, d, k, rc, masked)
        got = {"%s___%s.gz" % (r, s): abundance.vector_text(v, k) for r, per in vecs.items() for s, v in per.items()}
        assert got == ko.vec_count_np(recs, k, rc, masked), (k, rc, masked)


@pytest.mark.parametrize("k", [5, 12, 31, 33, 64])
@pytest.mark.parametrize("rc", [False, True])
def test_count_windows_equals_the_extraction_counts(k, rc):
    """Engine.count_windows (DeviceBatch.current_size: the 4-mer-histogram pre-pass for k >= 12) against
    the counts a full extraction reports, on records with N runs, IUPAC symbols, foreign bytes and
    records shorter than k."""
    import numpy as np

    from kman_b200 import fasta
    from kman_b200.engine import get_engine

    eng = get_engine(0)
    rng = np.random.default_rng(k * 2 + rc)
    recs = []
    for r, n in enumerate((50_000, 3, 0, 70, 20_000)):
        s = "".join(rng.choice(list("ACGTacgtNRX"), p=[.24, .24, .24, .24, .01, .01, .005, .005, .004, .003, .003], size=n))
        recs.append(("r%d" % r, s))
    for alphabet in ("IUPAC", "ACGT"):
        d = eng.upload(fasta.from_records(recs), alphabet=alphabet)
        a = eng.extract(d, k, rc, wide=False, val_bytes=0)
        assert eng.count_windows(d, k, rc) == (a.n, a.n_other), (alphabet, k, rc)
