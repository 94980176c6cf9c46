"""Host-side API mirror (kman_b200.seq / batch / batcher / join) -- the parts of the reference's
own unit tests that need no extraction, restated against the drop-in
(/root/reference/tests/test_seq.py:11-114,184-201 and tests/test_batch.py:11-184)."""
import os

import pytest

from kman_b200.alphabet import NATYPES
from kman_b200.batch import Batch, BatchAppendable
from kman_b200.batcher import BatcherThreading, FastaBatcher, load_batches
from kman_b200.join import Crawler, KJoiner, KJoinerThreading
from kman_b200.seq import KMer, Sequence, SequenceCoords, SequenceCount


def test_sequence_coords():
    for bad in ((-1, 5), (1, -5)):
        with pytest.raises(AssertionError):
            SequenceCoords("chr1", *bad)
    with pytest.raises(AssertionError):
        SequenceCoords("chr1", 1, 5, "+")
    c = SequenceCoords("chr1", 0, 1000)
    assert (c.ref, c.start, c.end, c.strand) == ("chr1", 0, 1000, SequenceCoords.STRAND.PLUS)
    assert str(c) == "chr1:0-1000:+"
    m = SequenceCoords("chr:1 x", 3, 7, SequenceCoords.STRAND.MINUS)
    assert str(m) == "chr:1 x:3-7:-" and SequenceCoords.from_str(str(m)) == m
    assert SequenceCoords.rev(SequenceCoords.STRAND.PLUS) == SequenceCoords.STRAND.MINUS
    assert SequenceCoords.STRAND.MINUS.label == "-"
    with pytest.raises(AssertionError):
        SequenceCoords.from_str("chr1:a-b:+")


def test_kmer_record():
    with pytest.raises(AssertionError):
        KMer("chr1", 0, 5, "ACGT")  # length mismatch (seq.py:449-450)
    with pytest.raises(AssertionError):
        Sequence("ACGATCGATCG", "DNA")
    k = KMer("chr1", 10, 14, "acga")
    assert k.seq == "ACGA" and k.header == "chr1:10-14:+"
    assert k == KMer.from_fasta((k.header, k.seq))
    assert k.as_fasta() == ">chr1:10-14:+\nACGA\n" and str(k) == "chr1:10-14:+\tACGA"
    assert k.is_ab_checked()
    assert Sequence("ACGATCGATCG", NATYPES.DNA) == Sequence("ACGATCGATCG", NATYPES.DNA)
    assert Sequence("ACGATCGATCG", NATYPES.DNA) != Sequence("ACGATCGATCG", NATYPES.RNA)
    s = Sequence("ACGATCGATCG", NATYPES.DNA)
    assert list(s.batches(3, 5)) == [("ACGAT", 0), ("ATCGA", 3), ("GATCG", 6)]  # test_seq.py:136-138
    assert Sequence.mkrc("ACGA", NATYPES.DNA) == "TCGT"  # test_seq.py:152-181


def test_sequence_count_text_round_trip():
    with pytest.raises(AssertionError):
        SequenceCount("ACGATCGATCG", [1, 2, 3], NATYPES.DNA)
    h = [str(SequenceCoords("chr1", 0, 1000)), str(SequenceCoords("chr1", 1000, 2000))]
    sc = SequenceCount("ACGATCGATCG", h, NATYPES.DNA)
    assert sc.header == h and sc.seq == sc.text
    assert str(sc) == "ACGATCGATCG\t" + " ".join(h)
    assert sc == SequenceCount.from_text(str(sc)) and sc.as_text() == str(sc) + "\n"


def test_host_batch_lifecycle(tmp_path):
    with pytest.raises(AssertionError):
        Batch(KMer, str(tmp_path), 0)
    b = Batch(KMer, str(tmp_path), 5)
    assert (b.size, b.remaining, b.current_size, b.is_written, b.type) == (5, 5, 0, False, KMer)
    with pytest.raises(AssertionError):
        b.add("not a kmer")
    recs = [KMer("c", i, i + 4, s) for i, s in enumerate(["TTTT", "ACGT", "ACGA", "GGGG", "ACGT"])]
    b.add_all(recs)
    assert b.is_full() and b.remaining == 0
    with pytest.raises(AssertionError):
        b.add(recs[0])
    assert [r.seq for r in b.sorted()] == ["ACGA", "ACGT", "ACGT", "GGGG", "TTTT"]
    assert [r.coords.start for r in b.sorted() if r.seq == "ACGT"] == [1, 4]  # stable
    assert list(b.record_gen()) == recs
    b.write()
    assert b.is_written and os.path.isfile(b.tmp)
    assert open(b.tmp).read() == "".join(r.as_fasta() for r in recs)
    assert list(b.record_gen()) == recs
    linked = Batch.from_file(b.tmp)
    assert linked.current_size == 5 and list(linked.record_gen()) == recs
    with pytest.raises(AssertionError):
        b.add(recs[0])
    part = Batch(KMer, str(tmp_path), 9)
    part.add_all(recs[:3])
    part.write(doSort=True)
    assert [r.seq for r in part.record_gen()] == ["ACGA", "ACGT", "TTTT"]
    part.unwrite()
    assert not part.is_written and part.current_size == 3 and not os.path.isfile(part.tmp)
    b.reset()
    assert (b.current_size, b.remaining, b.is_written) == (0, 5, False)


def test_batch_appendable(tmp_path):
    h = ["chr1:0-4:+"]
    b = BatchAppendable(SequenceCount, str(tmp_path), 4)
    for s in ("TTTT", "ACGT", "GGGG"):
        b.add(SequenceCount(s, h))
    assert b.current_size == 3 and [r.seq for r in b.record_gen()] == ["TTTT", "ACGT", "GGGG"]
    with pytest.raises(AssertionError):
        b.add("x")
    b.reset()
    assert b.current_size == 0 and list(b.record_gen()) == []


def test_batcher_and_joiner_argument_validation(tmp_path):
    with pytest.raises(AssertionError):
        BatcherThreading(0)
    bt = BatcherThreading(10, threads=10**6)
    assert 1 <= bt.threads <= os.cpu_count() and bt.size == 10 and len(bt.collection) == 1
    fb = FastaBatcher.__new__(FastaBatcher)
    with pytest.raises(AssertionError):
        FastaBatcher.mode.fset(fb, "KMERS")
    with pytest.raises(AssertionError):
        FastaBatcher.doReverseComplement.fset(fb, 1)
    assert [m.name for m in FastaBatcher.MODE] == ["KMERS", "RECORDS"]
    assert [m.name for m in BatcherThreading.FEED_MODE] == ["REPLACE", "FLOW", "APPEND"]
    assert [m.name for m in KJoiner.MODE] == ["UNIQUE", "SEQ_COUNT", "VEC_COUNT", "VEC_COUNT_MASKED"]
    assert [m.name for m in KJoiner.MEMORY] == ["NORMAL", "LOCAL"]
    j = KJoinerThreading()
    assert j.mode == KJoiner.MODE.UNIQUE and j.memory == KJoiner.MEMORY.NORMAL and j.threads == 1
    with pytest.raises(AssertionError):
        j.batch_size = 1
    with pytest.raises(AssertionError):
        j.batch_size = 2.0
    with pytest.raises(AssertionError):
        j.doSort = 1
    with pytest.raises(AssertionError):
        KJoiner(mode="UNIQUE")
    j.threads = 10**6
    assert 1 <= j.threads <= os.cpu_count()
    # -B: the folder must exist and be non-empty (the reference's docstring; its shipped guard is inverted,
    # batcher.py:631, SURVEY Appendix A5 -- fixed here on purpose).  A non-empty folder needs the GPU.
    (tmp_path / "empty").mkdir()
    with pytest.raises(AssertionError):
        load_batches(str(tmp_path / "empty"))
    with pytest.raises(AssertionError):
        load_batches(str(tmp_path / "missing"))


def test_host_batches_join_like_the_reference(tmp_path):
    """KJoiner.join over host Batch objects follows the reference's crawler protocol
    (join.py:63-130,243-285): merge of sorted batches, grouping, emit rules."""
    a, b = Batch(KMer, str(tmp_path), 4), Batch(KMer, str(tmp_path), 4)
    a.add_all([KMer("c", 0, 4, "ACGT"), KMer("c", 1, 5, "CGTA"), KMer("c", 2, 6, "GTAC")])
    b.add_all([KMer("d", 0, 4, "ACGT"), KMer("d", 1, 5, "TTTT")])
    a.write(doSort=True)
    b.write(doSort=True)
    groups = list(Crawler().do_batch([a, b]))
    assert groups == [(["c:0-4:+", "d:0-4:+"], "ACGT"), (["c:1-5:+"], "CGTA"), (["c:2-6:+"], "GTAC"), (["d:1-5:+"], "TTTT")]
    out = tmp_path / "count.tsv"
    KJoinerThreading(KJoiner.MODE.SEQ_COUNT).join([a, b], str(out))
    assert out.read_text() == "ACGT\t2\nCGTA\t1\nGTAC\t1\nTTTT\t1\n"
    out = tmp_path / "uniq.fa"
    KJoinerThreading().join([a, b], str(out))
    assert out.read_text() == ">c:1-5:+\nCGTA\n>c:2-6:+\nGTAC\n>d:1-5:+\nTTTT\n"


def test_host_batches_join_abundance_vectors(tmp_path):
    """VEC_COUNT / VEC_COUNT_MASKED over host Batch objects (join.py:287-372, abundance.py:92-172): the
    reference's record-by-record protocol, with its crashing abstract call removed."""
    import gzip

    from kman_b200.abundance import AbundanceVector, vector_text

    a, b = Batch(KMer, str(tmp_path), 4), Batch(KMer, str(tmp_path), 4)
    a.add_all([KMer("c", 0, 4, "ACGT"), KMer("c", 1, 5, "CGTA"), KMer("c", 3, 7, "ACGT")])
    b.add_all([KMer("d", 0, 4, "ACGT"), KMer("d", 2, 6, "TTTT")])
    a.write(doSort=True)
    b.write(doSort=True)
    out = tmp_path / "vec.txt"
    KJoinerThreading(KJoiner.MODE.VEC_COUNT).join([a, b], str(out))
    read = lambda f: gzip.open(tmp_path / "vec" / f, "rb").read()  # noqa: E731
    assert read("c___+.gz") == b"# k=4\n3\n1\n0\n3\n"
    assert read("d___+.gz") == b"# k=4\n3\n0\n1\n"
    out2 = tmp_path / "masked.txt"
    KJoinerThreading(KJoiner.MODE.VEC_COUNT_MASKED).join([a, b], str(out2))
    read2 = lambda f: gzip.open(tmp_path / "masked" / f, "rb").read()  # noqa: E731
    assert read2("c___+.gz") == b"# k=4\n1\n0\n0\n1\n"  # ACGT: one occurrence in the other record
    assert read2("d___+.gz") == b"# k=4\n2\n"
    v = AbundanceVector()
    v.add_count("r", "+", 2, 5, 4)
    with pytest.raises(AssertionError):
        v.add_count("r", "+", 2, 1, 4)  # abundance.py:128-131
    with pytest.raises(AssertionError):
        v.add_count("r", "+", 3, 1, 5)  # abundance.py:37-39: one k per join
    assert vector_text([0, 7, 10, 123456, 99], 21) == b"# k=21\n0\n7\n10\n123456\n99\n"
    assert vector_text([], 3) == b"# k=3\n"
