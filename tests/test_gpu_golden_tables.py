"""BASELINE.json's configurations 2, 3 (scaled), 4 and 5 at sizes the reference cannot finish: the
binary result tables of the CUDA path are hashed and compared with tests/golden/table_hashes.json,
which oracle/gen_table_hashes.py computed in the build container with the oracle's numpy tier
(kmermaid/seq.py:284-328 extraction, batch.py:156-168 order, join.py:95-130 grouping, :243-285 emit).
The inputs are rebuilt here from the same seeded generators (oracle/synth_configs.py)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import synth_configs as sc  # noqa: E402  (oracle/ is on sys.path, see conftest.py)

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "table_hashes.json")))

MAKERS = {
    "cfg2_100mbp_k31_count": lambda: sc.cfg2(100_000_000),
    "cfg2dup_100mbp_k31_count": lambda: sc.cfg2(100_000_000, dup=True),
    "cfg2_100mbp_k31_uniq": lambda: sc.cfg2(100_000_000),
    "cfg3_24mbp_24rec_k31_count": lambda: sc.cfg3(24_000_000, 24),
    "cfg3_24mbp_24rec_k31_uniq": lambda: sc.cfg3(24_000_000, 24),
    "cfg4_10mbp_k25_uniq_iupac": lambda: sc.cfg4(10_000_000),
    "cfg4_10mbp_k25_uniq_acgt": lambda: sc.cfg4(10_000_000),
    "cfg4_10mbp_k25_uniq_rc_iupac": lambda: sc.cfg4(10_000_000),
    "cfg4_10mbp_k25_count_iupac": lambda: sc.cfg4(10_000_000),
    "cfg4_100mbp_k25_uniq_iupac": lambda: sc.cfg4(100_000_000),
    "cfg4_10mbp_k45_uniq_iupac": lambda: sc.cfg4(10_000_000),
    "cfg5_20mbp_k63_count": lambda: sc.cfg5(20_000_000),
    "cfg5dup_20mbp_k63_count": lambda: sc.cfg5(20_000_000, dup=True),
    "cfg5_20mbp_k63_uniq_rc": lambda: sc.cfg5(20_000_000),
}


@pytest.fixture(scope="module")
def eng():
    from kman_b200.engine import get_engine

    return get_engine()


def _rows(host_keys: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(host_keys, dtype=np.uint64)


def test_every_golden_case_is_tested():
    assert set(GOLD) == set(MAKERS)


@pytest.mark.parametrize("name", sorted(MAKERS))
def test_table_hash_matches_oracle(eng, name):
    import torch

    from kman_b200 import fasta

    g = GOLD[name]
    recs = MAKERS[name]()
    d = eng.upload(fasta.from_records(recs), alphabet=g["alphabet"], with_names=False)
    del recs
    k, rc = g["k"], g["rc"]
    streams = eng.count(d, k, rc) if g["mode"] == "count" else eng.uniq(d, k, rc)
    for si, sname in enumerate(("narrow", "wide")):
        want = g[sname]
        if si >= len(streams):
            assert want["rows"] == 0, f"{sname}: the device returned no such stream, the oracle has {want['rows']} rows"
            continue
        st = streams[si]
        assert st.n == want["rows"], (sname, st.n, want["rows"])
        keys = _rows(st.keys_host())
        if g["mode"] == "count":
            got = sc.count_digest(keys, st.counts_host())
            assert got["total"] == want["total"]
            assert got["counts_sha256"] == want["counts_sha256"], sname
        else:
            got = sc.uniq_digest(keys, st.vals_host().astype(np.uint64))
            assert got["vals_sha256"] == want["vals_sha256"], sname
        assert got["keys_sha256"] == want["keys_sha256"], sname
    del streams, d
    torch.cuda.empty_cache()
