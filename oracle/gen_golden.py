#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference (`/root/reference`).

TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no
/root/reference); the vectors it writes are committed.  The reference's missing
third-party imports are satisfied by the stand-ins under oracle/shims/ (no kmermaid
code is copied or edited).  Every run uses the reference's one correct configuration:
default scan mode (KMERS), 1 thread.

usage: python oracle/gen_golden.py [--big]     (--big adds the 1 Mbp config-1 run, ~2 min)
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, HERE)
import kmer_oracle as ko  # noqa: E402


def run_ref(cmd, fasta, k, alphabet, extra=(), timeout=900):
    """Run `kmer <cmd> [extra] fasta OUT k` with the reference; return output bytes
    (for `batch`: list of file bytes sorted by content)."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), REF])
    env["KMG_ORACLE_ALPHABET"] = alphabet
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out")
        argv = [sys.executable, "-c", "from kmermaid.scripts.kmer import main; main()", cmd, *extra, fasta, out, str(k)]
        p = subprocess.run(argv, env=env, cwd=td, capture_output=True, timeout=timeout)
        if p.returncode != 0:
            return {"error": p.stderr.decode(errors="replace").strip().splitlines()[-1]}
        if cmd == "batch":
            return sorted(open(os.path.join(out, f), "rb").read() for f in os.listdir(out))
        return open(out, "rb").read()


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


TINY = ">chr1 first record\nACGATCGATCGnnNACGTacgtACGA\nTTGCA\n>chr2\nACGATCGGGTTTACGT\n"

# Edge-case FASTA texts (SURVEY.md §8a "exact semantics", Appendix B edge probes)
EDGE = {
    "tiny": TINY,
    "no_trailing_newline": ">a\nACGTACGTAC",
    "preamble_comments": "; comment\n\n>a desc here\nACGTTGCA\nGGCC\n",
    "crlf_spaces_blank": ">a\r\nACG TAC\r\n\r\nGTA CGT\r\n>b x\r\nTTTTACGT\r\n",
    "short_record": ">a\nACG\n>b\nACGTACGT\n",
    "duplicate_names": ">a\nACGTACGGT\n>a\nACGTACCGT\n",
    "tab_in_name": ">a\tb c\nACGTACGT\n",
    "iupac_mix": ">r1\nACGTRYKMSWBDHVNACGTNNACGT\n>r2\nacgtnACGTuACGTXACGT-ACGT*ACGT\n",
    "all_n": ">n\nNNNNNNNNNN\n",
    "k_gt_len": ">a\nACGT\n",
    "palindromes": ">p\nACGTACGTTCGAATTCGGATCC\n",
    "lower_mixed": ">soft\nacgtACGTacgtNNNNacgtACGTTTGA\n>soft2\nggggCCCCaaaaTTTT\n",
    "repeat": ">r\nACGACGACGACGACGACGACG\n>r2\nACGACGACG\n",
}
EDGE_KS = {"k_gt_len": [5], "all_n": [3]}


def main():
    os.makedirs(GOLD, exist_ok=True)
    big = "--big" in sys.argv
    index = {"_about": "produced by oracle/gen_golden.py from the unmodified reference; do not edit by hand",
             "cases": []}
    with tempfile.TemporaryDirectory() as td:
        # ---- small edge cases: keep full outputs --------------------------------------
        for name, text in EDGE.items():
            fa = os.path.join(td, name + ".fa")
            with open(fa, "w", newline="") as fh:
                fh.write(text)
            for k in EDGE_KS.get(name, [2, 4, 7]):
                for alphabet in ("IUPAC", "ACGT"):
                    for rc in (False, True):
                        extra = ("-r",) if rc else ()
                        case = {"name": name, "fasta_text": text, "k": k, "alphabet": alphabet, "rc": rc}
                        for cmd in ("count", "uniq"):
                            r = run_ref(cmd, fa, k, alphabet, extra)
                            case[cmd] = r if isinstance(r, dict) else r.decode("latin-1")
                        if name in ("tiny", "repeat") and k == 4:
                            r = run_ref("batch", fa, k, alphabet, extra + ("-b", "10"))
                            case["batch_b10"] = [x.decode("latin-1") for x in r]
                        index["cases"].append(case)
                        print("edge", name, k, alphabet, rc, "ok", flush=True)
        # batch-size independence (probe in SURVEY §8a): b=3 vs default must agree
        fa = os.path.join(td, "tiny.fa")
        for b in (3, 7):
            for cmd in ("count", "uniq"):
                r = run_ref(cmd, fa, 4, "IUPAC", ("-b", str(b)))
                ref = [c for c in index["cases"] if c["name"] == "tiny" and c["k"] == 4 and c["alphabet"] == "IUPAC" and not c["rc"]][0]
                assert r.decode("latin-1") == ref[cmd], ("batch-size dependence", b, cmd)
        # ---- synthetic configs: keep hashes ------------------------------------------
        synth = [("syn_100k", 100_000, 1234, 21)]
        if big:
            synth.append(("syn_1m", 1_000_000, 1234, 21))
        index["synthetic"] = []
        for name, n, seed, k in synth:
            data = ko.synth_fasta_bytes([("chr1 synthetic seed=%d" % seed, ko.synth_bases(n, seed))])
            fa = os.path.join(td, name + ".fa")
            open(fa, "wb").write(data)
            ent = {"name": name, "n": n, "seed": seed, "k": k, "fasta_sha256": sha(data), "fasta_bytes": len(data)}
            for cmd in ("count", "uniq"):
                r = run_ref(cmd, fa, k, "IUPAC", timeout=3600)
                ent[cmd + "_sha256"] = sha(r)
                ent[cmd + "_bytes"] = len(r)
                ent[cmd + "_lines"] = r.count(b"\n")
                ent[cmd + "_head"] = r[:200].decode()
            index["synthetic"].append(ent)
            print("synthetic", name, "ok", flush=True)
        # a multi-record duplicated synthetic (counts > 1, -r) small enough for full reference run
        seq = ko.synth_bases(3000, 7)
        recs = [("chrA dup", seq[:2000] + seq[:1000]), ("chrB", seq[1500:3000]), ("chrC tail", seq[100:400])]
        data = ko.synth_fasta_bytes(recs)
        fa = os.path.join(td, "syn_dup.fa")
        open(fa, "wb").write(data)
        for k in (11, 31, 33, 63):
            for rc in (False, True):
                ent = {"name": "syn_dup", "k": k, "rc": rc, "fasta_text": data.decode()}
                for cmd in ("count", "uniq"):
                    r = run_ref(cmd, fa, k, "IUPAC", ("-r",) if rc else ())
                    ent[cmd + "_sha256"] = sha(r)
                    ent[cmd + "_lines"] = r.count(b"\n")
                index["synthetic"].append(ent)
                print("syn_dup", k, rc, "ok", flush=True)
    if not big:
        # keep a previously generated 1 Mbp entry
        old = os.path.join(GOLD, "golden.json")
        if os.path.exists(old):
            prev = json.load(open(old))
            index["synthetic"] += [e for e in prev.get("synthetic", []) if e["name"] == "syn_1m"]
    with open(os.path.join(GOLD, "golden.json"), "w") as fh:
        json.dump(index, fh, indent=1)
    print("wrote", os.path.join(GOLD, "golden.json"))


if __name__ == "__main__":
    main()
