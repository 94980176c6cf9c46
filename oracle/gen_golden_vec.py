#!/usr/bin/env python
"""Generate tests/golden/golden_vec.json: `kmer count -m VEC_COUNT / VEC_COUNT_MASKED` outputs of the reference.

TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).  As shipped the reference
crashes in these modes: AbundanceVector.add_count calls super().add_count and the ABSTRACT base method
raises NotImplementedError (kmermaid/abundance.py:123 -> :60).  The generator
imports the reference unmodified and neutralises exactly that one abstract method at run time
(`AbundanceVectorBase.add_count = no-op`); every other line of kmermaid/join.py:287-335 and
abundance.py:92-172 runs as written.  Stored: the DECOMPRESSED content of every REF___STRAND.gz file
(gzip headers carry a time stamp).

usage: python oracle/gen_golden_vec.py
"""
from __future__ import annotations

import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, HERE)
import kmer_oracle as ko  # noqa: E402

PATCHED_MAIN = ("import kmermaid.abundance as ab\n"
                "ab.AbundanceVectorBase.add_count = lambda self, *a, **k: None\n"
                "from kmermaid.scripts.kmer import main; main()\n")

TEXTS = {
    "tiny": ">chr1 first record\nACGATCGATCGnnNACGTacgtACGA\nTTGCA\n>chr2\nACGATCGGGTTTACGT\n",
    "shared": ">a x\nACGTACGTTGCAACGT\n>b\nTTGCAACGTACGAAA\n>c\nGGGACGTACG\n",
    "iupac_mix": ">r1\nACGTRYKMSWBDHVNACGTNNACGT\n>r2\nacgtnACGTuACGTXACGT-ACGT*ACGT\n",
    "gaps": ">g1\nACGTACGTNNNNNNACGTACGT\n>g2\nNNNACGTACGTNNN\n>short\nAC\n",
    "palindromes": ">p\nACGTACGTTCGAATTCGGATCC\n>q\nGGATCCGAATTCG\n",
}


def run_ref(fasta, k, alphabet, mode, rc):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), REF])
    env["KMG_ORACLE_ALPHABET"] = alphabet
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.txt")
        argv = [sys.executable, "-c", PATCHED_MAIN, "count", "-m", mode] + (["-r"] if rc else []) + [fasta, out, str(k)]
        p = subprocess.run(argv, env=env, cwd=td, capture_output=True, timeout=600)
        if p.returncode != 0:
            return {"error": p.stderr.decode(errors="replace").strip().splitlines()[-1]}
        d = os.path.join(td, "out")
        if not os.path.isdir(d):
            return {}
        return {f: gzip.open(os.path.join(d, f), "rb").read().decode("latin-1") for f in sorted(os.listdir(d))}


def main():
    cases = []
    with tempfile.TemporaryDirectory() as td:
        for name, text in TEXTS.items():
            fa = os.path.join(td, name + ".fa")
            with open(fa, "w", newline="") as fh:
                fh.write(text)
            for k in (3, 4, 7):
                for alphabet in ("IUPAC", "ACGT"):
                    for rc in (False, True):
                        for mode in ("VEC_COUNT", "VEC_COUNT_MASKED"):
                            files = run_ref(fa, k, alphabet, mode, rc)
                            cases.append({"name": name, "fasta_text": text, "k": k, "alphabet": alphabet, "rc": rc,
                                          "mode": mode, "files": files})
                            print(name, k, alphabet, rc, mode, len(files), flush=True)
        # a duplicated multi-record synthetic input at real k
        seq = ko.synth_bases(3000, 7)
        recs = [("chrA dup", seq[:2000] + seq[:1000]), ("chrB", seq[1500:3000]), ("chrC tail", seq[100:400])]
        text = ko.synth_fasta_bytes(recs).decode()
        fa = os.path.join(td, "syn_dup.fa")
        open(fa, "w").write(text)
        for k in (11, 31, 45):
            for rc in (False, True):
                for mode in ("VEC_COUNT", "VEC_COUNT_MASKED"):
                    files = run_ref(fa, k, "IUPAC", mode, rc)
                    cases.append({"name": "syn_dup", "fasta_text": text, "k": k, "alphabet": "IUPAC", "rc": rc, "mode": mode,
                                  "files": files})
                    print("syn_dup", k, rc, mode, len(files), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "golden_vec.json")
    with open(out, "w") as fh:
        json.dump({"_about": "produced by oracle/gen_golden_vec.py from the reference with AbundanceVectorBase.add_count "
                             "neutralised at run time; do not edit by hand", "cases": cases}, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
