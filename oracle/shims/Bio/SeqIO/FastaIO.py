"""Stand-in for Bio.SeqIO.FastaIO.SimpleFastaParser (biopython 1.79 semantics).
TEST INFRASTRUCTURE ONLY (see oracle/shims/oligo_melting/__init__.py)."""


def SimpleFastaParser(handle):
    for line in handle:
        if line[0] == ">":
            title = line[1:].rstrip()
            break
    else:
        return
    lines = []
    for line in handle:
        if line[0] == ">":
            yield title, "".join(lines).replace(" ", "").replace("\r", "")
            lines = []
            title = line[1:].rstrip()
            continue
        lines.append(line.rstrip())
    yield title, "".join(lines).replace(" ", "").replace("\r", "")
