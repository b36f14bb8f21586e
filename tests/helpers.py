"""Shared test helpers: golden-case inputs and digest of a dense canonical table."""
import hashlib
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REF_DATA = os.path.join(GOLDEN, "ref")      # the reference's own swga/tests/data fixtures, committed as test vectors

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def code_to_kmer(code, k):
    return "".join("ACTG"[(int(code) >> (2 * (k - 1 - j))) & 3] for j in range(k))


def table_digest(table, k):
    """(distinct, sum, md5[:12]) of a dense canonical table, in the golden file's format."""
    nz = np.nonzero(table)[0]
    lines = sorted("%s %d" % (code_to_kmer(c, k), table[c]) for c in nz)
    md5 = hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest()[:12]
    return len(nz), int(table.astype(np.int64).sum()), md5


def table_dict(table, k):
    return {code_to_kmer(c, k): int(table[c]) for c in np.nonzero(table)[0]}


def golden_case_bytes(name):
    """FASTA image of a golden case ("ref:<file>" = a fixture of the reference's swga/tests/data, kept under
    tests/golden/ref/ so that it also runs on the GPU box)."""
    from swga_b200 import synth
    if name == "edge_a":
        return open(os.path.join(GOLDEN, "edge_a.fa"), "rb").read()
    if name == "edge_q":
        return open(os.path.join(GOLDEN, "edge_q.fq"), "rb").read()
    if name in make_golden.RECIPES:
        import tempfile
        seq, rs = make_golden.RECIPES[name]()
        with tempfile.TemporaryDirectory() as tmp:
            fp = os.path.join(tmp, "x.fa")
            synth.write_fasta(fp, seq, rs)
            return open(fp, "rb").read()
    if name.startswith("ref:"):
        fp = os.path.join(REF_DATA, name[4:])
        return open(fp, "rb").read() if os.path.isfile(fp) else None
    raise KeyError(name)
