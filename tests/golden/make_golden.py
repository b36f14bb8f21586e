#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REAL reference programs.

Run in the build container (needs /root/reference and oracle/_ref/{dsk,set_finder}, built by
`make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  edge_q.fq                hand-made FASTQ / mixed quirk file (Bank.cpp:173-201)
  edge_a.fa                hand-made mode-A quirk file (N runs, IUPAC, lower case, CRLF, blank
                           lines, empty / short records, no trailing newline)
  dsk_golden.json          per case and k: distinct, sum, md5[:12] of the sorted "KMER COUNT"
                           lines of the real DSK output; full {kmer: count} for small k
  set_finder_golden.json   set_finder stdout for the graph of swga/tests/test_find_sets_cmd.py:35-48

Cases whose FASTA is produced by a seeded recipe (swga_b200.synth) are re-created by the tests
from the same recipe, so only the DSK answers are stored.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import swga_oracle as O          # noqa: E402
from swga_b200 import synth                  # noqa: E402

REF_DATA = os.path.join(HERE, "ref")         # copied from /root/reference/swga/tests/data (test vectors)

EDGE_A = (
    b">r1 multi-line record with N runs > and @ inside the header\r\n"
    b"NNACGTACGTNNNNACGTTGCA\r\n"
    b"ACGTNACGTnACGTacgtRYKMSWN\r\n"
    b"\n"
    b"GGGGCCCCAAAATTTTN\n"
    b">r2_empty\n"
    b">r3_short\n"
    b"ACG\n"
    b">r4 trailing N and homopolymers\n"
    b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n"
    b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTNNNN\n"
    b">r5 palindromes\n"
    b"ACGTACGTACGTACGTAATTAATTGGCCGGCCGCGCGCGCATATATAT\n"
    b">r6 no trailing newline\n"
    b"GATTACAGATTACAGATTACANGATTACA"
)


# FASTQ and mixed input (Bank.cpp:173-201): 4-line records, CRLF, quality lines that start with '@' / contain '>',
# multi-line reads whose quality is cut short by a long header comment (dummy still holds the rest of the header
# line), a FASTA record in between, left-over quality that contains '@' and is then taken for a header, no final
# newline.
EDGE_Q = (
    b"@r1 some comment\n"
    b"ACGTACGTNNACGTTGCAAC\n"
    b"+\n"
    b"IIIIIIIIIIIIIIIIIIII\n"
    b"@r2\r\n"
    b"ACGTTGCATGCA\r\n"
    b"+r2\r\n"
    b"@@@@>>>>++++\r\n"
    b"@r3 x\n"
    b"ACGTACGTAC\n"
    b"GGGTTTAAAC\n"
    b"+\n"
    b"IIIIIIIIII\n"
    b"IIIIIIIIII\n"
    b">fa1 fasta record inside\n"
    b"ACGTNNNNACGTACGTAGCTAGCTAGCTAG\n"
    b"\n"
    b"acgtnRYKM\n"
    b"@r4 long comment so that dummy is long already\n"
    b"ACGTAC\n"
    b"+\n"
    b"II>III\n"
    b"@r6 0123456789\n"
    b"ACGTACGTAC\n"
    b"GGGTTTAAAC\n"
    b"+\n"
    b"IIIIIIIIII\n"
    b"III@IIIIII\n"
    b"GATTACAGATTACA\n"
    b"@r5\n"
    b"TTTTTTTTGGGGGGGGCCCCCCCCAAAAAAAA\n"
    b"+\n"
    b"!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!"
)


def recipe_mode_b():
    """One 1,000,050-bp record (forces DSK mode B) with N / n / IUPAC sprinkled in, plus short
    records containing N."""
    seq = synth.bases(101, 0, 1_000_050 + 400, 0.41)
    idx = synth.splitmix64(np.arange(300, dtype=np.uint64) + np.uint64(7 << 40)) % np.uint64(len(seq))
    seq[idx[:150].astype(np.int64)] = ord("N")
    seq[idx[150:200].astype(np.int64)] = ord("n")
    seq[idx[200:250].astype(np.int64)] = ord("R")
    low = idx[250:].astype(np.int64)
    seq[low] = seq[low] | 0x20                       # lower case
    seq[500:540] = ord("N")                          # a long N run inside the big record
    seq[1_000_049] = ord("N")                        # N as the last base of the big record
    seq[1_000_050] = ord("N")                        # N as the first base of the next record
    rs = np.array([0, 1_000_050, 1_000_050 + 7, 1_000_050 + 7, 1_000_050 + 200, 1_000_050 + 400],
                  dtype=np.uint64)
    return seq, rs


def recipe_mode_a():
    """300 records of varying length (<= 1 Mbp => mode A), AT-rich, with N runs."""
    n = 600_000
    seq = synth.bases(102, 0, n, 0.25)
    cuts = np.sort(synth.splitmix64(np.arange(299, dtype=np.uint64) + np.uint64(9 << 40)) % np.uint64(n))
    rs = np.concatenate([[0], cuts, [n]]).astype(np.uint64)
    runs = synth.splitmix64(np.arange(400, dtype=np.uint64) + np.uint64(11 << 40))
    for r in runs:
        p = int(r % np.uint64(n))
        ln = int((r >> np.uint64(40)) % np.uint64(20)) + 1
        seq[p:p + ln] = ord("N")
    return seq, rs


RECIPES = {"synth_mode_b": recipe_mode_b, "synth_mode_a": recipe_mode_a}


def digest(table_dict):
    lines = sorted("%s %d" % kv for kv in table_dict.items())
    return hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest()[:12]


def dsk_counts(fp, k, tmp):
    out = O.run_dsk(fp, k, tmp)
    d = dict(O.parse_kmer_binary(out))
    os.remove(out)
    return d


def main():
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    gold = {}
    if only:                                   # regenerate some cases, keep the others as they are
        with open(os.path.join(HERE, "dsk_golden.json")) as f:
            gold = json.load(f)
    with tempfile.TemporaryDirectory() as tmp:
        cases = {}
        with open(os.path.join(HERE, "edge_a.fa"), "wb") as f:
            f.write(EDGE_A)
        cases["edge_a"] = (os.path.join(HERE, "edge_a.fa"), [3, 4, 5, 8, 12], [3, 4, 5, 8, 12])
        with open(os.path.join(HERE, "edge_q.fq"), "wb") as f:
            f.write(EDGE_Q)
        cases["edge_q"] = (os.path.join(HERE, "edge_q.fq"), [3, 4, 5, 8, 12], [3, 4, 5, 8, 12])
        for name, fn in RECIPES.items():
            seq, rs = fn()
            fp = os.path.join(tmp, name + ".fa")
            synth.write_fasta(fp, seq, rs)
            cases[name] = (fp, list(range(5, 13)), [5, 6])
        for fn in ("test.fasta", "bburgdorferi_1k.fa", "ecoli_2k.fa", "bburgdorferi_wgs.fa",
                   "simple_fg_genome.fa"):
            cases["ref:" + fn] = (os.path.join(REF_DATA, fn), [4] + list(range(5, 13)), [4, 5])
        for name, (fp, ks, full_ks) in cases.items():
            if only and name not in only:
                continue
            gold[name] = {}
            for k in ks:
                d = dsk_counts(fp, k, tmp)
                e = {"distinct": len(d), "sum": int(sum(d.values())), "md5": digest(d)}
                if k in full_ks:
                    e["counts"] = d
                gold[name][str(k)] = e
                print(name, k, e["distinct"], e["sum"], e["md5"], flush=True)
    with open(os.path.join(HERE, "dsk_golden.json"), "w") as f:
        json.dump(gold, f, indent=0, sort_keys=True)
    if only:
        return

    # set_finder: the graph of swga/tests/test_find_sets_cmd.py:35-48 and its expected line
    graph = ("p sp 6 7\n" "n 1 1\nn 2 2\nn 3 1\nn 4 4\nn 5 5\nn 6 6\n"
             "e 1 2\ne 1 3\ne 2 3\ne 2 4\ne 4 5\ne 4 6\ne 5 6\n")
    sf = {}
    with tempfile.TemporaryDirectory() as tmp:
        gp = os.path.join(tmp, "g.dimacs")
        with open(gp, "w") as f:
            f.write(graph)
        out = subprocess.run(
            [os.path.join(ROOT, "oracle", "_ref", "set_finder"), "-q", "-q", "--bg_freq", "2", "--bg_len", "10",
             "--min", "3", "--max", "3", "--unweighted", "--all", "--reorder", "weighted-coloring", gp],
            stdout=subprocess.PIPE, check=False).stdout.decode()
    sf["graph"] = graph
    sf["stdout"] = out
    with open(os.path.join(HERE, "set_finder_golden.json"), "w") as f:
        json.dump(sf, f, indent=1, sort_keys=True)
    print("set_finder:", repr(out))


if __name__ == "__main__":
    main()
