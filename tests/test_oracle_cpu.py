"""The oracle against the real DSK binary's golden vectors (tests/golden/dsk_golden.json, made by
tests/golden/make_golden.py) and the reference's own known-answer tests."""
import os

import numpy as np
import pytest

from helpers import golden_case_bytes, table_digest, table_dict

CASES = ["edge_a", "edge_q", "synth_mode_a", "synth_mode_b", "ref:test.fasta", "ref:bburgdorferi_1k.fa",
         "ref:ecoli_2k.fa", "ref:bburgdorferi_wgs.fa", "ref:simple_fg_genome.fa"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_counts_match_real_dsk(case, oracle, dsk_golden):
    data = golden_case_bytes(case)
    if data is None:
        pytest.skip("reference tree not present")
    seq, rs = oracle.fasta_parse_dsk(data)
    for ks, gold in dsk_golden[case].items():
        k = int(ks)
        t = oracle.count_dense(seq, rs, k)
        assert table_digest(t, k) == (gold["distinct"], gold["sum"], gold["md5"]), (case, k)
        if "counts" in gold:
            assert table_dict(t, k) == gold["counts"]


def test_oracle_modes(oracle):
    assert oracle.dsk_mode(np.array([0, 1000000], dtype=np.uint64)) == 0
    assert oracle.dsk_mode(np.array([0, 1000001], dtype=np.uint64)) == 1
    # only the first 10,001 records are inspected (Bank.cpp:470-494)
    rs = np.arange(10003, dtype=np.uint64)
    rs[-1] += 2000000
    assert oracle.dsk_mode(rs) == 0
    rs = np.arange(10002, dtype=np.uint64)
    rs[-1] += 2000000
    assert oracle.dsk_mode(rs) == 1


def test_survey_b1_known_answer(oracle):
    # real DSK on swga/tests/data/test.fasta, k=4
    fa = b"> record1\nAAGGAAGGCCTTCCTT\n> record2\nAAAATTTT\n> record3\nAAGGCCTT"
    assert oracle.count_kmers_dict(fa, 4) == {
        "AAAA": 2, "AAAT": 2, "AAGG": 6, "AATT": 1, "AGGA": 2, "AGGC": 4, "CTTC": 2, "GGCC": 2, "TTCC": 2}


def test_survey_a3_probe(oracle):
    # SURVEY.md App. A3 probe: split at 'N', 'n' -> G
    fa = b">x\nACGTNACGTnACGTacgt\n"
    assert oracle.count_kmers_dict(fa, 4) == {
        "ACGT": 4, "CACG": 1, "CGTA": 2, "CGTC": 1, "GTAC": 1, "TCAC": 1, "TGAC": 1}


def test_parse_kmer_binary_roundtrip(oracle, tmp_path):
    import struct
    fp = tmp_path / "x.solid_kmers_binary"
    with open(fp, "wb") as f:
        f.write(struct.pack("ii", 64, 4))
        f.write(struct.pack("<QI", oracle.kmer_to_code("AAGG"), 6))
        f.write(struct.pack("<QI", oracle.kmer_to_code("CTTC"), 2))
        f.write(b"\x00\x00")                                   # short trailing read is ignored
    assert list(oracle.parse_kmer_binary(str(fp))) == [("AAGG", 6), ("CTTC", 2)]


# ---- reference known-answer tests for locate / score (swga/tests/test_locate.py, test_score.py)

TEST_FASTA = b"> record1\nAAGGAAGGCCTTCCTT\n> record2\nAAAATTTT\n> record3\nAAGGCCTT"


def test_locate_kat(oracle):
    recs = oracle.fasta_records_pyfaidx(TEST_FASTA)
    loc = oracle.binding_sites("AAGG", recs)
    assert len(loc["record1"]) == 4 and loc["record2"] == []          # test_locate.py:12-15
    ends = oracle.chromosome_ends(recs)
    assert ends == {"record1": [0, 15], "record2": [16, 23], "record3": [24, 31]}   # :18-22
    lin = oracle.linearize_binding_sites([loc], ends)
    assert len(lin) == 10                                                 # :25-37
    for rec, (s, e) in ends.items():
        assert s in lin and e in lin
        for site in loc[rec]:
            assert site + s in lin
    assert oracle.revcomp("GCAT") == "ATGC"                               # :40-41


def test_score_kat(oracle):
    assert oracle.read_set_finder_line("1,2,3,4 3500.01234") == ([1, 2, 3, 4], 3500.01234)
    assert oracle.seq_diff([1, 3, 4, 5]) == [2, 1, 1]
    assert oracle.seq_diff([3, 1, 5, 4]) == [2, 1, 1]
    score, ns = oracle.default_score_set("fg_dist_mean/bg_dist_mean", [1, 2, 3, 4], [1, 3, 5, 7], 2, 3.0)
    assert score == 2.0 / 3.0 and "__builtins__" not in ns               # test_score.py:18-29


def test_gini_python2_floor(oracle):
    # gaps [1,1,1] -> fair_area = 4 (floor of 4.5), area = 4.5, Gini = -0.125
    assert oracle.gini([1, 1, 1]) == -0.125


def test_survey_b3_simple_genome(oracle):
    # simple_fg_genome.fa (AAAATTTT x 72) with the palindromic primer AAAATTTT
    fa = b">simple_genome\n" + b"\n".join([b"AAAATTTT" * 8] * 9) + b"\n"
    recs = oracle.fasta_records_pyfaidx(fa)
    loc = oracle.binding_sites("AAAATTTT", recs)
    assert len(loc["simple_genome"]) == 144                               # each site twice
    assert loc["simple_genome"][:3] == [0, 8, 16]
    ends = oracle.chromosome_ends(recs)
    score, var, max_dist = oracle.score_set([loc], 36000, float("inf"), ends)
    assert max_dist == 8 and var["fg_dist_mean"] == 7.986111111111111
    assert var["fg_dist_std"] == pytest.approx(0.11785113019775792, rel=1e-12)
    assert var["fg_dist_gini"] == 0.001714975845410628
    assert score == 0.0
