"""CPU-only: the checkers themselves.  Known-answer vectors for the literal oracle, and the C
oracle against the literal oracle on seeded genomes."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import c_oracle, py_oracle
from polycracker_b200 import binding, synth
from tests import util


def test_known_answer_survey_8c():
    r = py_oracle.run_path([("s1", "ACGTACGTAC")], chunk_size=5, k=3, kmer_low_count=3)
    assert [b[3] for b in r["bed"]] == ["s1_0_5", "s1_5_10"]
    assert r["counts"] == {"ACG": 3, "GTA": 3}
    assert r["kmers"] == ["ACG", "GTA"] and r["scaffolds"] == ["s1_0_5", "s1_5_10"]
    m = r["matrix"]
    assert m.format == "csc" and m.dtype == np.float32
    assert m.indptr.tolist() == [0, 2, 4] and m.indices.tolist() == [0, 1, 0, 1] and m.data.tolist() == [2, 1, 1, 2]
    t = r["tfidf"].toarray()
    np.testing.assert_allclose(t, np.array([[2, 1], [1, 2]]) / np.sqrt(5), rtol=1e-6)
    # max (BBTools) orientation: same counts, labels are the reverse complements
    r2 = py_oracle.run_path([("s1", "ACGTACGTAC")], 5, 3, 3, mode="max")
    assert r2["kmers"] == ["CGT", "TAC"] and (r2["matrix"] != m).nnz == 0
    # threshold 4: empty repeat set, matrix keeps its rows
    r3 = py_oracle.run_path([("s1", "ACGTACGTAC")], 5, 3, 4)
    assert r3["matrix"].shape == (2, 0)


def test_chunk_geometry_edge_cases():
    # L == m*C exactly: no empty tail; L == m*C + 1: one-base tail; L < C: single chunk; L == 0: nothing
    bed = py_oracle.split_chunks([("a", 10), ("b", 11), ("c", 3), ("d", 0)], 5)
    assert [r[3] for r in bed] == ["a_0_5", "a_5_10", "b_0_5", "b_5_10", "b_10_11", "c_0_3"]
    # row universe is string-sorted: chr1_100_... sorts before chr1_50_...
    bed = py_oracle.split_chunks([("chr1", 150)], 50)
    assert py_oracle.find_scaffolds(bed) == ["chr1_0_50", "chr1_100_150", "chr1_50_100"]
    assert py_oracle.filter_scaffolds(["x_0_50", "x_50_70"], 50) == ["x_0_50"]
    assert py_oracle.filter_scaffolds(["x_0_50", "x_50_70"], 50, remove_non_chunk=0, min_chunk_threshold=1,
                                      min_chunk_size=20) == ["x_0_50", "x_50_70"]


def test_counting_semantics():
    # N breaks windows, lower case folds, palindromes count once, chunk boundaries are not spanned
    c = py_oracle.count_kmers([("r", "ACGTNACGT")], 4)
    assert c == {"ACGT": 2}                                # ACGT is its own reverse complement
    c = py_oracle.count_kmers([("r", "acGTaC")], 3)
    assert c == {"ACG": 2, "GTA": 2}                       # ACG, CGT->ACG, GTA, TAC->GTA
    c = py_oracle.count_kmers([("r1", "AAAT"), ("r2", "TTAA")], 4)
    assert c == {"AAAT": 1, "TTAA": 1}
    c = py_oracle.count_kmers([("r", "T" * 40)], 32)
    assert c == {"A" * 32: 9}
    assert py_oracle.count_kmers([("r", "ACG")], 4) == {}  # chunk shorter than k
    assert py_oracle.dump_kcount({"AAA": 2, "AAC": 3}, 3) == [("AAC", 3)]          # mincount=3 for k<=31
    assert py_oracle.dump_kcount({"A" * 32: 1}, 32) == [("A" * 32, 1)]             # jellyfish dumps all


def test_kmer2fasta_thresholds_and_sampling():
    lines = [("AAA", 5), ("AAC", 10), ("AAG", 30), ("AAT", 31), ("ACA", 2000001)]
    assert py_oracle.kmer2fasta(lines, 10) == ["AAC", "AAG", "AAT", "ACA"]           # >= low (inclusive)
    assert py_oracle.kmer2fasta(lines, 10, high_bool=1, kmer_high_count=31) == ["AAC", "AAG", "AAT"]
    assert py_oracle.kmer2fasta(lines, 1, sampling_sensitivity=2) == ["AAA", "AAG", "ACA"]
    assert py_oracle.kmer2fasta(lines, 1, sampling_sensitivity=0) == [l[0] for l in lines]


def test_blast_either_strand_and_multiplicity():
    recs = [("c_0_8", "ACGAACGA"), ("c_8_16", "TCGTGGGG")]
    hits = py_oracle.blast_kmers(["ACGA"], recs)           # rc(ACGA) = TCGT
    assert sorted((h[2], h[3], h[1]) for h in hits) == [("c_0_8", 1, 0), ("c_0_8", 5, 0), ("c_8_16", 1, 16)]
    merged = py_oracle.blast2bed(hits)
    assert merged == [("c_0_8", 0, 8, "ACGA,ACGA"), ("c_8_16", 0, 8, "ACGA")]


@pytest.mark.parametrize("k,low", [(5, 8), (11, 3), (23, 4), (32, 2)])
def test_c_oracle_matches_literal_oracle(k, low):
    p = synth.small_params(seed=11 + k, genome_size=24_000, n_chrom=2)
    seq, off, names = synth.generate(p)
    chunk = 2500
    lit = py_oracle.run_path(util.to_sequences(seq, off, names), chunk, k, low)
    got = util.c_oracle_path(seq, off, names, chunk, k, low)
    # a2: every count, not only the dumped ones
    assert dict(zip(binding.kmer_decode(got["keys"], k), got["counts"].tolist())) == dict(lit["counts"])
    assert got["n_windows"] == sum(lit["counts"].values())
    # a3/a6: repeat set and its order
    assert binding.kmer_decode(got["sel"], k) == lit["kmers"]
    assert got["scaffolds"] == lit["scaffolds"]
    # a4-a7: the matrix, bit-exact
    want = sps.csr_matrix(lit["matrix"]); want.sort_indices()
    assert got["indptr"].tolist() == want.indptr.tolist()
    assert got["indices"].tolist() == want.indices.tolist()
    assert got["data"].tolist() == want.data.tolist()
    assert len(lit["kmers"]) > 0 and want.nnz > 0          # the case is not vacuous
    # a8
    util.assert_tfidf_close(got["tfidf"], lit["tfidf"].data)
