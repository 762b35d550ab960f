"""where the reference's real external tools are installed, pin the oracle's restatement of a2
(k-mer counting) and a4 (exact mapping) against them; where they are not (this image, the GPU boxes), say so and skip."""
import numpy as np
import pytest

from oracle import external_tools, py_oracle

TOOLS = external_tools.available()


def _small_genome():
    rng = np.random.default_rng(11)
    unit = "".join(rng.choice(list("ACGT"), 400))
    seqs = [("s%d" % i, "".join(rng.choice(list("ACGT"), 3000)) + unit * (i + 1) + "NNNNN" + "ACGT" * 20) for i in range(4)]
    return seqs


def _oracle_counts(seqs, k):
    return dict(py_oracle.count_kmers(seqs, k))                  # the literal oracle itself: what the tools must agree with


def test_tool_inventory_is_reported():
    """Always runs: states which of the reference's executables this machine has (none on the build image)."""
    assert set(TOOLS) == {"jellyfish", "kmercountexact.sh", "bbmap.sh"}
    print("external tools on PATH:", {k: bool(v) for k, v in TOOLS.items()})


@pytest.mark.skipif(not TOOLS["jellyfish"], reason="jellyfish is not on PATH: stage a2 (k > 31) stays restated from its documentation")
@pytest.mark.parametrize("k", [23, 32])
def test_oracle_counts_equal_jellyfish(k):
    seqs = _small_genome()
    assert external_tools.jellyfish_counts(seqs, k) == _oracle_counts(seqs, k)


@pytest.mark.skipif(not TOOLS["kmercountexact.sh"], reason="kmercountexact.sh (BBTools) is not on PATH: stage a2 (k <= 31) stays restated from its documentation")
@pytest.mark.parametrize("k", [15, 23, 31])
def test_oracle_counts_equal_kmercountexact(k):
    seqs = _small_genome()
    want = {kmer: c for kmer, c in _oracle_counts(seqs, k).items() if c >= 3}
    assert external_tools.kmercountexact_counts(seqs, k, mincount=3) == want


@pytest.mark.skipif(not TOOLS["bbmap.sh"], reason="bbmap.sh (BBTools) is not on PATH: stage a4 stays restated from its documentation")
def test_oracle_matches_equal_bbmap_perfect_mode():
    k = 23
    seqs = _small_genome()
    counts = _oracle_counts(seqs, k)
    kmers = sorted(kmer for kmer, c in counts.items() if c >= 3)[:200]
    hits = external_tools.bbmap_perfect_hits(seqs, [(km, km) for km in kmers], k)
    want = sorted((q, rname) for q, _, rname, _ in py_oracle.blast_kmers(kmers, seqs))
    assert hits == want
