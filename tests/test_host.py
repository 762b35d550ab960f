"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares, the host-side integer logic
matches the literal oracle, the device bit tricks (run on the host through the same inline routines) match the
oracle's k-mers, and compute calls fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import c_oracle, py_oracle
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import ChunkLayout
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "polycracker_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 35
    lib = ctypes.CDLL(binding._build.LIB)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "declared in the header but not exported: %s" % missing
    assert set(binding.EXPORTED) == set(declared), "binding.py and the header disagree"
    assert lib.pc_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if binding.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(binding.PolycrackerError, match="no CPU fallback"):
        binding.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "polycracker-unofficial-mirror_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/" not in text, f


@pytest.mark.parametrize("chunk", [5, 7, 50, 1000])
def test_chunk_table_matches_reference_geometry(chunk):
    fai = [("chrB", 150), ("chrA", 10), ("s_3", 11), ("tiny", 3), ("empty", 0), ("exact", 2 * chunk), ("plus1", 2 * chunk + 1)]
    names = [n for n, _ in fai]
    offsets = np.concatenate([[0], np.cumsum([l for _, l in fai])])
    lay = ChunkLayout(names, offsets, chunk)
    want = py_oracle.split_chunks(fai, chunk)
    assert lay.bed_rows() == want
    assert lay.original_scaffolds() == py_oracle.find_scaffolds(want)
    for kw in (dict(remove_non_chunk=1), dict(remove_non_chunk=0, min_chunk_threshold=1, min_chunk_size=4),
               dict(remove_non_chunk=0), dict(remove_non_chunk=1, prefilter=1)):
        scaffolds, chunk_row = lay.rows(**kw)
        assert scaffolds == py_oracle.filter_scaffolds(lay.original_scaffolds(), chunk, **kw)
        assert sorted(r for r in chunk_row.tolist() if r >= 0) == list(range(len(scaffolds)))
        assert [lay.chunk_names[i] for i in np.argsort(np.where(chunk_row < 0, 1 << 30, chunk_row))[:len(scaffolds)]] == scaffolds


def test_pack_word_swar():
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKMxX-*.0", dtype=np.uint8)
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    for _ in range(300):
        n = int(rng.integers(0, 33))
        b = alphabet[rng.integers(0, len(alphabet), n)]
        w, m = binding.host_pack_word(b.tobytes())
        for j in range(32):
            c = code.get(int(b[j]), None) if j < n else None
            assert ((m >> (31 - j)) & 1) == (c is None)
            assert ((w >> (62 - 2 * j)) & 3) == (c or 0)
    for byte in range(256):                                # every byte value, every lane position
        w, m = binding.host_pack_word(bytes([byte]) * 32)
        assert m == (0 if byte in code else 0xFFFFFFFF)


def test_kmer_codec_and_canonical():
    rng = np.random.default_rng(1)
    for k in (1, 2, 13, 23, 31, 32):
        kmers = ["".join(rng.choice(list("ACGT"), k)) for _ in range(50)] + ["A" * k, "T" * k, ("AT" * k)[:k]]
        keys = binding.kmer_encode(kmers, k)
        assert binding.kmer_decode(keys, k) == kmers
        assert sorted(kmers) == binding.kmer_decode(np.sort(keys), k)        # integer order == string order
        for km, key in zip(kmers, keys.tolist()):
            can = binding.host_canonical(key, k)
            assert binding.kmer_decode(np.array([can], np.uint64), k)[0] == py_oracle.canonical(km)
    assert binding.kmer_encode(["ACGN"], 4)[0] == np.uint64(0xFFFFFFFFFFFFFFFF)
    assert binding.kmer_encode(["acgt"], 4)[0] == binding.kmer_encode(["ACGT"], 4)[0]


@pytest.mark.parametrize("k", [1, 4, 15, 23, 26, 31, 32])
def test_device_roller_on_host_matches_oracle_counts(k):
    """k_pack's layout + the Roller (the code K3/K5 run per lane), executed on the host, must reproduce the
    oracle's k-mer multiset: N runs, lower case, chunk tails, chunk < k, padding words."""
    p = synth.small_params(seed=5 + k, genome_size=30_000, n_chrom=2, n_fraction=0.02)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 997)                    # odd chunk size: tails and unaligned chunk starts
    words, nmask, w0 = binding.host_pack_chunks(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    keys, valid = binding.host_extract(words, nmask, k)
    got_keys, got_counts = np.unique(keys[valid], return_counts=True)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    assert int(valid.sum()) == n_windows
    assert got_keys.tolist() == want_keys.tolist()
    assert got_counts.tolist() == want_counts.tolist()
    # windows never leave their chunk: the last k-1 starts of every chunk are invalid
    for c in range(lay.n_chunks):
        ln = int(lay.chunk_end[c] - lay.chunk_start[c])
        base = int(w0[c]) * 32
        assert not valid[base + max(0, ln - k + 1): int(w0[c + 1]) * 32].any()


def test_fasta_reader_and_writers(tmp_path):
    fa = tmp_path / "g.fa"
    fa.write_text(">s1 description here\nACGT\nacgtNN\n\n>s2\r\nTT\r\nGG\n>empty\n>s3\nA")
    seq, off, names = binding.read_fasta(str(fa))
    assert names == ["s1", "s2", "empty", "s3"]
    assert seq.tobytes() == b"ACGTacgtNNTTGGA" and off.tolist() == [0, 10, 14, 14, 15]
    keys = binding.kmer_encode(["ACG", "TTT"], 3)
    binding.write_kcount(str(tmp_path / "a.kcount"), keys, np.array([5, 7], np.uint32), 3)
    binding.write_kcount(str(tmp_path / "a.kcount"), keys[:1], np.array([9], np.uint32), 3, append=True)
    assert (tmp_path / "a.kcount").read_text() == "ACG\t5\nTTT\t7\nACG\t9\n"
    binding.write_kmer_fasta(str(tmp_path / "a.fa"), keys, 3)
    assert (tmp_path / "a.fa").read_text() == ">ACG\nACG\n>TTT\nTTT\n"
    binding.write_kmer_fasta(str(tmp_path / "b.fa"), keys, 3, rc_labels=True)
    assert (tmp_path / "b.fa").read_text() == ">CGT\nCGT\n>AAA\nAAA\n"


def test_synth_is_deterministic_and_shaped():
    p = synth.small_params(seed=3, genome_size=40_000)
    a, off, names = synth.generate(p)
    b, _, _ = synth.generate(p, threads=1)
    assert a.tobytes() == b.tobytes()
    assert names == ["SubA_chr1", "SubA_chr2", "SubB_chr1", "SubB_chr2"]
    one, _, _ = synth.generate(p, chroms=[2])              # any rank can make any chromosome on its own
    assert one.tobytes() == a[off[2]:off[3]].tobytes()
    text = a.tobytes()
    assert set(text) <= set(b"ACGTacgtNn") and b"N" in text and b"a" in text
