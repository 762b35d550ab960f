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


def test_partition_planning_and_count_floors():
    """Host integer logic around the count: bucket bits shared by all ranks (table count: ~64 MB sub-tables; shared-memory
    count: level 1 takes the smaller half of the bits), the counter's dump floor (kmercountexact mincount=3 for k <= 31,
    polycracker.py:199; everything for k > 31, polycracker.py:203) and the threshold the selection really applies."""
    from polycracker_b200 import distributed as pcd
    from polycracker_b200.pipeline import dump_floor, effective_low_count
    # table count: 5e8 windows / 0.6 * 12 B = 10 GB -> 2^8 sub-tables of 39 MB
    assert pcd.plan_partition_bits(500_000_000, 1) == 8
    assert pcd.plan_partition_bits(1000, 8) == 3                       # at least one bucket per rank
    # shared-memory count: windows / 4096 sub-buckets in total, 2 * 4^pb >= that
    for windows, want in ((500_000_000, 8), (1_000_000_000, 9), (4_000_000_000, 10), (4_460_000_000, 10)):
        pb = pcd.plan_partition_bits(windows, 1, smem_mean=4096)
        assert pb == want and (2 << (2 * pb)) >= windows // 4096 and (pb == 1 or (2 << (2 * (pb - 1))) < windows // 4096)
    assert pcd.plan_partition_bits(10_000, 4, smem_mean=4096) == 2
    assert [dump_floor(k) for k in (15, 23, 31, 32)] == [3, 3, 3, 1]
    assert effective_low_count(23, 1) == 3 and effective_low_count(23, 30) == 30 and effective_low_count(32, 1) == 1
    assert pcd.bind_to_gpu_numa(0) is None or isinstance(pcd.bind_to_gpu_numa(0), str)   # no NVML / GPU here: None


def test_n2_file_stage_writers(tmp_path):
    """kcount_hist / number_repeatmers_per_subsequence mirrors: the tables they write (plots only when matplotlib is there)."""
    from polycracker_b200 import stages
    kdir = tmp_path / "k"
    kdir.mkdir()
    (kdir / "23.kcount").write_text("ACG\t3\nCCC\t5\nGGA\t3\nTTT\t12\n")
    (kdir / "31.kcount").write_text("AAA 7\nAAC 7\n")
    (kdir / "notes.txt").write_text("x")
    out = str(tmp_path / "hist.png")
    tables = stages.kcount_hist(str(kdir), out_file=out, kcount_range="3,100", kmer_lengths="23")
    assert list(tables) == ["23.kcount"]
    v, f = tables["23.kcount"]
    assert v.tolist() == [3, 5, 12] and f.tolist() == [2, 1, 1]
    assert open(out + ".tsv").read().split("\n") == ["23.kcount\t3\t2", "23.kcount\t5\t1", "23.kcount\t12\t1"]
    tables = stages.kcount_hist(str(kdir), out_file=out)
    assert sorted(tables) == ["23.kcount", "31.kcount"] and tables["31.kcount"][1].tolist() == [2]
    bed = tmp_path / "blasted_merged.bed"
    bed.write_text("c_0_10\t0\t10\tAAA,CCC,AAA\nc_10_20\t0\t10\tGGG\n")
    got = stages.number_repeatmers_per_subsequence(str(bed), out_file=str(tmp_path / "per"))
    assert got.tolist() == [3, 1]
    assert open(str(tmp_path / "per.png.tsv")).read() == "1\t1\n3\t1"


@pytest.mark.parametrize("k,chunk", [(23, 20000), (31, 7777), (32, 5000), (17, 20000), (12, 3000), (26, 50000)])
def test_super_kmer_records_partition_the_kmers_exactly(k, chunk):
    """Host run of the device's own super-k-mer routines (kmer_skm.cuh: window validity, rolling m-mer hashes, sliding
    minimum, record boundaries, record assembly, rolling and direct expansion): the records hold exactly the multiset of
    canonical k-mers of the packed genome, every k-mer goes to ONE destination (so counting bucket by bucket is exact),
    destinations and record lengths respect the plan."""
    from polycracker_b200 import synth
    from polycracker_b200.pipeline import ChunkLayout
    p = synth.small_params(seed=11, genome_size=200_000)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, chunk)
    words, nmask, _ = binding.host_pack_chunks(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    keys, valid = binding.host_extract(words, nmask, k)
    want = np.sort(keys[valid])
    for nsub, pb, nb2, nmax in [(10, 1, 5, 16), (121000, 8, 473, 16), (1_090_000, 10, 1065, 8), (4000, 2, 2048, 1)]:
        m = binding.host_skm_m(k, nsub)
        assert 9 <= m <= 16 and (k - m + 1) % 2 == 0 and 4 <= k - m + 1 <= 22
        nmax = min(nmax, 49 - k + 1)
        recs = binding.host_skm_records(words, nmask, k, m, nmax, pb, nb2)
        got, dest = binding.host_skm_expand(recs, k)
        assert np.array_equal(np.sort(got), want)
        order = np.lexsort((dest, got))
        g, d = got[order], dest[order]
        same = g[1:] == g[:-1]
        assert np.all(d[1:][same] == d[:-1][same]), "a canonical k-mer went to two destinations"
        assert (d >> 11).max() < (1 << pb) and (d & 2047).max() < nb2
        n = (recs[:, 1] & np.uint64(63)).astype(np.int64) + 1
        assert n.max() <= nmax and n.sum() == len(want)
        if nmax == 1:
            assert len(recs) == len(want)
    assert binding.host_skm_m(9, 1000) == 0                     # too short for a minimizer window


def test_hipmer_output_to_kcount_matches_the_reference_awk_line(tmp_path):
    """format.py:123-133 shells out to `cat <files> | awk '{OFS="\\t"; sum=0; for (i=2; i<=7; i++) { sum+= $i }; if (sum >= 3)
    print $1, sum}'`; the host mirror must write the same bytes (run against the real awk when the box has one)."""
    import shutil
    import subprocess
    from polycracker_b200 import stages
    rng = np.random.default_rng(5)
    d = tmp_path / "hip"
    d.mkdir()
    for f in range(3):
        with open(d / ("part%d.txt" % f), "w") as out:
            for _ in range(400):
                kmer = "".join(rng.choice(list("ACGT"), 21))
                cols = rng.integers(0, 3, size=int(rng.integers(4, 9))).tolist()      # fewer or more than six count columns
                sep = "\t" if rng.random() < 0.5 else "  "
                out.write(kmer + sep + sep.join(str(c) for c in cols) + "\n")
            out.write("ACGTACGTACGTACGTACGTA\t2\t0.5\t0.5\tx\t0\t0\n")                     # fractional sum, a non-numeric field
            out.write("\n")
    stages.hipmer_output_to_kcount(str(d), str(tmp_path / "mine.kcount"), run_on_dir=True)
    stages.hipmer_output_to_kcount(str(d / "part0.txt"), str(tmp_path / "one.kcount"))
    lines = open(tmp_path / "mine.kcount").read().splitlines()
    assert lines and all(float(l.split("\t")[1]) >= 3 for l in lines)
    if shutil.which("awk"):
        files = " ".join(str(d / f) for f in os.listdir(d))
        cmd = "cat %s | awk '{OFS = \"\\t\"; sum=0; for (i=2; i<=7; i++) { sum+= $i }; if (sum >= 3) print $1, sum}' > %s" % (
            files, tmp_path / "awk.kcount")
        subprocess.check_call(cmd, shell=True)
        assert open(tmp_path / "mine.kcount").read() == open(tmp_path / "awk.kcount").read()
    k, keys, counts = stages.read_kcount(str(tmp_path / "mine.kcount"))
    assert k == 21 and len(keys) == len(np.unique(keys)) and counts.min() >= 3


def test_compact_handoff_rebuild_is_exact_for_every_packing():
    """Host side of the compact matrix hand-off (pipeline.rebuild_indices / rebuild_counts8): the three packings the device
    ships -- one-byte gaps + one-byte counts (pc_csr_get8), two-byte gaps (pc_csr_get_gap16), 12-bit gap + 4-bit count in one
    uint16 (pc_csr_get_g12c4) -- restated in numpy, rebuild the int32 columns and float32 counts exactly; the width rule
    (Context.gap_dtype) picks the wide form only for a large column universe."""
    from polycracker_b200.pipeline import rebuild_counts8, rebuild_indices
    rng = np.random.default_rng(3)
    n_rows, n_cols = 300, 100_000                               # mean column gap of a row: 500
    lens = rng.poisson(200, n_rows); lens[::7] = 0
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n_cols, l, replace=False)) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    data = rng.geometric(0.4, len(indices)).astype(np.float32)
    data[rng.random(len(data)) < 0.01] = 70000.0                   # counts beyond every field width
    first = np.zeros(len(indices), bool); first[indptr[:-1][lens > 0]] = True
    gap = np.where(first, 1 << 40, np.diff(indices.astype(np.int64), prepend=0))

    def pack(gap_esc, cnt_esc):
        g = np.minimum(gap, gap_esc); cn = np.minimum(data, cnt_esc)
        ce = np.flatnonzero(gap >= gap_esc); ke = np.flatnonzero(data >= cnt_esc)
        order = rng.permutation(len(ce))                        # the device appends escapes in no particular order
        return g, cn, (ce[order], indices[ce[order]]), (ke, data[ke])

    g, cn, ce, ke = pack(255, 255)
    assert np.array_equal(rebuild_indices(indptr, g.astype(np.uint8), ce), indices)
    assert np.array_equal(rebuild_counts8(cn.astype(np.uint8), ke), data)
    g, cn, ce, ke = pack(65535, 255)
    assert np.array_equal(rebuild_indices(indptr, g.astype(np.uint16), ce), indices)
    g, cn, ce, ke = pack(4095, 15)
    packed = ((g.astype(np.uint32) << 4) | cn.astype(np.uint32)).astype(np.uint16)
    assert np.array_equal(rebuild_indices(indptr, packed, ce, packed_counts=True), indices)
    assert np.array_equal(rebuild_counts8(None, ke, packed=packed), data)
    assert len(ce[0]) < 0.01 * len(indices) < 0.4 * len(indices) < (gap >= 255).sum()   # what the wide form is for
    assert binding.Context.gap_dtype(n_cols, n_rows, len(indices)) == np.uint16
    assert binding.Context.gap_dtype(673_070, 6621, 96_226_654) == np.uint8
    assert binding.Context.gap_dtype(10, 5, 0) == np.uint8
