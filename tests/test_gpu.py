"""GPU parity tests (run with -m gpu on the B200 box).  Every check goes through the C ABI
(polycracker_b200.binding -> libpolycracker_b200.so) and compares against the CPU checkers in oracle/.

Bars (north_star): k-mer counts, repeat set, CSR matrix bit-exact; TF-IDF within 1e-6 relative (fp32).
"""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import c_oracle, py_oracle
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import ChunkLayout, run_hot_path
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if binding.device_count() == 0:
        pytest.fail("no CUDA device: -m gpu tests need the B200 (there is no CPU fallback to test)")
    c = binding.Context(0)
    yield c
    c.close()


def _check_against_c_oracle(ctx, seq, off, names, chunk, k, low, **kw):
    res = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=low, **kw)
    want = util.c_oracle_path(seq, off, names, chunk, k, low,
                              high_bool=kw.get("high_bool", 0), high=kw.get("kmer_high_count", 2000000),
                              sampling=kw.get("sampling_sensitivity", 1),
                              remove_non_chunk=kw.get("remove_non_chunk", 1),
                              min_chunk_threshold=kw.get("min_chunk_threshold", 0),
                              min_chunk_size=kw.get("min_chunk_size", 0))
    assert res.stats["valid_windows"] == want["n_windows"]
    assert res.stats["distinct"] == len(want["keys"])
    assert res.kmer_keys.tolist() == want["sel"].tolist()
    assert res.scaffolds == want["scaffolds"]
    assert np.array_equal(res.indptr, want["indptr"])
    assert np.array_equal(res.indices, want["indices"])
    assert np.array_equal(res.data, want["data"])
    util.assert_tfidf_close(res.tfidf, want["tfidf"])
    return res, want


def test_known_answer_survey_8c(ctx):
    seq, off, names = util.from_sequences([("s1", "ACGTACGTAC")])
    res = run_hot_path(ctx, seq, off, names, chunk_size=5, k=3, kmer_low_count=3)
    assert res.kmers() == ["ACG", "GTA"] and res.kmers("max") == ["CGT", "TAC"]
    assert res.scaffolds == ["s1_0_5", "s1_5_10"]
    assert res.csr_matrix().toarray().tolist() == [[2, 1], [1, 2]]
    csc = res.csr_matrix().tocsc()
    assert csc.indptr.tolist() == [0, 2, 4] and csc.indices.tolist() == [0, 1, 0, 1] and csc.data.tolist() == [2, 1, 1, 2]
    np.testing.assert_allclose(res.csr_matrix("tfidf").toarray(), np.array([[2, 1], [1, 2]]) / np.sqrt(5), rtol=1e-6)
    res = run_hot_path(ctx, seq, off, names, chunk_size=5, k=3, kmer_low_count=4)
    assert res.shape == (2, 0) and len(res.indices) == 0 and res.indptr.tolist() == [0, 0, 0]


@pytest.mark.parametrize("chunk", [997, 1024, 75000])
def test_pack_matches_host_layout(ctx, chunk):
    p = synth.small_params(seed=21, genome_size=300_000, n_fraction=0.02)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, chunk)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    words, nmask, w0 = ctx.packed()
    hw, hm, h0 = binding.host_pack_chunks(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    assert np.array_equal(w0, h0)
    assert np.array_equal(words, hw[:-1]) and np.array_equal(nmask, hm[:-1])


def test_pack_every_byte_phase(ctx):
    # chunk starts at every offset mod 16 (the 128-bit load realignment) and every length mod 32
    rng = np.random.default_rng(5)
    seq = np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)[rng.integers(0, 10, 5000)].copy()
    begins = np.arange(0, 4000, 101, dtype=np.int64)
    ends = begins + (np.arange(len(begins)) % 67) + 30
    ctx.load_sequences(seq, begins, ends)
    words, nmask, w0 = ctx.packed()
    hw, hm, h0 = binding.host_pack_chunks(seq, begins, ends)
    assert np.array_equal(words, hw[:-1]) and np.array_equal(nmask, hm[:-1]) and np.array_equal(w0, h0)


@pytest.mark.parametrize("k", [1, 2, 5, 11, 16, 23, 26, 31, 32])
def test_counts_bit_exact(ctx, k):
    p = synth.small_params(seed=100 + k, genome_size=400_000, n_fraction=0.01)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 5000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    ctx.count(k)
    keys, counts = ctx.dump_counts(1)
    order = np.argsort(keys)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    assert np.array_equal(keys[order], want_keys)
    assert np.array_equal(counts[order].astype(np.uint64), want_counts)
    info = ctx.count_info()
    assert info["valid_windows"] == n_windows and info["distinct"] == len(want_keys)
    # the .kcount dump: mincount=3 (polycracker.py:199)
    k3, c3 = ctx.dump_counts(3)
    assert sorted(k3.tolist()) == want_keys[want_counts >= 3].tolist()


@pytest.mark.parametrize("k,mb", [(23, 1), (31, 4), (11, 1), (32, 2)])
def test_partitioned_count_equals_direct_and_oracle(ctx, k, mb):
    """The L2-partitioned count path (radix partition of the k-mers, one sub-table per bucket) forced on at a size
    where the direct path would normally run; both must reproduce the oracle bit for bit."""
    p = synth.small_params(seed=200 + k, genome_size=1_500_000, n_fraction=0.01)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 20000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    try:
        for mode in (2, 1, 0):
            ctx.tune(part_mode=mode, part_mbytes=mb)
            ctx.count(k)
            assert ctx.count_detail()["kind"] == ("list" if mode == 2 else "table")
            keys, counts = ctx.dump_counts(1)
            order = np.argsort(keys)
            assert np.array_equal(keys[order], want_keys), "mode %d" % mode
            assert np.array_equal(counts[order].astype(np.uint64), want_counts), "mode %d" % mode
            info = ctx.count_info()
            assert info["valid_windows"] == n_windows and info["distinct"] == len(want_keys)
        # insert variants of the bucket kernel (load-first / CAS-first, 2 or 4 k-mers in flight per thread)
        for casfirst, items in ((1, 4), (0, 2), (0, 4)):
            ctx.tune(part_mode=1, part_mbytes=mb, part_casfirst=casfirst, count_items=items)
            ctx.count(k)
            keys, counts = ctx.dump_counts(1)
            order = np.argsort(keys)
            assert np.array_equal(keys[order], want_keys), (casfirst, items)
            assert np.array_equal(counts[order].astype(np.uint64), want_counts), (casfirst, items)
        # the multi-GPU building blocks on one device: partition, then count the buckets as two "ranks" would
        # (bucket ranges [0, P/2) and [P/2, P), each fed by two segments per bucket)
        pb = 3
        for mode in (1, 2):
            ctx.tune(part_mode=mode, part_mbytes=mb, part_casfirst=0, count_items=2)
            off_b, keys_ptr = ctx.partition(k, pb)
            assert off_b[-1] == n_windows and np.all(np.diff(off_b) >= 0)
            got = {}
            for lo, hi in ((0, 4), (4, 8)):
                seg_begin = np.array([[off_b[q], off_b[q] + (off_b[q + 1] - off_b[q]) // 3] for q in range(lo, hi)])
                seg_len = np.array([[(off_b[q + 1] - off_b[q]) // 3, (off_b[q + 1] - off_b[q]) - (off_b[q + 1] - off_b[q]) // 3]
                                    for q in range(lo, hi)])
                ctx.count_buckets(k, pb, keys_ptr, seg_begin, seg_len)
                kk, cc = ctx.dump_counts(1)
                assert not (set(kk.tolist()) & set(got))                         # owners hold disjoint k-mers
                got.update(zip(kk.tolist(), cc.tolist()))
            assert got == dict(zip(want_keys.tolist(), want_counts.tolist())), "mode %d" % mode
        ctx.tune(part_mode=1, part_mbytes=mb)
        _check_against_c_oracle(ctx, seq, off, names, 20000, k, 5)
    finally:
        ctx.tune(part_mode=-1, part_mbytes=64, part_casfirst=0, count_items=2)


@pytest.mark.parametrize("k", [9, 23, 32])
def test_shared_memory_count(ctx, k):
    """The shared-memory count (two-level partition, one smem table per sub-bucket, compact list of survivors): every
    kernel variant, tiny and huge sub-buckets (the latter overflow their table and take the global-table fall-back),
    keep_min floors, all against the oracle."""
    p = synth.small_params(seed=300 + k, genome_size=2_000_000, n_fraction=0.01)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 20000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    distinct = len(want_keys)

    def check(keep_min, expect_overflow=None):
        ctx.count(k)
        d = ctx.count_detail()
        assert d["kind"] == "list" and d["keep_min"] == keep_min
        if expect_overflow is not None:
            assert (d["overflowed"] > 0) == expect_overflow, d
        keys, counts = ctx.dump_counts(keep_min)
        order = np.argsort(keys)
        m = want_counts >= keep_min
        assert d["kept"] == int(m.sum())
        assert np.array_equal(keys[order], want_keys[m]) and np.array_equal(counts[order].astype(np.uint64), want_counts[m])
        info = ctx.count_info()
        assert info["valid_windows"] == n_windows and info["distinct"] == distinct
        return d

    try:
        for variant in range(5):
            ctx.tune(part_mode=2, smem_variant=variant, scatter_variant=variant & 1, smem_target=0, keep_min=1)
            check(1, expect_overflow=False)
        ctx.tune(smem_variant=0, keep_min=3)
        check(3)
        with pytest.raises(binding.PolycrackerError, match="not kept"):
            ctx.dump_counts(2)
        k5, c5 = ctx.dump_counts(5, 9)
        m = (want_counts >= 5) & (want_counts <= 9)
        assert sorted(k5.tolist()) == want_keys[m].tolist()
        ctx.tune(smem_target=100, keep_min=2)                    # hundreds of tiny sub-buckets per bucket
        d = check(2, expect_overflow=False)
        assert d["sub_per_bucket"] > 50
        # sub-buckets far larger than a table: with enough distinct k-mers they overflow and are re-counted globally
        ctx.tune(smem_target=200_000, keep_min=1)
        check(1, expect_overflow=distinct > 100_000)
        ctx.tune(keep_min=4)
        check(4)
        ctx.tune(smem_target=0, keep_min=3)
        _check_against_c_oracle(ctx, seq, off, names, 20000, k, 5)
    finally:
        ctx.tune(part_mode=-1, smem_variant=0, scatter_variant=0, smem_target=0, keep_min=1)


@pytest.mark.parametrize("k", [12, 17, 23, 26, 31, 32])
def test_super_kmer_count(ctx, k):
    """The super-k-mer count (part_mode 3: minimizer-partitioned 16-byte records through both scatters, k-mers expanded
    inside the count kernel): record lengths, staging overflow, tiny and huge sub-buckets (the latter overflow their
    shared-memory table and take the record fall-back through a global table), keep_min floors, forced level-1 widths,
    all against the oracle."""
    p = synth.small_params(seed=500 + k, genome_size=2_000_000, n_fraction=0.01)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 20000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    distinct = len(want_keys)

    def check(keep_min, expect_overflow=None):
        ctx.count(k)
        d = ctx.count_detail()
        assert d["kind"] == "list" and d["keep_min"] == keep_min and d["skm_records"] > 0 and d["skm_m"] >= 9, d
        if expect_overflow is not None:
            assert (d["overflowed"] > 0) == expect_overflow, d
        keys, counts = ctx.dump_counts(keep_min)
        order = np.argsort(keys)
        m = want_counts >= keep_min
        assert d["kept"] == int(m.sum())
        assert np.array_equal(keys[order], want_keys[m]) and np.array_equal(counts[order].astype(np.uint64), want_counts[m])
        info = ctx.count_info()
        assert info["valid_windows"] == n_windows and info["distinct"] == distinct
        f = ctx.count_flags()                                    # uint32 counters: no counting unit anywhere near 2^32 instances
        assert not f["could_wrap"] and 0 < f["largest_unit"] <= n_windows and f["overflowed_sub_buckets"] == d["overflowed"]
        return d

    try:
        # (no "no overflow" expectation here: one minimizer inside a high-copy repeat family collects the mutated k-mers
        # of every copy, so a few sub-buckets outgrow their table on any repeat-rich genome and take the fall-back)
        ctx.tune(part_mode=3, smem_target=0, keep_min=1, skm_nmax=16, skm_stage=2560, smem_pb=0)
        d0 = check(1)
        for nmax in (1, 5, 32):                                  # one k-mer per record ... as long as a record can hold
            ctx.tune(skm_nmax=nmax)
            d = check(1)
            assert (d["skm_records"] == n_windows) if nmax == 1 else (d["skm_records"] < n_windows)
        ctx.tune(skm_nmax=16, skm_stage=256)                     # staging area too small: records leave one by one
        assert check(1)["skm_records"] == d0["skm_records"]
        ctx.tune(skm_stage=2560, keep_min=3)
        check(3)
        with pytest.raises(binding.PolycrackerError, match="not kept"):
            ctx.dump_counts(2)
        for pb in (1, 5, 11):                                    # 2 ... 2048 level-1 buckets
            ctx.tune(smem_pb=pb)
            assert check(3)["skm_records"] == d0["skm_records"]
        ctx.tune(smem_pb=0, smem_target=100, keep_min=2)         # hundreds of tiny sub-buckets per bucket
        d = check(2)
        assert d["sub_per_bucket"] > 50
        ctx.tune(smem_target=200_000, keep_min=1)                # sub-buckets far larger than a table: record fall-back
        check(1, expect_overflow=distinct > 100_000)
        ctx.tune(keep_min=4)
        check(4)
        ctx.tune(smem_target=0, keep_min=3)
        _check_against_c_oracle(ctx, seq, off, names, 20000, k, 5)
        assert ctx.count_detail()["skm_records"] > 0
    finally:
        ctx.tune(part_mode=-1, smem_target=0, keep_min=1, skm_nmax=16, skm_stage=2560, smem_pb=0)


def test_super_kmer_count_odd_shapes(ctx):
    """part_mode 3 on inputs that unbalance the minimizer buckets (poly-A, tandem repeats: every window of 300 kb shares
    ONE minimizer), on thousands of tiny scaffolds (chunks shorter than k, words that hold several chunks) and on a chunk
    with 2 % N; and too short a k is refused loudly."""
    rng = np.random.default_rng(3)
    tiny = [("t%d" % i, "".join(rng.choice(list("ACGT"), int(rng.integers(1, 90))))) for i in range(3000)]
    seqs = [("polyA", "A" * 300000), ("at", "AT" * 100000), ("unit", "ACGTTGCAAC" * 30000)] + tiny + \
           [("big", "".join(rng.choice(list("ACGTN"), 400000, p=[.245, .245, .245, .245, .02])))]
    seq, off, names = util.from_sequences(seqs)
    try:
        for target in (0, 300):
            ctx.tune(part_mode=3, smem_target=target)
            _check_against_c_oracle(ctx, seq, off, names, 50000, 23, 3, remove_non_chunk=0)
            assert ctx.count_detail()["skm_records"] > 0
            _check_against_c_oracle(ctx, seq, off, names, 1000000, 16, 3, remove_non_chunk=0)
            _check_against_c_oracle(ctx, seq, off, names, 1000000, 32, 3, remove_non_chunk=0)
        ctx.tune(smem_target=0)
        lay = ChunkLayout(names, off, 50000)
        ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
        with pytest.raises(binding.PolycrackerError, match="too short"):
            ctx.count(9)
    finally:
        ctx.tune(part_mode=-1, smem_target=0)


def test_small_table_reports_overflow_not_garbage(ctx):
    p = synth.small_params(seed=9, genome_size=100_000)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 5000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    with pytest.raises(binding.PolycrackerError, match="overflow"):
        ctx.count(23, capacity=1000)
    with pytest.raises(binding.PolycrackerError, match="unsupported"):
        ctx.count(33)


def test_full_path_vs_literal_oracle(ctx):
    p = synth.small_params(seed=31, genome_size=30_000, n_chrom=2)
    seq, off, names = synth.generate(p)
    for k, low in ((5, 8), (23, 4), (32, 2)):
        lit = py_oracle.run_path(util.to_sequences(seq, off, names), 2500, k, low)
        res = run_hot_path(ctx, seq, off, names, chunk_size=2500, k=k, kmer_low_count=low)
        assert res.kmers() == lit["kmers"] and res.scaffolds == lit["scaffolds"]
        want = sps.csr_matrix(lit["matrix"]); want.sort_indices()
        assert np.array_equal(res.indptr, want.indptr) and np.array_equal(res.indices, want.indices)
        assert np.array_equal(res.data, want.data)
        got_csc = res.csr_matrix().tocsc()                 # what clusteringMatrix.npz holds
        assert np.array_equal(got_csc.indptr, lit["matrix"].indptr) and np.array_equal(got_csc.indices, lit["matrix"].indices)
        assert np.array_equal(got_csc.data, lit["matrix"].data)
        util.assert_tfidf_close(res.tfidf, lit["tfidf"].data)      # sklearn TfidfTransformer itself
        assert want.nnz > 0


@pytest.mark.parametrize("k,low,chunk", [(23, 10, 75000), (26, 30, 50000), (15, 50, 10000), (31, 5, 33333), (32, 4, 75000)])
def test_full_path_vs_c_oracle(ctx, k, low, chunk):
    p = synth.small_params(seed=40 + k, genome_size=3_000_000, n_shared_families=6, n_specific_families=8,
                           max_family_len=1500)
    seq, off, names = synth.generate(p)
    res, want = _check_against_c_oracle(ctx, seq, off, names, chunk, k, low)
    assert len(want["sel"]) > 100 and len(want["data"]) > 1000
    # sklearn on the GPU-built matrix (the reference's own downstream call, polycracker.py:395-396)
    from sklearn.feature_extraction.text import TfidfTransformer
    sk = sps.csr_matrix(TfidfTransformer().fit_transform(res.csr_matrix().tocsc()))
    sk.sort_indices()
    util.assert_tfidf_close(res.tfidf, sk.data)


def test_thresholds_sampling_and_row_filters(ctx):
    p = synth.small_params(seed=77, genome_size=1_000_000)
    seq, off, names = synth.generate(p)
    _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 10, high_bool=1, kmer_high_count=40)
    _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 10, sampling_sensitivity=7)
    _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 1)                          # mincount=3 floor
    _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 10, remove_non_chunk=0)     # tails are rows too
    _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 10, remove_non_chunk=0, min_chunk_threshold=1,
                            min_chunk_size=15000)
    res, _ = _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 10**6)             # nothing selected
    assert res.shape[1] == 0


def test_degenerate_inputs(ctx):
    # all N, chunk shorter than k, one-base tail, empty sequence in the middle, no rows at all
    seqs = [("n", "N" * 100), ("short", "ACGTACG"), ("e", ""), ("r", "ACGTTGCA" * 20 + "A")]
    seq, off, names = util.from_sequences(seqs)
    for chunk, kw in ((40, {}), (40, dict(remove_non_chunk=0)), (1000, {}), (1000, dict(remove_non_chunk=0))):
        lit = py_oracle.run_path(seqs, chunk, 8, 3, **kw)
        res = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=8, kmer_low_count=3, **kw)
        assert res.kmers() == lit["kmers"] and res.scaffolds == lit["scaffolds"]
        want = sps.csr_matrix(lit["matrix"]); want.sort_indices()
        assert np.array_equal(res.indptr, want.indptr) and np.array_equal(res.indices, want.indices)
        assert np.array_equal(res.data, want.data)
        util.assert_tfidf_close(res.tfidf, sps.csr_matrix(lit["tfidf"]).data)
    seq, off, names = util.from_sequences([("e", "")])
    res = run_hot_path(ctx, seq, off, names, chunk_size=10, k=5, kmer_low_count=3)
    assert res.shape == (0, 0)


def test_low_complexity_heavy_hitters(ctx):
    # poly-A / tandem repeats: thousands of atomics on one slot, long runs in the row sort
    seqs = [("a", "A" * 70000), ("t", "T" * 70000), ("at", "AT" * 35000), ("mix", ("ACGTTGCAAC" * 7000))]
    seq, off, names = util.from_sequences(seqs)
    _check_against_c_oracle(ctx, seq, off, names, 10000, 23, 3)
    _check_against_c_oracle(ctx, seq, off, names, 70000, 32, 3)


def test_partitioned_path_skewed_buckets_and_odd_shapes(ctx):
    """Forced partitioned count + stream probe on inputs that unbalance the buckets (poly-A / tandem repeats put
    hundreds of thousands of instances of ONE k-mer into one bucket), on thousands of tiny scaffolds (every warp tile
    spans many chunks) and on one chunk longer than any tile."""
    rng = np.random.default_rng(3)
    tiny = [("t%d" % i, "".join(rng.choice(list("ACGT"), int(rng.integers(1, 90))))) for i in range(3000)]
    seqs = [("polyA", "A" * 300000), ("at", "AT" * 100000), ("unit", "ACGTTGCAAC" * 30000)] + tiny + \
           [("big", "".join(rng.choice(list("ACGTN"), 400000, p=[.245, .245, .245, .245, .02])))]
    seq, off, names = util.from_sequences(seqs)
    try:
        for mb in (1, 4):
            ctx.tune(part_mode=1, part_mbytes=mb)
            _check_against_c_oracle(ctx, seq, off, names, 50000, 23, 3, remove_non_chunk=0)
            _check_against_c_oracle(ctx, seq, off, names, 1000000, 16, 3, remove_non_chunk=0)
        for target in (0, 300):                              # the shared-memory count on the same inputs
            ctx.tune(part_mode=2, smem_target=target)
            _check_against_c_oracle(ctx, seq, off, names, 50000, 23, 3, remove_non_chunk=0)
            _check_against_c_oracle(ctx, seq, off, names, 1000000, 16, 3, remove_non_chunk=0)
        ctx.tune(smem_target=0)
    finally:
        ctx.tune(part_mode=-1, part_mbytes=64)


def test_probe_filters(ctx):
    """K5 with no filter, the shared-memory bitmap, the L2-resident bitmap (large repeat sets) and the fingerprint walk,
    roller and stream variants: identical matrices."""
    p = synth.small_params(seed=88, genome_size=4_000_000, n_shared_families=6, n_specific_families=8, max_family_len=1500)
    seq, off, names = synth.generate(p)
    try:
        for kw in (dict(probe_filter=0), dict(probe_filter=1), dict(probe_filter=1, filter_bits=4096),
                   dict(probe_filter=2), dict(probe_filter=2, probe_fp=1), dict(probe_filter=0, probe_fp=1),
                   dict(probe_filter=2, probe_mode=0)):
            ctx.tune(part_mode=2, probe_mode=-1, probe_fp=-1, filter_bits=1310720)
            ctx.tune(**kw)
            res, want = _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 5)
            assert len(want["sel"]) > 1000
        # no k-mer stream (super-k-mer count): the roller probe with the shared-memory bitmap, the L2 bitmap, the compact
        # 8-byte-slot repeat table (verified against the key array), the fingerprint walk, and the plain round-1 kernel
        for kw in (dict(), dict(probe_table8=1), dict(probe_filter=2), dict(probe_filter=2, probe_table8=0),
                   dict(probe_filter=0, probe_table8=1), dict(probe_filter=0, probe_table8=0, probe_fp=1),
                   dict(probe_filter=1, filter_bits=4096, probe_table8=1), dict(probe_roll=0)):
            ctx.tune(part_mode=3, probe_mode=-1, probe_fp=-1, filter_bits=1310720, probe_filter=-1, probe_table8=-1, probe_roll=1)
            ctx.tune(**kw)
            res, want = _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 5)
            assert ctx.count_detail()["skm_records"] > 0
    finally:
        ctx.tune(part_mode=-1, probe_filter=-1, probe_fp=-1, probe_mode=-1, filter_bits=1310720, probe_table8=-1, probe_roll=1)


def test_rows_to_csr_bitmap_equals_sort_and_scipy(ctx):
    """K6: the shared-memory column-bitmap kernel (kernels_csr.cuh) against the per-row radix sort and against scipy, on
    raw hit lists (pc_matrix_from_hits): empty rows, one heavy column, more rows than resident CTAs x 32 (the chained scan
    looks back over several windows), several column ranges per row (csr_range_bits) and rows with more distinct columns
    than shared count slots (csr_ncap: the counts of the tail live in global memory)."""
    rng = np.random.default_rng(7)
    for n_rows, n_cols, mean_hits in ((1, 1, 5), (700, 5000, 300), (5200, 40000, 40), (300, 2_500_000, 3000)):
        lens = rng.poisson(mean_hits, n_rows).astype(np.int64)
        lens[rng.random(n_rows) < 0.1] = 0
        ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        cols = rng.integers(0, n_cols, int(ro[-1])).astype(np.int32)
        cols[rng.random(len(cols)) < 0.3] = n_cols // 2            # a heavy column: counts far above 1
        if len(cols) > 10:
            cols[:5] = n_cols - 1                                  # the last column of the last range
        rows = np.repeat(np.arange(n_rows), lens)
        want = sps.coo_matrix((np.ones(len(cols), np.float32), (rows, cols)), shape=(n_rows, n_cols)).tocsr()
        want.sort_indices()
        got = {}
        try:
            for name, kw in (("sort", dict(csr_mode=0)), ("bitmap", dict(csr_mode=1)),
                             ("ranges", dict(csr_mode=1, csr_range_bits=1024)),
                             ("ncap", dict(csr_mode=1, csr_ncap=3)),
                             ("ranges+ncap", dict(csr_mode=1, csr_range_bits=2048, csr_ncap=1))):
                if name.startswith("ranges") and n_cols > 64 * 2048:
                    continue                                       # more ranges than the kernel takes (it would fall back to the sort)
                ctx.tune(csr_mode=-1, csr_range_bits=0, csr_ncap=0)
                ctx.tune(**kw)
                nnz = ctx.matrix_from_hits(ro, cols, n_cols)
                got[name] = ctx.csr(n_rows, nnz)
                df = ctx.df(n_cols)
                assert nnz == want.nnz, name
                assert np.array_equal(got[name][0], want.indptr), name
                assert np.array_equal(got[name][1], want.indices), name
                assert np.array_equal(got[name][2], want.data), name
                assert np.array_equal(df, np.bincount(want.indices, minlength=n_cols)), name
        finally:
            ctx.tune(csr_mode=-1, csr_range_bits=0, csr_ncap=0)
    with pytest.raises(Exception):
        ctx.matrix_from_hits(np.array([0, 1]), np.array([7], np.int32), 7)       # column out of range


def test_stage_timeline_and_raw_select(ctx):
    """pc_stage_timeline: the stages of the last step in start order, each as long as pc_timings says, inside the step;
    pc_tune select_raw (what multi-GPU owners use): the survivors leave unsorted, the matrix refuses to build on them until
    the (gathered) set is installed with pc_set_selected -- and then equals the plain path's."""
    p = synth.small_params(seed=91, genome_size=1_500_000)
    seq, off, names = synth.generate(p)
    ctx.reset_counters()
    res = run_hot_path(ctx, seq, off, names, chunk_size=20000, k=23, kmer_low_count=5)
    tl = ctx.stage_timeline()
    tm = ctx.timings()
    assert [n for n, _, _ in tl][:2] == ["h2d", "pack"] and {"count", "probe", "row_sort", "tfidf"} <= {n for n, _, _ in tl}
    assert all(b >= a >= 0 for _, a, b in tl) and all(tl[i][1] <= tl[i + 1][1] for i in range(len(tl) - 1))
    for name, a, b in tl:
        assert abs((b - a) - tm[name]) < 0.05, (name, a, b, tm[name])
    lay = ChunkLayout(names, off, 20000)
    scaffolds, chunk_row = lay.rows()
    try:
        ctx.tune(select_raw=1)
        n = ctx.select(5, 2000000, False, 1)
        raw = ctx.selected(n)
        assert n == len(res.kmer_keys) and not np.array_equal(raw, res.kmer_keys) and np.array_equal(np.sort(raw), res.kmer_keys)
        with pytest.raises(binding.PolycrackerError, match="select_raw"):
            ctx.build_matrix(chunk_row, len(scaffolds))
        ctx.tune(select_raw=0)
        assert ctx.set_selected(raw, 23) == n
        nnz = ctx.build_matrix(chunk_row, len(scaffolds))
        indptr, indices, data = ctx.csr(len(scaffolds), nnz)
        assert np.array_equal(indptr, res.indptr) and np.array_equal(indices, res.indices) and np.array_equal(data, res.data)
    finally:
        ctx.tune(select_raw=0)


def test_compact_handoff_matches_full_arrays(ctx):
    """What crosses PCIe in the end-to-end path (pc_csr_get8: one byte per column gap and per count + escapes, 2 bytes per
    entry; or pc_csr_get16: 6 bytes; both with pc_tfidf_factors_get) rebuilds on the host to exactly the int32 / float32 arrays
    of pc_csr_get / pc_tfidf_get (12 bytes per entry) -- bit for bit, including a count >= 65535 (a poly-A chunk of 70 000
    windows), column gaps >= 255, and an all-N row with norm 0."""
    p = synth.small_params(seed=77, genome_size=600_000)
    seq, off, names = synth.generate(p)
    seqs = util.to_sequences(seq, off, names) + [("polyA", "A" * 70000), ("allN", "N" * 70000), ("at", "AT" * 35000)]
    seq, off, names = util.from_sequences(seqs)
    for k, chunk in ((23, 70000), (31, 20000)):
        full = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=5, remove_non_chunk=0, compact=False)
        comp = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=5, remove_non_chunk=0, compact=16)
        c8 = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=5, remove_non_chunk=0, compact=True)
        assert c8.dcol8.dtype == np.uint8 and c8.cnt8.dtype == np.uint8 and c8._indices is None and c8._data is None
        assert np.array_equal(c8.indptr, full.indptr) and np.array_equal(c8.indices, full.indices) and c8.indices.dtype == np.int32
        assert np.array_equal(c8.data, full.data) and np.array_equal(c8.tfidf.view(np.uint32), full.tfidf.view(np.uint32))
        assert len(c8.escapes8[0][0]) >= (np.diff(full.indptr) > 0).sum()   # every non-empty row starts with an absolute column
        assert (np.asarray(c8.cnt8) == 255).sum() == len(c8.escapes8[1][0]) == (full.data >= 255).sum()
        # two-byte column gaps (pc_csr_get_gap16: what gap_dtype picks for a large column universe), forced here
        keep = binding.Context.__dict__["gap_dtype"]             # (the staticmethod object, not the function it wraps)
        try:
            binding.Context.gap_dtype = staticmethod(lambda n_sel, n_rows, nnz: np.uint16)
            c16 = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=5, remove_non_chunk=0, compact=True)
        finally:
            binding.Context.gap_dtype = keep
        # ... which ships gap and count in ONE uint16 (pc_csr_get_g12c4: gap << 4 | count, 4095 / 15 escape)
        assert c16.dcol8.dtype == np.uint16 and c16.cnt8 is None and c16._indices is None and c16._data is None
        assert np.array_equal(c16.indices, full.indices) and np.array_equal(c16.data, full.data)
        assert np.array_equal(c16.tfidf.view(np.uint32), full.tfidf.view(np.uint32))
        assert len(c16.escapes8[0][0]) <= len(c8.escapes8[0][0]) and ((np.asarray(c16.dcol8) >> 4) == 4095).sum() == len(c16.escapes8[0][0])
        assert ((np.asarray(c16.dcol8) & 15) == 15).sum() == len(c16.escapes8[1][0]) == (full.data >= 15).sum()
        # two-byte gaps next to one-byte counts (pc_csr_get_gap16), straight through the binding on the matrix just built
        from polycracker_b200.pipeline import rebuild_counts8, rebuild_indices
        nnz = len(full.indices)
        g16, k8, ip = np.empty(nnz, np.uint16), np.empty(nnz, np.uint8), np.empty(len(full.indptr), np.int64)
        n_col, n_cnt = ctx.csr8_into(ip, g16, k8)
        esc = ctx.csr8_escapes(n_col, n_cnt)
        assert np.array_equal(ip, full.indptr) and np.array_equal(rebuild_indices(ip, g16, esc[0]), full.indices)
        assert np.array_equal(rebuild_counts8(k8, esc[1]), full.data) and (g16 == 65535).sum() == n_col
        assert binding.Context.gap_dtype(5_600_000, 6600, 143_000_000) == np.uint16      # the 8-GPU weak-scaling shape
        assert binding.Context.gap_dtype(673_070, 6621, 96_226_654) == np.uint8          # the bench workload
        assert comp.data16 is not None and comp.data16.dtype == np.uint16 and full.data16 is None
        assert np.array_equal(comp.indptr, full.indptr) and np.array_equal(comp.indices, full.indices)
        assert np.array_equal(comp.kmer_keys, full.kmer_keys) and np.array_equal(comp.df, full.df)
        assert comp.data.dtype == np.float32 and np.array_equal(comp.data, full.data)
        assert comp.tfidf.dtype == np.float32 and np.array_equal(comp.tfidf.view(np.uint32), full.tfidf.view(np.uint32))
        if chunk == 70000:
            assert comp.escapes is not None and len(comp.escapes[0]) >= 1 and full.data.max() >= 65535
            assert (np.asarray(comp.data16) == 65535).sum() == len(comp.escapes[0])
        assert (comp.row_norm == 0).any()                              # the all-N chunks: empty rows stay zero


def test_external_kcount_ingest(ctx, tmp_path):
    """Scope row N4, second half (format.py:123-133): a .kcount written by ANOTHER counter (here: the oracle's counts, with
    half of the k-mers named by their reverse-complement strand, as a non-canonical counter would) is installed as the count
    result (pc_load_counts) and K4..K7 run on it: same repeat set, matrix and TF-IDF as counting on the GPU."""
    from polycracker_b200 import stages
    p = synth.small_params(seed=123, genome_size=400_000)
    seq, off, names = synth.generate(p)
    k, chunk, low = 23, 20000, 6
    want = util.c_oracle_path(seq, off, names, chunk, k, low)
    keep = want["counts"] >= 3
    kmers = binding.kmer_decode(want["keys"][keep], k)
    comp = str.maketrans("ACGT", "TGCA")
    path = str(tmp_path / "ext.kcount")
    with open(path, "w") as f:
        for i, (km, c) in enumerate(zip(kmers, want["counts"][keep].tolist())):
            f.write("%s%s%d\n" % (km.translate(comp)[::-1] if i % 2 else km, "\t" if i % 3 else " ", c))
        f.write("ACGTNNNNNNNNNNNNNNNNNNN\t50\n")                      # not a k-mer: skipped
    lay = ChunkLayout(names, off, chunk)
    scaffolds, chunk_row = lay.rows()
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    kk, n = stages.load_kcount(ctx, path)
    assert kk == k and n == int(keep.sum()) and ctx.count_info()["distinct"] == n
    n_sel = ctx.select(low)
    assert ctx.selected(n_sel).tolist() == want["sel"].tolist()
    nnz = ctx.build_matrix(chunk_row, len(scaffolds))
    indptr, indices, data = ctx.csr(len(scaffolds), nnz)
    assert np.array_equal(indptr, want["indptr"]) and np.array_equal(indices, want["indices"]) and np.array_equal(data, want["data"])
    ctx.tfidf(len(scaffolds))
    util.assert_tfidf_close(ctx.tfidf_values(nnz), want["tfidf"])
    values, freq, above = ctx.count_histogram(3, 1000)
    assert freq.sum() + above == n
    with pytest.raises(binding.PolycrackerError, match="more than once"):
        ctx.load_counts(np.array([5, 5], np.uint64), np.array([3, 4], np.uint32), k)
    with pytest.raises(binding.PolycrackerError, match="not k-mers"):
        ctx.load_counts(np.array([5, 0xFFFFFFFFFFFFFFFF], np.uint64), np.array([3, 4], np.uint32), k)


def test_n2_diagnostics(ctx):
    """Scope row N2: the data of kcount_hist (Counter over the dumped counts) and of number_repeatmers_per_subsequence
    (k-mer names per merged-bed line), from the table count, the list count and against the literal oracle."""
    from collections import Counter
    p = synth.small_params(seed=91, genome_size=60_000, n_chrom=2)
    seq, off, names = synth.generate(p)
    k, chunk, low = 11, 2500, 4
    lit = py_oracle.run_path(util.to_sequences(seq, off, names), chunk, k, low, remove_non_chunk=0)
    want_hist = py_oracle.kcount_histogram(["%s\t%d" % kc for kc in lit["kcount"]])
    per_line = dict(zip([m[0] for m in lit["merged"]],
                        py_oracle.repeatmers_per_subsequence(["\t".join(map(str, m)) for m in lit["merged"]])))
    try:
        for mode in (0, 2):
            ctx.tune(part_mode=mode, keep_min=3)
            res = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=low, remove_non_chunk=0)
            ctx.tune(keep_min=3)
            ctx.count(k)
            vals, freq, above = ctx.count_histogram(3, 50)
            want = [(c, n) for c, n in want_hist if c <= 50]
            assert list(zip(vals.tolist(), freq.tolist())) == want and above == sum(n for c, n in want_hist if c > 50)
            ctx.select(low)
            lay = res.layout
            scaffolds, chunk_row = lay.rows(remove_non_chunk=0)
            ctx.build_matrix(chunk_row, len(scaffolds))
            hits = ctx.row_hits(len(scaffolds))
            assert hits.tolist() == [per_line.get(s_, 0) for s_ in scaffolds]
            assert hits.tolist() == np.asarray(res.csr_matrix().sum(axis=1)).ravel().astype(np.int64).tolist()
        with pytest.raises(binding.PolycrackerError, match="not kept"):
            ctx.count_histogram(1, 10)
    finally:
        ctx.tune(part_mode=-1, keep_min=1)
    # larger: against the C oracle's counts
    p = synth.small_params(seed=92, genome_size=3_000_000)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 20000)
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    ctx.count(23)
    vals, freq, above = ctx.count_histogram(1, 200)
    _, want_counts, _ = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, 23)
    cnt = Counter(want_counts.tolist())
    assert dict(zip(vals.tolist(), freq.tolist())) == {c: n for c, n in cnt.items() if c <= 200}
    assert above == sum(n for c, n in cnt.items() if c > 200)


def test_n3_standard_scaler(ctx):
    """Scope row N3: length normalisation (polycracker.py:357-359) + sklearn StandardScaler(with_mean=False)
    (polycracker.py:392-393) on the GPU-built matrix, against sklearn itself; tolerance 1e-6 relative (float32)."""
    from sklearn.preprocessing import StandardScaler
    p = synth.small_params(seed=93, genome_size=3_000_000, n_shared_families=6, n_specific_families=8, max_family_len=1500)
    seq, off, names = synth.generate(p)
    res = run_hot_path(ctx, seq, off, names, chunk_size=20000, k=23, kmer_low_count=5, remove_non_chunk=0)
    m = res.csr_matrix()
    assert m.nnz > 10000
    # raw counts
    got = ctx.standard_scale(m.nnz)
    want = sps.csr_matrix(StandardScaler(with_mean=False).fit_transform(m.tocsc())); want.sort_indices()
    assert np.array_equal(want.indices, res.indices)
    util.assert_tfidf_close(got, want.data)
    # after the tfidf = 0 length normalisation of generate_Kmer_Matrix
    lens = dict(zip(res.layout.chunk_names, np.abs(res.layout.chunk_end - res.layout.chunk_start).tolist()))
    x = np.array([lens[s_] / 5000. for s_ in res.scaffolds], np.float64)
    rs = 1.0 / x
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    ln = sps.csr_matrix(((m.data.astype(np.float64) * rs[rows]).astype(np.float32), m.indices, m.indptr), shape=m.shape)
    got = ctx.standard_scale(m.nnz, rs)
    want = sps.csr_matrix(StandardScaler(with_mean=False).fit_transform(ln.tocsc())); want.sort_indices()
    util.assert_tfidf_close(got, want.data)
    # a constant column (every row holds the same value) keeps its values: scale 1
    seqs = [("c%d" % i, "ACGTTGCAACGGTTAA" * 4) for i in range(6)]
    s2, o2, n2 = util.from_sequences(seqs)
    r2 = run_hot_path(ctx, s2, o2, n2, chunk_size=64, k=8, kmer_low_count=3)
    m2 = r2.csr_matrix()
    got2 = ctx.standard_scale(m2.nnz)
    want2 = sps.csr_matrix(StandardScaler(with_mean=False).fit_transform(m2.tocsc())); want2.sort_indices()
    util.assert_tfidf_close(got2, want2.data)
    assert m2.nnz > 0 and np.array_equal(got2, m2.data)


def test_n3_gram_matrix(ctx):
    """Scope row N3, second half: K = X X^T for KernelPCA(kernel='linear' | 'cosine') (polycracker.py:387, 398-399) from the
    matrix on the device, against sklearn's linear_kernel / cosine_similarity on the same sparse TF-IDF matrix."""
    from sklearn.metrics.pairwise import cosine_similarity, linear_kernel
    from polycracker_b200 import gram
    p = synth.small_params(seed=42, genome_size=1_500_000)
    seq, off, names = synth.generate(p)
    seqs = util.to_sequences(seq, off, names) + [("allN", "N" * 30000)]
    seq, off, names = util.from_sequences(seqs)
    res = run_hot_path(ctx, seq, off, names, chunk_size=15000, k=23, kmer_low_count=5, remove_non_chunk=0, compact=False)
    X = res.csr_matrix("tfidf")
    assert X.shape[0] > 80 and X.shape[1] > 1000
    want = linear_kernel(X)
    for panel in (257, 8192):
        got = gram.gram_matrix_numpy(ctx, X.shape[0], X.shape[1], values="tfidf", kernel="linear", panel_cols=panel, engine="cublas")
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6)
    got = gram.gram_matrix_numpy(ctx, X.shape[0], X.shape[1], values="tfidf", kernel="linear", tf32=True, engine="cublas")
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-3)                  # TF32 products: <= 1e-3 of the unit diagonal
    # the hand-written tcgen05 kernel (TF32 operands rounded to nearest, float32 accumulation in TMEM): 1e-3 relative;
    # panel widths that leave a partial K block, a partial last panel, one panel for everything
    # (gram_variant 0: one CTA per 128 x 128 tile; 1, the default: persistent CTAs, 128 x 256 tiles, two TMEM accumulators)
    try:
        for variant in (0, 1):
            ctx.tune(gram_variant=variant)
            for panel in (100, 1000, 4096, 1 << 20):
                got = gram.gram_matrix_numpy(ctx, X.shape[0], X.shape[1], values="tfidf", kernel="linear", panel_cols=panel)
                np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-5)
                assert np.array_equal(got, got.T)
    finally:
        ctx.tune(gram_variant=1)
    C = res.csr_matrix("counts")
    got = gram.gram_matrix_numpy(ctx, C.shape[0], C.shape[1], values="counts", kernel="cosine")
    np.testing.assert_allclose(got, cosine_similarity(C), rtol=1e-3, atol=1e-6)
    assert np.all(got[-1] == 0)                                               # the all-N chunk: an empty row stays zero
    try:
        # several tiles per persistent CTA (3 000 rows: 156 tiles of 128 x 256 on 148 SMs, so the two accumulators alternate),
        # ragged last row / column blocks, small integer counts: TF32 products and float32 sums are exact here
        rng = np.random.default_rng(5)
        n_rows, n_cols = 3000, 5000
        lens = rng.poisson(40, n_rows).astype(np.int64); lens[::11] = 0
        ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        cols = rng.integers(0, n_cols, int(ro[-1])).astype(np.int32)
        nnz = ctx.matrix_from_hits(ro, cols, n_cols)
        ip, ix, dt = ctx.csr(n_rows, nnz)
        M = sps.csr_matrix((dt, ix, ip), shape=(n_rows, n_cols))
        exact = (M @ M.T).toarray().astype(np.float32)
        for variant in (0, 1):
            ctx.tune(gram_variant=variant)
            for panel in (512, 5000):
                got = gram.gram_matrix_numpy(ctx, n_rows, n_cols, values="counts", kernel="linear", panel_cols=panel)
                assert np.array_equal(got, exact), (variant, panel, np.abs(got - exact).max())
    finally:
        ctx.tune(gram_variant=1)


def test_threshold_watch_fused_select(ctx):
    """pc_watch_threshold (K4 fused into the table count through returning atomics on threshold crossings): same repeat
    set and matrix as the scan; a no-op for the shared-memory count, whose list already is the thresholded dump."""
    p = synth.small_params(seed=95, genome_size=2_000_000, n_shared_families=6, n_specific_families=8, max_family_len=1500)
    seq, off, names = synth.generate(p)
    try:
        for mode in (1, 2, 0):
            ctx.tune(part_mode=mode, part_mbytes=1)
            _check_against_c_oracle(ctx, seq, off, names, 20000, 23, 7, fuse_select=True)
        ctx.watch_threshold(0)
    finally:
        ctx.tune(part_mode=-1, part_mbytes=64)
        ctx.watch_threshold(0)


def test_overlap_pipeline_and_sm_copy():
    """StepPipeline (H2D of the next genome / D2H of the previous result overlapped with this step's kernels) returns, one
    call late, exactly what a plain call returns -- for a series of DIFFERENT genomes, with copy-engine and with SM-driven
    bulk copies; pc_copy_async moves bytes faithfully in both directions, odd sizes included."""
    import torch
    from polycracker_b200.overlap import StepPipeline
    dev = torch.device("cuda", 0)
    genomes = []
    for i in range(4):
        p = synth.small_params(seed=400 + i, genome_size=600_000 + 100_003 * i)
        seq, off, names = synth.generate(p)
        genomes.append((torch.from_numpy(seq).pin_memory(), seq, off, names))
    with binding.Context(0) as plain:
        want = [run_hot_path(plain, seq, off, names, chunk_size=20000, k=23, kmer_low_count=5) for _, seq, off, names in genomes]
    keep = binding.Context.__dict__["gap_dtype"]
    for ctas, wide in ((0, False), (8, False), (0, True)):     # wide: the packed two-byte hand-off of a large column universe
      binding.Context.gap_dtype = staticmethod(lambda n_sel, n_rows, nnz: np.uint16) if wide else keep
      try:
        with binding.Context(0) as c2:
            pipe = StepPipeline(c2, dev, sm_copy_ctas=ctas)
            got = []
            pipe.prefetch(genomes[0][0])
            for i, (_, seq, off, names) in enumerate(genomes):
                if i + 1 < len(genomes):
                    pipe.prefetch(genomes[i + 1][0])
                prev = pipe.step(lambda d, off=off, names=names: run_hot_path(c2, d, off, names, chunk_size=20000, k=23,
                                                                              kmer_low_count=5, fetch=False))
                if prev is not None:
                    got.append((prev.kmer_keys.copy(), prev.indptr.copy(), prev.indices.copy(), prev.data.copy(),
                                prev.tfidf.copy(), prev.df.copy(), prev.scaffolds))
            last = pipe.flush()
            got.append((last.kmer_keys.copy(), last.indptr.copy(), last.indices.copy(), last.data.copy(), last.tfidf.copy(),
                        last.df.copy(), last.scaffolds))
            pipe.close()
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert np.array_equal(g[0], w.kmer_keys) and np.array_equal(g[1], w.indptr) and np.array_equal(g[2], w.indices)
                assert np.array_equal(g[3], w.data) and np.array_equal(g[4], w.tfidf) and np.array_equal(g[5], w.df)
                assert g[6] == w.scaffolds and len(w.data) > 1000
            # pc_copy_async on its own
            for n in (1 << 20, 1_000_003):
                src = torch.randint(0, 255, (n,), dtype=torch.uint8).pin_memory()
                d = torch.empty(n, dtype=torch.uint8, device=dev)
                back = torch.empty(n, dtype=torch.uint8).pin_memory()
                st = torch.cuda.Stream()
                c2.copy_async(d, src, stream=st.cuda_stream, n_ctas=4)
                c2.copy_async(back, d, stream=st.cuda_stream, n_ctas=4)
                st.synchronize()
                assert torch.equal(back, src)
      finally:
        binding.Context.gap_dtype = keep


def test_external_repeat_set(ctx):
    """blast_kmers as a separately invoked stage: the repeat set comes from a k-mer FASTA, in either orientation,
    unsorted, with duplicates."""
    p = synth.small_params(seed=55, genome_size=500_000)
    seq, off, names = synth.generate(p)
    k, chunk = 23, 20000
    want = util.c_oracle_path(seq, off, names, chunk, k, 10)
    lay = ChunkLayout(names, off, chunk)
    scaffolds, chunk_row = lay.rows()
    ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
    kmers = binding.kmer_decode(want["sel"], k)
    shuffled = [py_oracle.revcomp(km) if i % 2 else km for i, km in enumerate(kmers)][::-1] + kmers[:5]
    n = ctx.set_selected(binding.kmer_encode(shuffled, k), k)
    assert n == len(kmers) and ctx.selected(n).tolist() == want["sel"].tolist()
    nnz = ctx.build_matrix(chunk_row, len(scaffolds))
    indptr, indices, data = ctx.csr(len(scaffolds), nnz)
    assert np.array_equal(indptr, want["indptr"]) and np.array_equal(indices, want["indices"]) and np.array_equal(data, want["data"])


def test_owner_partitioned_counts_two_ranks_on_one_gpu(ctx):
    """The multi-GPU count exchange with both ranks emulated on one device: shard by chromosome,
    count locally, export by owner, merge at the owners, select, gather."""
    import torch
    p = synth.small_params(seed=66, genome_size=800_000, n_chrom=2)
    seq, off, names = synth.generate(p)
    k, chunk, low, world = 23, 20000, 10, 2
    want = util.c_oracle_path(seq, off, names, chunk, k, low)
    shards = [[0, 2], [1, 3]]
    exported = []
    for r in range(world):
        s_seq = np.concatenate([seq[off[c]:off[c + 1]] for c in shards[r]])
        s_off = np.concatenate([[0], np.cumsum([off[c + 1] - off[c] for c in shards[r]])])
        lay = ChunkLayout([names[c] for c in shards[r]], s_off, chunk)
        ctx.load_sequences(s_seq, lay.chunk_begin_abs, lay.chunk_end_abs)
        ctx.count(k)
        d = ctx.count_info()["distinct"]
        keys_t = torch.empty(d, dtype=torch.int64, device="cuda"); counts_t = torch.empty(d, dtype=torch.int32, device="cuda")
        offs = ctx.export_by_owner(world, keys_t, counts_t)
        assert offs[-1] == d
        exported.append((keys_t, counts_t, offs))
    selected = []
    total_distinct = 0
    for owner in range(world):
        ks = torch.cat([exported[r][0][exported[r][2][owner]:exported[r][2][owner + 1]] for r in range(world)])
        cs = torch.cat([exported[r][1][exported[r][2][owner]:exported[r][2][owner + 1]] for r in range(world)])
        ctx.merge_counts(ks, cs, ks.numel(), k, fresh=True)
        total_distinct += ctx.count_info()["distinct"]
        n = ctx.select(low)
        selected.append(ctx.selected(n))
    assert total_distinct == len(want["keys"])
    union = np.sort(np.concatenate(selected))
    assert union.tolist() == want["sel"].tolist()


@pytest.mark.parametrize("world,part_mode", [(2, 3), (4, 3), (3, -1), (2, 2), (4, 2)])
def test_emulated_ranks_end_to_end_on_one_gpu(world, part_mode):
    """The multi-GPU path with every rank emulated by its own context on ONE device, end to end and with the driver's own
    helpers (shard_chunks, local_buffer_plan, owner_ranges, owner_segments): shard the chunks, partition locally, exchange
    the bucket sizes and sub-bucket histograms (summed on the host here instead of all-gather / all-to-all), every owner
    counts its buckets straight out of ALL ranks' buffers (pointer segments, as over NVLink peer memory), thresholds its
    share, the union of the repeat sets is installed everywhere, local matrices, global df, TF-IDF, row blocks
    concatenated -- against the single-process C oracle.  part_mode 3 / -1: super-k-mer records; 2: 8-byte k-mers."""
    import torch
    from polycracker_b200 import distributed as pcd
    from polycracker_b200.pipeline import dump_floor, effective_low_count
    p = synth.small_params(seed=191, genome_size=3_000_000, n_chrom=3, n_shared_families=6, n_specific_families=8)
    seq, off, names = synth.generate(p)
    k, chunk, low = 23, 20000, 10
    want = util.c_oracle_path(seq, off, names, chunk, k, low)
    layout = ChunkLayout(names, off, chunk)
    scaffolds_all, chunk_row_all = layout.rows()
    windows_global = int(np.maximum(0, (layout.chunk_end - layout.chunk_start) - k + 1).sum())
    blocks = pcd.shard_chunks(layout, world)
    ctxs = [binding.Context(0) for _ in range(world)]
    dev = torch.device("cuda", 0)
    try:
        plans, shard = [], []
        for r, c in enumerate(ctxs):
            c.tune(part_mode=part_mode, keep_min=dump_floor(k))
            ids, begins, ends, total = pcd.local_buffer_plan(layout, blocks[r])
            buf = np.empty(total, np.uint8)
            for j, i in enumerate(ids.tolist()):
                a = int(off[layout.chunk_seq[i]])
                buf[begins[j]:ends[j]] = seq[a + layout.chunk_start[i]:a + layout.chunk_end[i]]
            c.load_sequences(buf, begins, ends)
            shard.append((ids, begins, ends))
            plans.append(c.skm_plan(k, windows_global, world))
        skm = int(plans[0][5]) == 1
        assert skm == (part_mode != 2) and all(np.array_equal(pl, plans[0]) for pl in plans)
        plan = plans[0]
        if skm:
            pb, nb2 = int(plan[0]), int(plan[1])
            parts = [c.skm_partition(k, plan) for c in ctxs]                # (off, records ptr, hist ptr)
            hists = [pcd.device_view(torch, h, (1 << pb) * nb2, dev) for _, _, h in parts]
        else:
            smem_mean = (1 << 20) // ctxs[0].sub_plan(1 << 40, 1 << 20)
            pb = pcd.plan_partition_bits(windows_global, world, smem_mean=smem_mean)
            parts = [c.partition(k, pb) for c in ctxs]                      # (off, k-mer ptr)
        P = 1 << pb
        q_lo = pcd.owner_ranges(P, world)
        sizes_all = np.stack([np.diff(pt[0]) for pt in parts])              # [sender, bucket]
        ptrs = np.asarray([pt[1] for pt in parts], dtype=np.uint64)
        n_windows = 0
        selected, distinct = [], 0
        for r, c in enumerate(ctxs):
            seg_begin, seg_len = pcd.owner_segments(sizes_all, q_lo, r)
            seg_base = np.repeat(ptrs[None, :], seg_len.shape[0], axis=0)
            if skm:
                mine = slice(q_lo[r] * nb2, q_lo[r + 1] * nb2)
                staged = torch.stack([h[mine] for h in hists]).sum(dim=0).contiguous()
                torch.cuda.synchronize()
                c.skm_count_from(k, plan, seg_base, seg_begin, seg_len, staged)
            else:
                nb2 = max(c.sub_plan(int(sizes_all[:, q_lo[x]:q_lo[x + 1]].sum()), max(1, q_lo[x + 1] - q_lo[x])) for x in range(world))
                hs = []
                for c2 in ctxs:
                    h = torch.empty(P * nb2, dtype=torch.int64, device=dev)
                    c2.sub_hist(pb, nb2, h)
                    hs.append(h[q_lo[r] * nb2:q_lo[r + 1] * nb2])
                staged = torch.stack(hs).sum(dim=0).contiguous()
                torch.cuda.synchronize()
                c.sub_hist_set(nb2, staged)
                c.count_buckets_from(k, pb, seg_base, seg_begin, seg_len)
            info = c.count_info()
            n_windows += info["valid_windows"]; distinct += info["distinct"]
            n = c.select(effective_low_count(k, low))
            selected.append(c.selected(n))
        assert n_windows == want["n_windows"] and distinct == len(want["keys"])
        sel_all = np.concatenate(selected)
        assert len(np.unique(sel_all)) == len(sel_all)                       # owners select disjoint sets
        rows, dfs = [], []
        for r, c in enumerate(ctxs):
            ids = shard[r][0]
            my = chunk_row_all[ids]
            kept = my[my >= 0]
            row0 = int(kept.min()) if len(kept) else 0
            assert len(kept) == 0 or int(kept.max()) - row0 + 1 == len(kept)
            n_sel = c.set_selected(sel_all, k)
            nnz = c.build_matrix(np.where(my >= 0, my - row0, -1).astype(np.int32), len(kept))
            rows.append((row0, len(kept), nnz))
            dfs.append(c.df(n_sel))
        df = np.sum(dfs, axis=0).astype(np.int32)
        keys = ctxs[0].selected(len(sel_all))
        assert keys.tolist() == want["sel"].tolist()
        indptr, indices, data, tf = [np.zeros(1, np.int64)], [], [], []
        for r, c in sorted(enumerate(ctxs), key=lambda x: rows[x[0]][0] if rows[x[0]][1] else 1 << 60):
            row0, n_rows, nnz = rows[r]
            c.tfidf(len(scaffolds_all), df)
            ip, ix, dt = c.csr(n_rows, nnz)
            indptr.append(ip[1:] + int(indptr[-1][-1])); indices.append(ix); data.append(dt); tf.append(c.tfidf_values(nnz))
        assert np.array_equal(np.concatenate(indptr), want["indptr"])
        assert np.array_equal(np.concatenate(indices), want["indices"]) and np.array_equal(np.concatenate(data), want["data"])
        util.assert_tfidf_close(np.concatenate(tf), want["tfidf"])
    finally:
        for c in ctxs:
            c.close()


def test_cfg1_full_size_parity(ctx):
    """BASELINE.json configs[0] at full size (the reference's own CPU-runnable case: test_pipeline on the two-algae test
    genome with the shipped config -- k = 26, 50 kb chunks, kmer_low_count 30, removeNonChunk = 1, tfidf = 1,
    config_polyCRACKER.txt:27-53; the FASTA itself is not in the reference tree, so a synthetic stand-in of the same
    size and shape): the whole path on the GPU against the C oracle, bit-exact counts / repeat set / CSR, TF-IDF within
    1e-6.  k = 26 is even, so reverse-complement palindromes occur."""
    cfg = synth.CONFIGS["cfg1_algae_standin"]
    p = synth.make_params(cfg["seed"], cfg["genome_size"], n_subgenomes=cfg["n_subgenomes"])
    seq, off, names = synth.generate(p, threads=8)
    res, want = _check_against_c_oracle(ctx, seq, off, names, cfg["chunk_size"], cfg["k"], cfg["kmer_low_count"])
    assert len(seq) > 150_000_000 and res.shape[0] > 3000 and res.shape[1] > 10_000 and len(want["data"]) > 1_000_000
    assert ctx.count_detail()["kind"] == "list"                    # the shared-memory count ran
    got_csc = res.csr_matrix().tocsc()                              # what clusteringMatrix.npz holds
    assert got_csc.dtype == np.float32 and got_csc.shape == (len(want["scaffolds"]), len(want["sel"]))


def test_cfg2_full_size_parity(ctx):
    """BASELINE.json configs[1] -- the bench workload -- at full size: 497 Mbp synthetic allotetraploid, k = 23, 75 kb
    chunks, threshold 30.  The whole GPU path (shared-memory count, filtered probe) against the C oracle: 4.96e8 k-mer
    instances, 3.3e8 distinct, bit-exact repeat set and CSR (9.6e7 entries), TF-IDF within 1e-6."""
    cfg = synth.CONFIGS["cfg2_500Mbp_allotetraploid"]
    p = synth.make_params(cfg["seed"], cfg["genome_size"], n_subgenomes=cfg["n_subgenomes"])
    seq, off, names = synth.generate(p, threads=8)
    res, want = _check_against_c_oracle(ctx, seq, off, names, cfg["chunk_size"], cfg["k"], cfg["kmer_low_count"])
    assert res.stats["valid_windows"] > 490_000_000 and res.shape[1] > 500_000 and len(want["data"]) > 90_000_000
    d = ctx.count_detail()
    # the super-k-mer count ran; a handful of sub-buckets (heavy minimizers inside high-copy repeat families) may take the
    # global-table fall-back, never more than a fraction of a percent
    assert d["kind"] == "list" and d["skm_records"] > 50_000_000 and d["overflowed"] < d["sub_buckets"] // 500
    ctx.release(ctx.REL_BUCKETS | ctx.REL_TABLE | ctx.REL_STREAM | ctx.REL_HITS | ctx.REL_ASCII)   # 20 GB back for the next tests


def _sampled_rows_match(ctx, seq, lay, scaffolds, chunk_row, k, sel, rng, n_sample=600):
    """matrix of the current repeat set on the GPU against the oracle on a random sample of rows (bit-exact), and the
    size-independent invariants on all rows"""
    nnz = ctx.build_matrix(chunk_row, len(scaffolds))
    indptr, indices, data = ctx.csr(len(scaffolds), nnz)
    rows = np.sort(rng.choice(len(scaffolds), size=min(n_sample, len(scaffolds)), replace=False))
    chunk_of_row = np.full(len(scaffolds), -1, np.int64)
    chunk_of_row[chunk_row[chunk_row >= 0]] = np.flatnonzero(chunk_row >= 0)
    w_indptr, w_indices, w_data = c_oracle.matrix(seq, lay.chunk_begin_abs, lay.chunk_end_abs, chunk_of_row[rows], k, sel)
    for j, r in enumerate(rows.tolist()):
        a, b = indptr[r], indptr[r + 1]
        assert np.array_equal(indices[a:b], w_indices[w_indptr[j]:w_indptr[j + 1]]), r
        assert np.array_equal(data[a:b], w_data[w_indptr[j]:w_indptr[j + 1]]), r
    assert np.array_equal(ctx.df(len(sel)), np.bincount(indices, minlength=len(sel)))
    return nnz


def test_scale_1gbp_threshold_100_parity(ctx):
    """BASELINE configs[2]/[3] regime on one GPU (the oracle would need minutes at 4.5 Gbp): a 1 Gbp genome of the cfg3
    shape, k = 23, threshold 100.  Repeat set bit-exact against the C oracle; CSR bit-exact on 600 random rows for every
    probe variant that only appears at scale (L2 bitmap, compact 8-byte table, fingerprint walk, forced level-1 width
    pb = 10); then a threshold sweep (100 -> 300 -> 1000) on the kept count list, as rerun_threshold does per rank."""
    cfg = synth.CONFIGS["cfg3_4.5Gbp_tobacco"]
    p = synth.make_params(cfg["seed"], 1_000_000_000, n_subgenomes=cfg["n_subgenomes"])
    seq, off, names = synth.generate(p, threads=8)
    k, chunk = cfg["k"], cfg["chunk_size"]
    lay = ChunkLayout(names, off, chunk)
    scaffolds, chunk_row = lay.rows()
    rng = np.random.default_rng(1)
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    try:
        ctx.tune(part_mode=-1, smem_pb=10, keep_min=3)
        ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
        ctx.count(k)
        info, d = ctx.count_info(), ctx.count_detail()
        assert info["valid_windows"] == n_windows > 990_000_000 and info["distinct"] == len(want_keys)
        assert d["skm_records"] > 100_000_000 and info["partition_bits"] == 10 and not ctx.count_flags()["could_wrap"]
        for low, variants in ((100, (dict(), dict(probe_table8=1), dict(probe_filter=0, probe_table8=0, probe_fp=1), dict(probe_filter=2, probe_table8=0))),
                              (300, (dict(),)), (1000, (dict(),))):
            sel = c_oracle.select(want_keys, want_counts, k, low)
            for kw in variants:
                ctx.tune(probe_filter=-1, probe_table8=-1, probe_fp=-1)
                ctx.tune(**kw)
                n_sel = ctx.select(low)                            # (re)builds the look-up structures for this variant
                assert n_sel == len(sel) and np.array_equal(ctx.selected(n_sel), sel)
                nnz = _sampled_rows_match(ctx, seq, lay, scaffolds, chunk_row, k, sel, rng)
            assert len(sel) > 1000 and nnz > 1_000_000
    finally:
        ctx.tune(part_mode=-1, smem_pb=0, keep_min=1, probe_filter=-1, probe_table8=-1, probe_fp=-1)
        ctx.release(ctx.REL_BUCKETS | ctx.REL_TABLE | ctx.REL_STREAM | ctx.REL_HITS | ctx.REL_ASCII)


@pytest.mark.parametrize("k", [15, 31])
def test_k_sweep_300mbp_parity(ctx, k):
    """BASELINE configs[4] (k = 15 / 31 next to the k = 23 of the other tests) at 300 Mbp: k = 15 takes the k-mer partition
    (4^15 / 2 distinct k-mers at most), k = 31 the super-k-mer one with the longest minimizer window; counts of everything
    the reference's counter would dump (>= 3), repeat set and 600 sampled CSR rows bit-exact, TF-IDF rows unit length."""
    p = synth.make_params(20260005, 300_000_000)
    seq, off, names = synth.generate(p, threads=8)
    chunk, low = 75000, 100
    lay = ChunkLayout(names, off, chunk)
    scaffolds, chunk_row = lay.rows()
    want_keys, want_counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    try:
        ctx.tune(keep_min=3)
        ctx.load_sequences(seq, lay.chunk_begin_abs, lay.chunk_end_abs)
        ctx.count(k)
        info, d = ctx.count_info(), ctx.count_detail()
        assert info["valid_windows"] == n_windows and info["distinct"] == len(want_keys)
        assert (d["skm_m"] > 0) == (k >= 17)
        keys, counts = ctx.dump_counts(3)
        order = np.argsort(keys)
        m = want_counts >= 3
        assert np.array_equal(keys[order], want_keys[m]) and np.array_equal(counts[order].astype(np.uint64), want_counts[m])
        sel = c_oracle.select(want_keys, want_counts, k, low)
        n_sel = ctx.select(low)
        assert np.array_equal(ctx.selected(n_sel), sel) and n_sel > 100
        nnz = _sampled_rows_match(ctx, seq, lay, scaffolds, chunk_row, k, sel, np.random.default_rng(k))
        ctx.tfidf(len(scaffolds))
        indptr, _, _ = ctx.csr(len(scaffolds), nnz)
        tf = ctx.tfidf_values(nnz).astype(np.float64)
        norms = np.sqrt(np.add.reduceat(tf * tf, indptr[:-1][np.diff(indptr) > 0]))
        np.testing.assert_allclose(norms, 1.0, rtol=1e-6)
    finally:
        ctx.tune(keep_min=1)
        ctx.release(ctx.REL_BUCKETS | ctx.REL_TABLE | ctx.REL_STREAM | ctx.REL_HITS | ctx.REL_ASCII)


def test_properties_at_scale(ctx):
    """Size-independent invariants on a 60 Mbp genome (the oracle would take minutes here): column sums over all
    chunks equal the global counts, row sums bounded by windows, unit L2 rows, df bounds."""
    p = synth.make_params(20260002, 60_000_000)
    seq, off, names = synth.generate(p, threads=8)
    k, chunk, low = 23, 75000, 30
    res = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=k, kmer_low_count=low, remove_non_chunk=0)
    m = res.csr_matrix()
    assert res.shape[1] > 1000 and m.nnz > 10000
    keys, counts = ctx.dump_counts(low)
    order = np.argsort(keys)
    assert np.array_equal(keys[order], res.kmer_keys)                      # repeat set == thresholded dump
    assert np.array_equal(np.asarray(m.sum(axis=0)).ravel().astype(np.int64), counts[order].astype(np.int64))
    lens = res.layout.chunk_end - res.layout.chunk_start
    row_len = dict(zip(res.layout.chunk_names, lens.tolist()))
    rs = np.asarray(m.sum(axis=1)).ravel()
    assert all(rs[i] <= max(0, row_len[s] - k + 1) for i, s in enumerate(res.scaffolds))
    assert np.all(np.diff(res.kmer_keys.astype(np.uint64)) > 0)            # sorted, unique
    for r in range(0, m.shape[0], 97):                                     # sorted column ids inside rows
        seg = res.indices[res.indptr[r]:res.indptr[r + 1]]
        assert np.all(np.diff(seg) > 0)
    assert np.array_equal(res.df, np.bincount(res.indices, minlength=res.shape[1]))
    t = res.csr_matrix("tfidf")
    norms = np.sqrt(np.asarray(t.multiply(t).sum(axis=1)).ravel())
    nz = np.diff(res.indptr) > 0
    np.testing.assert_allclose(norms[nz], 1.0, rtol=1e-6)
    assert np.all(norms[~nz] == 0)
    # and the C oracle on a slice of the same genome (first 2 chromosomes), bit-exact
    cut = int(off[2])
    _check_against_c_oracle(ctx, seq[:cut].copy(), off[:3], names[:2], chunk, k, 5)
