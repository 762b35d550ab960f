"""Scope row N1 (subgenome extraction).  Golden fixtures tests/golden/sub_*: produced by running the REFERENCE'S OWN
subgenomeExtraction / compareKmers / blast2bed3 / bed2unionBed / kmerratio2scaffasta code
(tests/golden/make_golden_sub.py).  CPU tests pin the oracle (literal and numpy flavours) and the host-side binning to
them; GPU tests drive the file-based mirror and the in-memory pass against the fixtures and the oracle."""
import os
import shutil

import numpy as np
import pytest

from oracle import py_oracle, sub_oracle
from polycracker_b200 import binding, subgenome, synth
from polycracker_b200.pipeline import ChunkLayout

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = sorted(d for d in os.listdir(GOLD) if d.startswith("sub_"))


def _cfg(case):
    cfg = {}
    for line in open(os.path.join(GOLD, case, "params.txt")):
        k, v = line.strip().split(" = ")
        cfg[k] = v if k in ("kmer_length", "unionbed_threshold") else (float(v) if "." in v else int(v))
    return cfg


def _fasta(path):
    seqs = []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            seqs.append([line[1:].split()[0], []])
        else:
            seqs[-1][1].append(line)
    return [(n, "".join(p)) for n, p in seqs]


def _chunk_records(case, cfg):
    seqs = _fasta(os.path.join(GOLD, case, "fasta_files", "genome.fa"))
    rows = py_oracle.split_chunks([(n, len(s)) for n, s in seqs], cfg["chunk_size"])
    return py_oracle.chunk_records(seqs, rows)


def _groups(folder):
    out = []
    for fn in sorted(os.listdir(folder)):
        if fn.startswith("subgenome_") and fn.endswith(".txt") and os.path.getsize(os.path.join(folder, fn)):
            out.append([l.strip() for l in open(os.path.join(folder, fn)) if l.strip()])
    return out


def _iterations(case):
    base = os.path.join(GOLD, case, "analysisOutputs")
    return sorted(d for d in os.listdir(base) if d.startswith("bootstrap_") and
                  os.path.exists(os.path.join(base, d, "subgenomes.union.bedgraph")))


def _union(path):
    return [(l.split("\t")[0], [int(v) for v in l.rstrip("\n").split("\t")[3:]]) for l in open(path) if l.strip()]


# ------------------------------------------------------------------------------------------------ CPU: oracle vs golden
@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_outputs(case):
    cfg = _cfg(case)
    records = _chunk_records(case, cfg)
    by_name = dict(records)
    ks = [int(x) for x in cfg["kmer_length"].split(",")]
    a_thr, r_thr = [int(x) for x in cfg["unionbed_threshold"].split(",")]
    base = os.path.join(GOLD, case, "analysisOutputs")
    for it in _iterations(case):
        folder = os.path.join(base, it)
        groups = _groups(folder)
        dicts = sub_oracle.subgenome_dictionaries([[(n, by_name[n]) for n in g] for g in groups], ks)
        diff = sub_oracle.compare_kmers(dicts, ks, cfg["diff_kmer_threshold"], cfg["default_kmercount_value"],
                                        cfg["diff_sample_rate"])
        cov = []
        for g in range(len(groups)):
            want = open(os.path.join(folder, "kmercount_files", "model_subgenome_%d.higher.kmers.fa" % g)).read()
            assert sub_oracle.higher_kmers_fasta(diff[g], cfg["diff_sample_rate"] > 1) == want
            cov.append(sub_oracle.map_back_literal([d[0] for d in diff[g]], records))
            bedgraph = open(os.path.join(folder, "model_subgenome_%d.higher.kmers.sorted.cov.bedgraph" % g)).read()
            assert bedgraph == "".join("%s\t0\t%d\t%d\n" % (n, len(s), cov[g][n]) for n, s in records)
        union = _union(os.path.join(folder, "subgenomes.union.bedgraph"))
        assert [u[0] for u in union] == sorted((n for n, _ in records), key=lambda s: s.encode())
        assert [u[1] for u in union] == [[c[n] for c in cov] for n, _ in union]
        binned, ambiguous = sub_oracle.bin_chunks([u[0] for u in union], [u[1] for u in union], a_thr, r_thr)
        ex = os.path.join(folder, "extractedSubgenomes")
        prefix = "genome_split.subgenome"
        for g in range(len(groups)):
            assert binned[g] == [l.strip() for l in open(os.path.join(ex, prefix + chr(65 + g) + ".names"))]
        assert ambiguous == [l.strip() for l in open(os.path.join(ex, "ambiguousScaffolds.names"))]
        # product host logic (no GPU): the vectorised binning rule
        member = subgenome.bin_membership(np.array([u[1] for u in union]), a_thr, r_thr)
        for g in range(len(groups)):
            assert [u[0] for u, m in zip(union, member[:, g]) if m] == binned[g]


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_matches_literal(case):
    cfg = _cfg(case)
    records = _chunk_records(case, cfg)
    ks = [int(x) for x in cfg["kmer_length"].split(",")]
    groups = _groups(os.path.join(GOLD, case, "analysisOutputs", "bootstrap_0"))
    names = [n for n, _ in records]
    seq = np.frombuffer("".join(s for _, s in records).encode(), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for _, s in records])]).astype(np.int64)
    by_name = dict(records)
    dicts = sub_oracle.subgenome_dictionaries([[(n, by_name[n]) for n in g] for g in groups], ks)
    diff = sub_oracle.compare_kmers(dicts, ks, cfg["diff_kmer_threshold"], cfg["default_kmercount_value"], 1)
    diff_by_k = {}
    for k in ks:
        arrs = []
        for g in groups:
            idx = [names.index(n) for n in g]
            arrs.append(sub_oracle.count_group(seq, off[idx], off[np.array(idx) + 1], k, 3 if k <= 31 else 1))
        got = sub_oracle.compare_arrays(arrs, cfg["diff_kmer_threshold"], cfg["default_kmercount_value"], 1)
        for g in range(len(groups)):
            want = [d for d in diff[g] if len(d[0]) == k]
            assert binding.kmer_decode(got[g][0], k) == [d[0] for d in want]
            assert got[g][1][:, g].tolist() == [d[1] for d in want]
        diff_by_k[k] = [g_[0] for g_ in got]
    cov = sub_oracle.map_back(seq, off[:-1], off[1:], diff_by_k)
    for g in range(len(groups)):
        lit = sub_oracle.map_back_literal([d[0] for d in diff[g]], records)
        assert cov[:, g].tolist() == [lit[n] for n in names]


def test_py2_division_semantics():
    """The two quotient kinds of compareKmers: floor when the other dictionary holds the k-mer, float otherwise."""
    dicts = [{"ACGTA": 41, "CCCCA": 62, "GGGTA": 7}, {"ACGTA": 2, "GGGTA": 3}]
    got = sub_oracle.compare_kmers(dicts, [5], 20, 3., 1)
    # 41 // 2 = 20 is NOT > 20 (true division 20.5 would be); 62 / 3.0 = 20.67 > 20; 7 // 3 = 2
    assert [d[0] for d in got[0]] == ["CCCCA"] and got[0][0][2] == [3.0]
    assert got[1] == []
    assert sub_oracle.higher_kmers_fasta(got[0]) == ">CCCCA.62.3\nCCCCA\n"
    assert sub_oracle.higher_kmers_fasta(got[0], sampled=True) == ">CCCCA.62.3.0\nCCCCA\n"


# ------------------------------------------------------------------------------------------------ GPU
def _copy_inputs(case, tmp_path):
    cfg = _cfg(case)
    work = tmp_path / "work"
    (work / "fasta_files").mkdir(parents=True)
    (work / "analysisOutputs" / "bootstrap_0").mkdir(parents=True)
    records = _chunk_records(case, cfg)
    with open(work / "fasta_files" / "genome_split.fa", "w") as f:
        for n, s in records:
            f.write(">%s\n%s\n" % (n, s))
    src = os.path.join(GOLD, case, "analysisOutputs", "bootstrap_0")
    for fn in os.listdir(src):
        if fn.startswith("subgenome_") and fn.endswith(".txt"):
            shutil.copy(os.path.join(src, fn), work / "analysisOutputs" / "bootstrap_0" / fn)
    return cfg, work


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_subgenome_extraction_files_match_reference(case, tmp_path):
    cfg, work = _copy_inputs(case, tmp_path)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        subgenome.subgenomeExtraction("analysisOutputs/bootstrap_0", "analysisOutputs", "./fasta_files/",
                                      "genome_split.fa", "genome_split.fa", bootstrap=cfg["bootstrap"],
                                      kmer_length=cfg["kmer_length"], diff_kmer_threshold=cfg["diff_kmer_threshold"],
                                      unionbed_threshold=cfg["unionbed_threshold"],
                                      diff_sample_rate=cfg["diff_sample_rate"],
                                      default_kmercount_value=cfg["default_kmercount_value"])
    finally:
        os.chdir(cwd)
    gold = os.path.join(GOLD, case, "analysisOutputs")
    checked = 0
    for root, _dirs, files in os.walk(gold):
        for fn in files:
            rel = os.path.relpath(os.path.join(root, fn), gold)
            if fn.endswith(".names"):
                got = [l[1:].strip() for l in open(work / "analysisOutputs" / (rel[:-len(".names")] + ".fasta"))
                       if l.startswith(">")]
                assert got == [l.strip() for l in open(os.path.join(root, fn))], rel
            else:
                assert open(work / "analysisOutputs" / rel).read() == open(os.path.join(root, fn)).read(), rel
            checked += 1
    assert checked >= 10
    # nothing the reference does not write either (e.g. a finalResults file for an empty subgenome)
    for d in ("finalResults",):
        assert sorted(os.listdir(work / "analysisOutputs" / d)) == sorted(os.listdir(os.path.join(gold, d)))


@pytest.mark.gpu
@pytest.mark.parametrize("sampling,n_sub", [(1, 2), (3, 3)])
def test_gpu_extract_pass_matches_numpy_oracle(sampling, n_sub):
    p = synth.small_params(seed=9090 + n_sub, genome_size=2_400_000, n_subgenomes=n_sub, n_chrom=2,
                           n_shared_families=5, n_specific_families=4 * n_sub)
    seq, off, names = synth.generate(p)
    lay = ChunkLayout(names, off, 20000)
    begins, ends = lay.chunk_begin_abs, lay.chunk_end_abs
    rng = np.random.RandomState(5)
    truth = np.array(["ABDEFGH".index(n[3]) for n in lay.chunk_names])
    label = np.where(rng.rand(len(truth)) < 0.15, -1, truth)              # some chunks unassigned,
    flip = rng.rand(len(truth)) < 0.05                                    # a few in the wrong cluster
    label[flip & (label >= 0)] = (label[flip & (label >= 0)] + 1) % n_sub
    groups = [np.flatnonzero(label == g) for g in range(n_sub)]
    ks = [23, 31, 32]
    thr, dflt = 4, 3.
    with binding.Context(0) as ctx:
        cov, diff, n_diff = subgenome.extract_pass(ctx, seq, begins, ends, groups, ks, thr, dflt, sampling)
        timings = ctx.timings()
    assert timings["sub_dict"] > 0 and timings["sub_compare"] > 0 and timings["sub_mapback"] > 0
    diff_by_k = {}
    for k in ks:
        arrs = [sub_oracle.count_group(seq, begins[g], ends[g], k, 3 if k <= 31 else 1) for g in groups]
        want = sub_oracle.compare_arrays(arrs, thr, dflt, sampling)
        for g in range(n_sub):
            assert np.array_equal(diff[k][g][0], want[g][0]), (k, g)
            assert np.array_equal(diff[k][g][1].astype(np.int64), want[g][1]), (k, g)
            assert n_diff[k][g] == len(want[g][0])
        diff_by_k[k] = [w[0] for w in want]
        assert sum(len(w[0]) for w in want) > 50
    want_cov = sub_oracle.map_back(seq, begins, ends, diff_by_k)
    assert np.array_equal(cov.astype(np.int64), want_cov)
    assert want_cov.sum() > 1000


@pytest.mark.gpu
def test_gpu_sub_session_errors():
    seq = np.frombuffer(b"ACGTACGTTTGACCAGTAGGATCCATTTACGACGGGATATTTACGCGCGATATATTAGAGAGCTCGA" * 40, np.uint8)
    with binding.Context(0) as ctx:
        with pytest.raises(binding.PolycrackerError):
            ctx.sub_begin(9)
        ctx.sub_begin(2)
        ctx.load_sequences(seq, [0], [len(seq)])
        ctx.count(11)
        assert ctx.sub_add_counts(0, 3) > 0
        with pytest.raises(binding.PolycrackerError):                      # subgenome 1 has no dictionary yet
            ctx.sub_compare(2, 3.)
        ctx.count(13)
        with pytest.raises(binding.PolycrackerError):                      # k changed before compare
            ctx.sub_add_counts(1, 3)
        ctx.count(11)
        ctx.sub_add_counts(1, 3)
        with pytest.raises(binding.PolycrackerError):
            ctx.sub_compare(2, 0.)                                         # division by zero, as in the reference
        n = ctx.sub_compare(2, 3.)
        assert n.tolist() == [0, 0]                                        # identical subgenomes: nothing differential
        cov = ctx.sub_mapback(1)
        assert cov.tolist() == [[0, 0]]


def test_n4_literal_oracle_known_answer():
    """Scope row N4 (bio_hyp_class, utilities.py:255-258) on a hand-checkable case: floor division, implicit zeros."""
    recs_a = [("a", "ACGTTGCA" * 30)]
    recs_b = [("b", "ACGTTGCA" * 2 + "TTTTCCCC" * 30)]
    n_union, kept, rows = sub_oracle.kmer_genome_matrix_literal([recs_a, recs_b], 5, min_count=3, ratio_threshold=5,
                                                                 default_kcount_value=1, max_val=20)
    assert n_union >= len(kept) > 0
    for km, row in zip(kept, rows):
        mx, mn = max(row), min(row)
        assert mx >= 20 and mx // max(mn, 1) >= 5
    # a k-mer present 29 times in one genome and 6 times in the other: 29 // 6 = 4 < 5 -> dropped although 29 / 6 = 4.83
    _, kept2, rows2 = sub_oracle.kmer_genome_matrix_literal([[("x", "ACGTA" * 29)], [("y", "ACGTA" * 6)]], 5, 3, 5, 1, 20)
    assert all(r[0] // max(r[1], 1) >= 5 for r in rows2)
    assert "ACGTA" not in kept2


@pytest.mark.gpu
@pytest.mark.parametrize("n_gen,k,min_count,ratio,dflt,max_val", [(2, 23, 3, 20, 1, 100), (3, 15, 3, 4, 1, 30), (4, 11, 5, 3, 2, 50),
                                                                   (2, 32, 1, 2, 1, 3)])
def test_gpu_n4_kmer_genome_matrix(n_gen, k, min_count, ratio, dflt, max_val):
    """Scope row N4: the k-mer x genome count matrix of bio_hyp_class on the GPU against the restated oracle (literal on a
    small case, numpy on a Mbp-sized one)."""
    # small: literal oracle (string k-mers)
    p = synth.small_params(seed=7000 + n_gen, genome_size=40_000 * n_gen, n_subgenomes=n_gen, n_chrom=1, n_shared_families=3,
                           n_specific_families=3 * n_gen, max_family_len=600)
    seq, off, names = synth.generate(p)
    from tests import util
    seqs = util.to_sequences(seq, off, names)
    groups = [np.array([i]) for i in range(len(names))][:n_gen]
    begins, ends = off[:-1].copy(), off[1:].copy()
    small = (max(2, ratio // 4), max(3, max_val // 10))
    with binding.Context(0) as ctx:
        nu, keys, counts = subgenome.kmer_genome_matrix(ctx, seq, begins, ends, groups, k, min_count, small[0], dflt, small[1])
    want_nu, want_kept, want_rows = sub_oracle.kmer_genome_matrix_literal([[seqs[i]] for i in range(n_gen)], k, min_count,
                                                                           small[0], dflt, small[1])
    assert nu == want_nu and binding.kmer_decode(keys, k) == want_kept
    assert counts.astype(np.int64).tolist() == want_rows
    # Mbp-sized: numpy oracle from the C counter
    p = synth.small_params(seed=7100 + n_gen, genome_size=600_000 * n_gen, n_subgenomes=n_gen, n_chrom=1, n_shared_families=5,
                           n_specific_families=4 * n_gen)
    seq, off, names = synth.generate(p)
    begins, ends = off[:-1].copy(), off[1:].copy()
    groups = [np.array([i]) for i in range(n_gen)]
    with binding.Context(0) as ctx:
        nu, keys, counts = subgenome.kmer_genome_matrix(ctx, seq, begins, ends, groups, k, min_count, ratio, dflt, max_val)
    arrs = [sub_oracle.count_group(seq, begins[g], ends[g], k, min_count) for g in groups]
    union = np.unique(np.concatenate([a[0] for a in arrs]))
    mat = np.zeros((len(union), n_gen), np.int64)
    for g, (kk, cc) in enumerate(arrs):
        mat[np.searchsorted(union, kk), g] = cc
    mx, mn = mat.max(axis=1), mat.min(axis=1)
    keep = ((mx // np.maximum(mn, dflt)) >= ratio) & (mx >= max_val)
    assert nu == len(union)
    assert np.array_equal(keys, union[keep]) and np.array_equal(counts.astype(np.int64), mat[keep])
    assert keep.sum() > 0
