"""Golden fixtures (tests/golden/case_*): produced by running the REFERENCE'S OWN Python stage code
(tests/golden/make_golden.py).  CPU tests pin the oracle and the host-side stage mirror to them; the GPU test drives
the whole file-based pipeline (the six sub-commands, in order, through files) and compares every artefact."""
import os
import pickle
import re
import shutil

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import py_oracle
from polycracker_b200 import binding, stages
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = sorted(d for d in os.listdir(GOLD) if d.startswith("case_"))
REF = "/root/reference/polycracker/polycracker.py"


def _cfg(case):
    cfg = {}
    for line in open(os.path.join(GOLD, case, "params.txt")):
        k, v = line.strip().split(" = ")
        cfg[k] = int(v)
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


def _merged(path):
    out = {}
    for line in open(path):
        c, s, e, names = line.rstrip("\n").split("\t")
        out[c] = (s, e, sorted(names.split(",")))
    return out


def _assert_same_npz(a, b, exact=True):
    ma, mb = sps.load_npz(a), sps.load_npz(b)
    assert ma.format == mb.format and ma.shape == mb.shape and ma.dtype == mb.dtype
    assert np.array_equal(ma.indptr, mb.indptr) and np.array_equal(ma.indices, mb.indices)
    if exact:
        assert np.array_equal(ma.data, mb.data)
    else:
        util.assert_tfidf_close(ma.data, mb.data)


def _write_sam(path, hits):
    with open(path, "w") as f:
        for q, flag, rname, pos in hits:
            f.write("\t".join([q, str(flag), rname, str(pos), "255", "%dM" % len(q), "*", "0", "0", q, "*"]) + "\n")


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_artefacts(case):
    cfg, g = _cfg(case), os.path.join(GOLD, case)
    seqs = _fasta(os.path.join(g, "fasta_files", "genome.fa"))
    r = py_oracle.run_path(seqs, cfg["chunk_size"], cfg["k"], cfg["kmer_low_count"], cfg["high_bool"], cfg["kmer_high_count"],
                           cfg["sampling"], cfg["remove_non_chunk"])
    bed = [tuple(l.rstrip("\n").split("\t")) for l in open(os.path.join(g, "correspondence.bed"))]
    assert [(c, str(s), str(e), n) for c, s, e, n in r["bed"]] == bed
    assert r["kmers"] == pickle.load(open(os.path.join(g, "kmers.p"), "rb"))
    assert r["scaffolds"] == pickle.load(open(os.path.join(g, "scaffolds.p"), "rb"))
    assert r["original_scaffolds"] == pickle.load(open(os.path.join(g, "originalScaffolds.p"), "rb"))
    want = sps.load_npz(os.path.join(g, "clusteringMatrix.npz"))
    got = r["matrix"]
    assert got.format == "csc" and got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    tf = sps.load_npz(os.path.join(g, "clusteringMatrix.tfidf.npz"))
    mine = sps.csr_matrix(r["tfidf"]); mine.sort_indices()
    assert np.array_equal(mine.indices, tf.indices)
    util.assert_tfidf_close(mine.data, tf.data)
    # a5: grouping of the hits by chunk
    recs = py_oracle.chunk_records(seqs, r["bed"])
    merged = {c: (str(s), str(e), sorted(n.split(","))) for c, s, e, n in py_oracle.blast2bed(py_oracle.blast_kmers(r["kmers"], recs))}
    assert merged == _merged(os.path.join(g, "blasted_merged.bed"))


@pytest.mark.parametrize("case", CASES)
def test_host_stage_mirror_against_reference_artefacts(case, tmp_path, monkeypatch):
    """splitFasta, kmer2Fasta and the text route of blast2bed are host code: byte-identical files on CPU."""
    cfg, g = _cfg(case), os.path.join(GOLD, case)
    monkeypatch.chdir(tmp_path)
    os.makedirs("fasta_files"); os.makedirs("kmercount_files"); os.makedirs("blast_files")
    shutil.copy(os.path.join(g, "fasta_files", "genome.fa"), "fasta_files/genome.fa")
    stages.splitFasta("genome.fa", "./fasta_files/", cfg["chunk_size"], 0, 0, cfg["remove_non_chunk"], 0, 0)
    for name in ("correspondence.bed", "correspondenceOriginal.bed"):
        assert open(name).read() == open(os.path.join(g, name)).read(), name
    assert open("fasta_files/genome_split.fa").read() == open(os.path.join(g, "fasta_files", "genome_split.fa")).read()
    with pytest.raises(OSError):
        stages.splitFasta("genome.fa", "./nope/", 1000)
    shutil.copy(os.path.join(g, "kmercount_files", "genome_split.kcount"), "kmercount_files/genome_split.kcount")
    stages.kmer2Fasta("./kmercount_files/", cfg["kmer_low_count"], cfg["high_bool"], cfg["kmer_high_count"], cfg["sampling"])
    assert open("kmercount_files/genome_split.kcount.fa").read() == \
        open(os.path.join(g, "kmercount_files", "genome_split.kcount.fa")).read()
    recs = _fasta("fasta_files/genome_split.fa")
    kmers = [l.strip() for l in open("kmercount_files/genome_split.kcount.fa").readlines()[1::2]]
    _write_sam("blast_files/results.sam", py_oracle.blast_kmers(kmers, recs))
    stages.blast2bed("./blast_files/results.sam", 1, 0, "blasted.bed", "blasted_merged.bed", False)
    assert open("blasted_merged.bed").read() == open(os.path.join(g, "blasted_merged.bed")).read()
    with pytest.raises(binding.PolycrackerError, match="perfect_mode"):
        stages.blast_kmers(reference_genome="fasta_files/genome_split.fa", kmer_fasta="kmercount_files/genome_split.kcount.fa",
                           perfect_mode=0)
    with pytest.raises(binding.PolycrackerError, match="unsupported"):
        stages.writeKmerCount("./fasta_files/", "./kmercount_files", "33", "1")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted (GPU box)")
def test_cli_surface_matches_reference_click_options():
    """Same sub-command names, option names, short flags and defaults as polycracker.py:113-316."""
    import click
    from polycracker_b200 import cli
    text = open(REF).read()
    fmt_text = open(os.path.join(os.path.dirname(REF), "format.py")).read()         # hipmer_output_to_kcount lives in format.py
    for cmd_name, cmd in cli.polycracker.commands.items():
        m = re.search(r"@polycracker\.command\(name=['\"]%s['\"]\)\n((?:@click\.option\(.*\)\n)+)(?:@click\.pass_context\n)?def %s\(" % (cmd_name, cmd_name), text)
        if m is None:
            m = re.search(r"@format\.command\(\)\n((?:@click\.option\(.*\)\n)+)def %s\(" % cmd_name, fmt_text)
        assert m, cmd_name
        ref_opts = {}
        for line in m.group(1).strip().split("\n"):
            head = re.match(r"@click\.option\(((?:'-{1,2}[A-Za-z_0-9]+',\s*)+)", line).group(1)
            flags = re.findall(r"'(-{1,2}[A-Za-z_0-9]+)'", head)
            dm = re.search(r"[ ,]default\s*=\s*('[^']*'|[^,)]+)", line)
            ref_opts[tuple(flags)] = dm.group(1).strip().strip("'") if dm else None
        mine = {}
        for p in cmd.params:
            if isinstance(p, click.Option):
                mine[tuple(sorted(p.opts, key=len))] = p.default
        assert {tuple(sorted(k, key=len)) for k in ref_opts} == set(mine), cmd_name
        for flags, d in ref_opts.items():
            got = mine[tuple(sorted(flags, key=len))]
            if d is not None:
                assert str(got) == d or (isinstance(got, float) and float(d) == got), (cmd_name, flags, got, d)
    assert set(cli.polycracker.commands) == {"splitFasta", "writeKmerCount", "kmer2Fasta", "blast_kmers", "blast2bed",
                                             "generate_Kmer_Matrix", "subgenomeExtraction",
                                             "number_repeatmers_per_subsequence", "kcount_hist", "hipmer_output_to_kcount"}


@pytest.mark.gpu
@pytest.mark.parametrize("sidecar", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_file_based_pipeline_on_gpu_matches_reference_artefacts(case, sidecar, tmp_path, monkeypatch):
    """The six stages as separate calls through files, like polycracker.nf runs them."""
    cfg, g = _cfg(case), os.path.join(GOLD, case)
    monkeypatch.chdir(tmp_path)
    os.makedirs("fasta_files"); os.makedirs("blast_files")
    shutil.copy(os.path.join(g, "fasta_files", "genome.fa"), "fasta_files/genome.fa")
    stages.splitFasta("genome.fa", "./fasta_files/", cfg["chunk_size"], 0, 0, cfg["remove_non_chunk"], 0, 0)
    stages.writeKmerCount("./fasta_files/", "./kmercount_files", str(cfg["k"]), "5")
    assert open("kmercount_files/genome_split.kcount").read() == open(os.path.join(g, "kmercount_files", "genome_split.kcount")).read()
    stages.kmer2Fasta("./kmercount_files/", cfg["kmer_low_count"], cfg["high_bool"], cfg["kmer_high_count"], cfg["sampling"])
    assert open("kmercount_files/genome_split.kcount.fa").read() == \
        open(os.path.join(g, "kmercount_files", "genome_split.kcount.fa")).read()
    stages.blast_kmers("5", "./fasta_files/genome_split.fa", "./kmercount_files/genome_split.kcount.fa", "results.sam", 13, 1, 4)
    os.rename("results.sam", "blast_files/results.sam")                    # what the Nextflow process does
    if sidecar:
        os.rename("results.sam.pcb200.npz", "blast_files/results.sam.pcb200.npz")
    else:
        os.remove("results.sam.pcb200.npz")
    stages.blast2bed("./blast_files/results.sam", 1, 0, "blasted.bed", "blasted_merged.bed", False)
    assert _merged("blasted_merged.bed") == _merged(os.path.join(g, "blasted_merged.bed"))
    stages.generate_Kmer_Matrix("./kmercount_files/", "genome_split.fa", cfg["chunk_size"], 0, cfg["remove_non_chunk"], 0, 0, 0,
                                "./fasta_files/", "genome_split.fa", "genome.fa", "blasted.bed", "blasted_merged.bed",
                                "clusteringMatrix.npz", "kmers.p", cfg["tfidf"])
    _assert_same_npz("clusteringMatrix.npz", os.path.join(g, "clusteringMatrix.npz"))
    _assert_same_npz("clusteringMatrix.tfidf.npz", os.path.join(g, "clusteringMatrix.tfidf.npz"), exact=False)
    for name in ("kmers.p", "scaffolds.p", "originalScaffolds.p"):
        assert pickle.load(open(name, "rb")) == pickle.load(open(os.path.join(g, name), "rb")), name
        assert open(name, "rb").read()[:2] == b"\x80\x02"                  # pickle protocol 2
    for name in ("rowNames.txt", "colNames.txt"):
        assert open(name).read() == open(os.path.join(g, name)).read(), name
    # the reference's downstream call on OUR artefact gives the reference's TF-IDF (polycracker.py:384-396)
    from sklearn.feature_extraction.text import TfidfTransformer
    tf = sps.csr_matrix(TfidfTransformer().fit_transform(sps.load_npz("clusteringMatrix.npz"))); tf.sort_indices()
    util.assert_tfidf_close(sps.load_npz("clusteringMatrix.tfidf.npz").data, tf.data)


@pytest.mark.parametrize("case", ["case_k23", "case_k11_high"])
def test_n2_diagnostics_on_reference_artefacts(case):
    """Scope row N2 on the artefacts the reference's own stage code produced: k-mer names per merged-bed line
    (number_repeatmers_per_subsequence, polycracker.py:1374) and the count histogram of the .kcount file (kcount_hist,
    polycracker.py:1466-1469): host stage mirror == literal oracle == row sums of the reference's matrix."""
    import scipy.sparse as sps
    from oracle import py_oracle
    from polycracker_b200 import stages
    d = os.path.join(GOLD, case)
    merged = os.path.join(d, "blasted_merged.bed")
    got = stages.repeatmers_per_subsequence(merged)
    lines = open(merged).read().splitlines()
    assert got.tolist() == py_oracle.repeatmers_per_subsequence(lines)
    m = sps.load_npz(os.path.join(d, "clusteringMatrix.npz")).tocsr()
    rows = open(os.path.join(d, "rowNames.txt")).read().split("\n")
    row_of = {r.split("\t")[1]: int(r.split("\t")[0]) for r in rows if r}
    sums = np.asarray(m.sum(axis=1)).ravel()
    for line, n in zip(lines, got.tolist()):
        chunk = line.split("\t")[0]
        if chunk in row_of:                                  # chunks filtered out of the matrix still have a bed line
            assert sums[row_of[chunk]] == n
    kdir = os.path.join(d, "kmercount_files")
    kfiles = [f for f in os.listdir(kdir) if f.endswith(".kcount")]
    assert kfiles
    for f in kfiles:
        values, freq = stages.kcount_histogram(os.path.join(kdir, f))
        assert list(zip(values.tolist(), freq.tolist())) == py_oracle.kcount_histogram(open(os.path.join(kdir, f)).read().splitlines())
        assert values.min() >= 1 and freq.sum() == sum(1 for l in open(os.path.join(kdir, f)) if l.strip())
