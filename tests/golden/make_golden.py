#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ by EXECUTING THE REFERENCE'S OWN PYTHON STAGE CODE.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is pinned by what
  * splitFasta / kmer2Fasta / blast2bed / findScaffolds / findKmerNames / generate_Kmer_Matrix are the reference's
    function bodies (polycracker/polycracker.py:113-171, 212-238, 255-296, 299-367), extracted as text, with the click
    decorators stripped and two mechanical Python-2 fixes (`print x` -> `print(x)`; `map` returns a list), then run
    here.  The external executables they call are replaced by small stand-ins with the documented semantics:
    pyfaidx.Fasta (writes the .fai), `bedtools sort` / `merge -c 4 -o collapse` / `getfasta -name`.
  * the two stages whose arithmetic lives in un-vendored tools -- kmercountexact.sh/jellyfish (a2) and bbmap.sh
    perfectmode=t (a4) -- cannot run here; their output files (.kcount, results.sam) are written by
    oracle/py_oracle.py in the tools' formats.  Those two stay "parity unpinned by the reference".
  * TF-IDF is sklearn's TfidfTransformer itself (polycracker.py:395-396).

The fixtures are small (a ~24 kb genome) and committed; tests/test_golden.py compares the oracle, the host-side stage
mirror and (on the GPU box) the CUDA path against them.
"""
import builtins
import os
import pickle
import re
import shutil
import sys
from collections import Counter

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/polycracker/polycracker.py"

from oracle import py_oracle  # noqa: E402
from polycracker_b200 import synth  # noqa: E402

CASES = {
    # name: (k, chunk_size, kmer_low_count, high_bool, kmer_high_count, sampling, remove_non_chunk, tfidf)
    "case_k23": dict(k=23, chunk_size=2000, kmer_low_count=4, high_bool=0, kmer_high_count=2000000, sampling=1,
                     remove_non_chunk=1, tfidf=1),
    "case_k11_high": dict(k=11, chunk_size=1500, kmer_low_count=6, high_bool=1, kmer_high_count=40, sampling=1,
                          remove_non_chunk=0, tfidf=1),
}


# ------------------------------------------------------------------------------------------------ reference code
def reference_functions():
    text = open(REF).read().split("\n")
    wanted = ["splitFasta", "kmer2Fasta", "blast2bed", "findScaffolds", "findKmerNames", "generate_Kmer_Matrix"]
    out = {}
    for name in wanted:
        start = next(i for i, l in enumerate(text) if l.startswith("def %s(" % name))
        end = next(i for i in range(start + 1, len(text))
                   if text[i].startswith("@") or text[i].startswith("def ") or text[i].startswith("class "))
        src = "\n".join(text[start:end])
        src = re.sub(r"^(\s*)print (?!\()(.+)$", r"\1print(\2)", src, flags=re.M)        # py2 print statement
        out[name] = (src, start + 1, end)
    return out


class _Interval:
    def __init__(self, fields):
        self.fields = fields

    def __len__(self):
        return int(self.fields[2]) - int(self.fields[1])


class BedTool:
    """Stand-in for pybedtools.BedTool with the three bedtools operations the hot path uses."""

    def __init__(self, src):
        if isinstance(src, str):
            self.rows = [l.rstrip("\n").split("\t") for l in open(src) if l.strip()]
        else:
            self.rows = [list(r) for r in src]

    def sort(self):                       # bedtools sort: chrom (bytewise), then start (numeric)
        return BedTool(sorted(self.rows, key=lambda r: (r[0].encode(), int(r[1]))))

    def merge(self, c=4, o="collapse"):   # bedtools merge -c 4 -o collapse on a sorted file
        assert c == 4 and o == "collapse"
        out = []
        for r in self.rows:
            if out and out[-1][0] == r[0] and int(r[1]) <= int(out[-1][2]):
                out[-1][2] = str(max(int(out[-1][2]), int(r[2])))
                out[-1][3] += "," + r[3]
            else:
                out.append([r[0], r[1], r[2], r[3]])
        return BedTool(out)

    def filter(self, fn):
        return BedTool([r for r in self.rows if fn(_Interval(r))])

    def saveas(self, path):
        with open(path, "w") as f:
            f.write("".join("\t".join(map(str, r)) + "\n" for r in self.rows))
        return self


class _Pybedtools:
    BedTool = BedTool


def _read_fasta(path):
    seqs, name = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs.append([name, []])
        elif name is not None:
            seqs[-1][1].append(line)
    return [(n, "".join(p)) for n, p in seqs]


def Fasta(path):
    """pyfaidx.Fasta(path): the reference only uses its side effect, the .fai index (polycracker.py:127-128)."""
    with open(path + ".fai", "w") as f:
        for n, s in _read_fasta(path):
            f.write("%s\t%d\t0\t60\t61\n" % (n, len(s)))


class _Subprocess:
    """`bedtools getfasta -fi X -fo Y -bed Z -name` (polycracker.py:163-170) and nothing else."""

    @staticmethod
    def call(cmd, shell=True):
        m = re.match(r"bedtools getfasta -fi (\S+) -fo (\S+) -bed (\S+) -name", cmd)
        assert m, "unexpected command in the reference hot path: %s" % cmd
        fi, fo, bed = m.groups()
        seqs = dict(_read_fasta(fi))
        with open(fo, "w") as f:
            for l in open(bed):
                if l.strip():
                    chrom, s, e, name = l.rstrip("\n").split("\t")[:4]
                    f.write(">%s\n%s\n" % (name, seqs[chrom][int(s):int(e)]))
        return 0


def reference_namespace():
    ns = dict(np=np, sps=sps, os=os, pickle=pickle, Counter=Counter, BedTool=BedTool, pybedtools=_Pybedtools,
              Fasta=Fasta, subprocess=_Subprocess,
              map=lambda f, *a: list(builtins.map(f, *a)))          # py2 map() returns a list
    for name, (src, _, _) in reference_functions().items():
        exec(compile(src, "%s:%s" % (REF, name), "exec"), ns)
    return ns


# ------------------------------------------------------------------------------------------------ fixtures
def make_genome(path):
    p = synth.small_params(seed=424242, genome_size=24_000, n_chrom=2, n_fraction=0.01)
    seq, off, names = synth.generate(p)
    text = seq.tobytes().decode()
    seqs = [(names[i], text[off[i]:off[i + 1]]) for i in range(len(names))]
    seqs.append(("short_scaf", "ACGTTGCATGCAAACCCGGGTTTACGTACGTAGCTAGCTAGGATCCGATCGATTAGC" * 9))   # < one chunk
    seqs.append(("n_only", "N" * 300))
    with open(path, "w") as f:
        for n, s in seqs:
            f.write(">%s some description\n" % n)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    return seqs


def run_case(name, cfg, ns):
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(os.path.join(out, "fasta_files"))
    os.makedirs(os.path.join(out, "kmercount_files"))
    os.makedirs(os.path.join(out, "blast_files"))
    cwd = os.getcwd()
    os.chdir(out)
    try:
        seqs = make_genome("fasta_files/genome.fa")
        # a1: the reference's splitFasta
        ns["splitFasta"]("genome.fa", "./fasta_files/", cfg["chunk_size"], 0, 0, cfg["remove_non_chunk"], 0, 0)
        records = _read_fasta("fasta_files/genome_split.fa")
        # a2: kmercountexact.sh stand-in (oracle) -> .kcount in the tool's "KMER<tab>COUNT" format
        counts = py_oracle.count_kmers(records, cfg["k"])
        with open("kmercount_files/genome_split.kcount", "w") as f:
            for km, c in py_oracle.dump_kcount(counts, cfg["k"]):
                f.write("%s\t%d\n" % (km, c))
        # a3: the reference's kmer2Fasta
        ns["kmer2Fasta"]("./kmercount_files/", cfg["kmer_low_count"], cfg["high_bool"], cfg["kmer_high_count"], cfg["sampling"])
        kmers = [l.strip() for l in open("kmercount_files/genome_split.kcount.fa").readlines()[1::2]]
        # a4: bbmap.sh perfectmode=t stand-in (oracle) -> headerless SAM
        hits = py_oracle.blast_kmers(kmers, records)
        with open("blast_files/results.sam", "w") as f:
            for q, flag, rname, pos in hits:
                f.write("\t".join([q, str(flag), rname, str(pos), "255", "%dM" % len(q), "*", "0", "0",
                                   q if flag == 0 else py_oracle.revcomp(q), "*"]) + "\n")
        # a5: the reference's blast2bed
        ns["blast2bed"]("./blast_files/results.sam", 1, 0, "blasted.bed", "blasted_merged.bed", False)
        # a6+a7: the reference's generate_Kmer_Matrix
        ns["generate_Kmer_Matrix"]("./kmercount_files/", "genome_split.fa", cfg["chunk_size"], 0, cfg["remove_non_chunk"], 0, 0,
                                   0, "./fasta_files/", "genome_split.fa", "genome.fa", "blasted.bed", "blasted_merged.bed",
                                   "clusteringMatrix.npz", "kmers.p", cfg["tfidf"])
        # a8: the reference's TF-IDF call (polycracker.py:395-396)
        from sklearn.feature_extraction.text import TfidfTransformer
        data = sps.load_npz("clusteringMatrix.npz")
        if data.shape[0] and data.shape[1]:
            tf = sps.csr_matrix(TfidfTransformer().fit_transform(data))
            tf.sort_indices()
        else:
            tf = sps.csr_matrix(data.shape, dtype=np.float32)
        sps.save_npz("clusteringMatrix.tfidf.npz", tf)
        os.remove("blasted.bed")                       # large and redundant with blasted_merged.bed
        os.remove("blast_files/results.sam")           # the oracle's stand-in for bbmap output; tests regenerate it
        os.rmdir("blast_files")
        os.remove("fasta_files/genome.fa.fai")
        with open("params.txt", "w") as f:
            f.write("\n".join("%s = %s" % kv for kv in sorted(cfg.items())) + "\n")
        print(name, "matrix", data.shape, "nnz", data.nnz, "kmers", len(kmers), "hits", len(hits))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    fns = reference_functions()
    with open(os.path.join(HERE, "REFERENCE_LINES.txt"), "w") as f:
        f.write("reference functions executed by make_golden.py (polycracker/polycracker.py):\n")
        for n, (_, a, b) in fns.items():
            f.write("  %s: lines %d-%d\n" % (n, a, b))
    ns = reference_namespace()
    for case, cfg in CASES.items():
        run_case(case, cfg, ns)
