#!/usr/bin/env python
"""Golden fixtures for scope row N1 (subgenome extraction) made by EXECUTING THE REFERENCE'S OWN PYTHON CODE.

Run in the build container only (reads /root/reference):

    python tests/golden/make_golden_sub.py

What runs
  * the reference's function bodies fai2bed, create_path, writeKmerCountSubgenome, kmercounttodict, compareKmers,
    writeBlast, blast2bed3, bed2unionBed, kmerratio2scaffasta and subgenomeExtraction
    (polycracker/polycracker.py:57-65, 676-1020), extracted as text with the click decorators stripped and the
    mechanical Python-2 fixes listed in `_py2_fixes` -- including the bootstrap recursion through ``ctx.invoke``.
  * the external executables they shell out to are replaced by stand-ins with the tools' documented semantics:
    `kmercountexact.sh` / `jellyfish` and `bbmap.sh perfectmode=t` (oracle/py_oracle.py, as for the hot path -- these
    two stay "parity unpinned by the reference"), `bedtools getfasta -name`, `cat`, `rm`, `cut -f 1-2`,
    `reformat.sh` (ignored: it only re-wraps FASTA lines), pybedtools sort / coverage / union_bedgraphs, pyfaidx.Fasta.
  * ``os.listdir`` is sorted (the reference takes directory order, which decides nothing but the order of the
    "other subgenome" counts in FASTA headers and which union-bedgraph column is which subgenome).

Outputs kept per case under tests/golden/<case>/: the inputs (genome.fa, bootstrap_0/subgenome_*.txt, params.txt) and
every artefact a later stage or a user reads: *.higher.kmers.fa, *.sorted.cov.bedgraph, subgenomes.union.bedgraph,
bootstrap_<i>/subgenome_*.txt, finalResults/subgenome_*.txt, extractedSubgenomes/*.names.
"""
import builtins
import errno
import os
import re
import shutil
import sys
import tempfile
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference/polycracker/polycracker.py"

import make_golden as mg  # noqa: E402
from oracle import py_oracle  # noqa: E402
from polycracker_b200 import synth  # noqa: E402

CASES = {
    "sub_two_k15_k23": dict(n_subgenomes=2, seed=77001, genome_size=60_000, chunk_size=2000, kmer_length="15,23",
                            diff_kmer_threshold=2, default_kmercount_value=3., diff_sample_rate=1,
                            unionbed_threshold="3,2", bootstrap=1, drop_every=4),
    "sub_three_k23_k32": dict(n_subgenomes=3, seed=77002, genome_size=72_000, chunk_size=3000, kmer_length="23,32",
                              diff_kmer_threshold=1, default_kmercount_value=2., diff_sample_rate=1,
                              unionbed_threshold="2,1", bootstrap=0, drop_every=5),
}

WANTED = ["create_path", "fai2bed", "writeKmerCountSubgenome", "kmercounttodict", "compareKmers", "writeBlast",
          "blast2bed3", "bed2unionBed", "kmerratio2scaffasta", "subgenomeExtraction"]


def _py2_fixes(src):
    src = re.sub(r"^(\s*)print (?!\()(.+)$", r"\1print(\2)", src, flags=re.M)        # print statement
    src = src.replace(".iteritems()", ".items()")                                       # dict.iteritems
    src = src.replace("(val1 / val2)", "_py2div(val1, val2)")                           # int / int floors in Python 2
    return src


def _py2div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a // b
    return a / b


def reference_functions():
    text = open(REF).read().split("\n")
    out = {}
    for name in WANTED:
        start = next(i for i, l in enumerate(text) if l.startswith("def %s(" % name))
        end = next(i for i in range(start + 1, len(text))
                   if text[i].startswith("@") or text[i].startswith("def ") or text[i].startswith("class ")
                   or text[i].startswith("#"))
        out[name] = (_py2_fixes("\n".join(text[start:end])), start + 1, end)
    return out


# ------------------------------------------------------------------------------------------------ stand-ins
class BedTool(mg.BedTool):
    def __init__(self, src=None):
        if src is None:
            self.rows = []
        else:
            mg.BedTool.__init__(self, src)

    def sort(self):
        return BedTool(sorted(self.rows, key=lambda r: (r[0].encode(), int(r[1]))))

    @property
    def fn(self):
        f = tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False)
        f.write("".join("\t".join(map(str, r)) + "\n" for r in self.rows))
        f.close()
        return f.name

    def coverage(self, b):
        """bedtools coverage -a self -b b: A's columns + [#overlapping B features, #bases of A covered, len(A), frac]."""
        per = defaultdict(list)
        for r in b.rows:
            per[r[0]].append((int(r[1]), int(r[2])))
        out = []
        for r in self.rows:
            s, e = int(r[1]), int(r[2])
            n, covered = 0, set()
            for bs, be in per.get(r[0], ()):
                if bs < e and be > s:
                    n += 1
                    covered.update(range(max(bs, s), min(be, e)))
            out.append(list(r) + [str(n), str(len(covered)), str(e - s), "%.7f" % (len(covered) / float(e - s) if e > s else 0)])
        return BedTool(out)

    def union_bedgraphs(self, i, g, empty=True):
        """bedtools unionbedg -i ... -g genome -empty on bedgraphs whose intervals are whole chunks: one line per
        chunk, one value column per input, '0' where an input has no interval."""
        assert empty
        files = [dict(((r[0], r[1], r[2]), r[3]) for r in BedTool(fn).rows) for fn in i]
        keys = sorted(set().union(*[set(f) for f in files]), key=lambda t: (t[0].encode(), int(t[1])))
        lens = dict(l.split()[:2] for l in open(g) if l.strip())
        for chrom, s, e in keys:
            assert s == "0" and e == lens[chrom], "stand-in only handles whole-chunk intervals"
        return BedTool([[c, s, e] + [f.get((c, s, e), "0") for f in files] for c, s, e in keys])


class _Record:
    def __init__(self, s):
        self.s = s

    def __getitem__(self, sl):
        return self.s[sl]


class Fasta:
    """pyfaidx.Fasta: writes the .fai, indexable by record name."""

    def __init__(self, path):
        self.seqs = dict(mg._read_fasta(path))
        with open(path + ".fai", "w") as f:
            for n, s in mg._read_fasta(path):
                f.write("%s\t%d\t0\t60\t61\n" % (n, len(s)))

    def __getitem__(self, name):
        return _Record(self.seqs[name])


class _Subprocess:
    @staticmethod
    def call(cmd, shell=True):
        m = re.match(r"bedtools getfasta -fi (\S+) -fo (\S+) -bed (\S+) -name", cmd)
        if m:
            return mg._Subprocess.call(cmd)
        m = re.search(r"kmercountexact\.sh overwrite=true fastadump=f mincount=3 in=(\S+) out=(\S+) k=(\d+)", cmd)
        if m:
            fi, fo, k = m.group(1), m.group(2), int(m.group(3))
            counts = py_oracle.count_kmers(mg._read_fasta(fi), k)
            with open(fo, "w") as f:
                for km, c in py_oracle.dump_kcount(counts, k):
                    f.write("%s\t%d\n" % (km, c))
            return 0
        m = re.match(r"jellyfish count -m (\d+) -s \d+ -t 15 -C -o \S+ (\S+) && jellyfish dump \S+ -c > (\S+)", cmd)
        if m:
            k, fi, fo = int(m.group(1)), m.group(2), m.group(3)
            counts = py_oracle.count_kmers(mg._read_fasta(fi), k)
            with open(fo, "w") as f:
                for km, c in py_oracle.dump_kcount(counts, k):
                    f.write("%s %d\n" % (km, c))
            return 0
        m = re.match(r"cat (.+) > (\S+)$", cmd)
        if m:
            with open(m.group(2), "w") as f:
                for p in m.group(1).split():
                    f.write(open(p).read())
            return 0
        m = re.match(r"rm (.+)$", cmd)
        if m:
            for p in m.group(1).split():
                os.remove(p)
            return 0
        m = re.match(r"cut -f 1-2 (\S+) > (\S+)$", cmd)
        if m:
            with open(m.group(2), "w") as f:
                for l in open(m.group(1)):
                    f.write("\t".join(l.rstrip("\n").split("\t")[:2]) + "\n")
            return 0
        m = re.search(r"bbmap\.sh .*perfectmode=t .*ref=(\S+) in=(\S+) path=\S+ outm=(\S+)", cmd)
        if m:
            ref, fi, fo = m.groups()
            queries = mg._read_fasta(fi)                                   # header = kmer.val1.values, sequence = kmer
            by_seq = {s: n for n, s in queries}
            hits = py_oracle.blast_kmers([s for _, s in queries], mg._read_fasta(ref))
            with open(fo, "w") as f:
                for q, flag, rname, pos in hits:
                    f.write("\t".join([by_seq[q], str(flag), rname, str(pos), "3", "%d=" % len(q), "*", "0", "0",
                                       q if flag == 0 else py_oracle.revcomp(q), "*", "XT:A:R", "NM:i:0", "AM:i:3"]) + "\n")
            return 0
        if "reformat.sh" in cmd:
            return 0                                                       # only re-wraps FASTA lines
        raise AssertionError("unexpected command in the reference subgenome path: %s" % cmd)


class _Os:
    """os with a sorted listdir (see module docstring)."""

    def __getattr__(self, name):
        return getattr(os, name)

    @staticmethod
    def listdir(p):
        return sorted(os.listdir(p))


class _Ctx:
    """click context: ctx.invoke(command, **kwargs) calls the command's function with the context."""

    def invoke(self, fn, **kw):
        return fn(self, **kw)


def reference_namespace():
    ns = dict(os=_Os(), errno=errno, shutil=shutil, defaultdict=defaultdict, BedTool=BedTool, Fasta=Fasta,
              subprocess=_Subprocess, _py2div=_py2div,
              map=lambda f, *a: list(builtins.map(f, *a)))
    for name, (src, _, _) in reference_functions().items():
        exec(compile(src, "%s:%s" % (REF, name), "exec"), ns)
    return ns


# ------------------------------------------------------------------------------------------------ fixtures
def run_case(name, cfg, ns, ns_hot):
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(os.path.join(out, "fasta_files"))
    cwd = os.getcwd()
    os.chdir(out)
    try:
        p = synth.small_params(seed=cfg["seed"], genome_size=cfg["genome_size"], n_subgenomes=cfg["n_subgenomes"],
                               n_chrom=2, n_shared_families=3, n_specific_families=3 * cfg["n_subgenomes"],
                               min_family_len=120, max_family_len=400, n_fraction=0.005)
        seq, off, names = synth.generate(p)
        text = seq.tobytes().decode()
        with open("fasta_files/genome.fa", "w") as f:
            for i, n in enumerate(names):
                s = text[off[i]:off[i + 1]]
                f.write(">%s\n" % n)
                for j in range(0, len(s), 60):
                    f.write(s[j:j + 60] + "\n")
        ns_hot["splitFasta"]("genome.fa", "./fasta_files/", cfg["chunk_size"], 0, 0, 1, 0, 0)
        chunks = [n for n, _ in mg._read_fasta("fasta_files/genome_split.fa")]
        # starting clusters: the true subgenome of each chunk, with every drop_every-th chunk left unassigned and one
        # chunk deliberately put into the wrong cluster
        os.makedirs("analysisOutputs/bootstrap_0")
        letters = "ABDEFGH"
        groups = [[] for _ in range(cfg["n_subgenomes"])]
        for i, c in enumerate(sorted(chunks)):
            if i % cfg["drop_every"] == 0:
                continue
            g = letters.index(c[3])
            if i == 7:
                g = (g + 1) % cfg["n_subgenomes"]
            groups[g].append(c)
        for g, members in enumerate(groups):
            with open("analysisOutputs/bootstrap_0/subgenome_%d.txt" % g, "w") as f:
                f.write("\n".join(members))
        ns["subgenomeExtraction"](_Ctx(), subgenome_folder="analysisOutputs/bootstrap_0",
                                  original_subgenome_path="analysisOutputs", fasta_path="./fasta_files/",
                                  genome_name="genome_split.fa", original_genome="genome_split.fa", bb=1,
                                  bootstrap=cfg["bootstrap"], iteration=0, kmer_length=cfg["kmer_length"], run_final=0,
                                  original=0, blast_mem="100", kmer_low_count=100,
                                  diff_kmer_threshold=cfg["diff_kmer_threshold"],
                                  unionbed_threshold=cfg["unionbed_threshold"], diff_sample_rate=cfg["diff_sample_rate"],
                                  default_kmercount_value=cfg["default_kmercount_value"], search_length=13,
                                  perfect_mode=1)
        # keep the artefacts, drop the bulky intermediates
        summary = []
        for root, dirs, files in os.walk("analysisOutputs"):
            for fn in files:
                path = os.path.join(root, fn)
                keep = (fn.endswith(".higher.kmers.fa") or fn.endswith(".sorted.cov.bedgraph")
                        or fn == "subgenomes.union.bedgraph" or re.match(r"subgenome_\d+\.txt$", fn))
                if "extractedSubgenomes" in root and fn.endswith(".fasta") and "_wrapped" not in fn:
                    with open(path[:-len(".fasta")] + ".names", "w") as f:
                        f.write("".join(l[1:] for l in open(path) if l.startswith(">")))
                if not keep:
                    os.remove(path)
                else:
                    summary.append("%s %d" % (path, os.path.getsize(path)))
        for root, dirs, files in os.walk("analysisOutputs", topdown=False):
            if not os.listdir(root):
                os.rmdir(root)
        for fn in os.listdir("fasta_files"):
            if fn != "genome.fa":
                os.remove(os.path.join("fasta_files", fn))
        for fn in os.listdir("."):
            if os.path.isfile(fn):
                os.remove(fn)                                  # correspondence beds of splitFasta
        with open("params.txt", "w") as f:
            f.write("\n".join("%s = %s" % kv for kv in sorted(cfg.items())) + "\n")
        print(name, "\n  " + "\n  ".join(sorted(summary)))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    fns = reference_functions()
    with open(os.path.join(HERE, "REFERENCE_LINES_SUB.txt"), "w") as f:
        f.write("reference functions executed by make_golden_sub.py (polycracker/polycracker.py):\n")
        for n, (_, a, b) in fns.items():
            f.write("  %s: lines %d-%d\n" % (n, a, b))
    ns = reference_namespace()
    ns_hot = mg.reference_namespace()
    for case, cfg in CASES.items():
        run_case(case, cfg, ns, ns_hot)
