"""Cross-check of the oracle against the reference's REAL external tools, when they are installed (test infrastructure).

The arithmetic of stages a2 (k-mer counting) and a4 (exact mapping) lives in executables the reference shells out to
(polycracker.py:199 kmercountexact.sh, :203 jellyfish, :252 bbmap.sh) and that this image does not carry, so the oracle
restates them from their documented semantics ("parity unpinned").  Wherever one of the tools IS on
PATH these functions run it with the reference's own command line on a small FASTA and return its result in the
oracle's terms, so that `tests/test_external_tools.py` can pin the restatement.  Nothing here is imported by the product.
"""
import os
import shutil
import subprocess
import tempfile

_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def available():
    """-> {tool: path or None} for the three executables of the path"""
    return {t: shutil.which(t) for t in ("jellyfish", "kmercountexact.sh", "bbmap.sh")}


def _canonical_min(kmer):
    rc = kmer.translate(_COMP)[::-1]
    return min(kmer, rc)


def write_fasta(path, records):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(">%s\n%s\n" % (name, seq))


def jellyfish_counts(records, k, threads=2):
    """polycracker.py:203: `jellyfish count -m K -s SIZE -t 15 -C -o mer_counts.jf FASTA && jellyfish dump mer_counts.jf -c`
    -> {canonical k-mer (lexicographic minimum of the two strands, jellyfish -C): count}"""
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa"); jf = os.path.join(d, "mer_counts.jf")
        write_fasta(fa, records)
        subprocess.run(["jellyfish", "count", "-m", str(k), "-s", str(max(os.stat(fa).st_size, 1000)), "-t", str(threads), "-C",
                        "-o", jf, fa], check=True, capture_output=True)
        out = subprocess.run(["jellyfish", "dump", jf, "-c"], check=True, capture_output=True, text=True).stdout
    counts = {}
    for line in out.splitlines():
        kmer, c = line.split()
        counts[_canonical_min(kmer.upper())] = int(c)
    return counts


def kmercountexact_counts(records, k, mincount=3, mem_gb=2):
    """polycracker.py:199: `kmercountexact.sh overwrite=true fastadump=f mincount=3 in=FASTA out=KCOUNT k=K -XmxMg`
    -> {canonical-minimum k-mer: count} (BBTools prints the lexicographic MAXIMUM of the two strands; relabelled here)"""
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa"); kc = os.path.join(d, "out.kcount")
        write_fasta(fa, records)
        subprocess.run(["kmercountexact.sh", "overwrite=true", "fastadump=f", "mincount=%d" % mincount, "in=" + fa, "out=" + kc,
                        "k=%d" % k, "-Xmx%dg" % mem_gb], check=True, capture_output=True)
        counts = {}
        with open(kc) as f:
            for line in f:
                if line.strip():
                    kmer, c = line.split()
                    counts[_canonical_min(kmer.upper())] = int(c)
    return counts


def bbmap_perfect_hits(ref_records, kmer_records, k, threads=2, mem_gb=2):
    """polycracker.py:252: `bbmap.sh vslow=t ambiguous=all noheader=t secondary=t k=K perfectmode=t threads=T
    maxsites=2000000000 outputunmapped=f ref=REF in=KMERS outm=SAM` -> sorted list of (query name, reference name): one
    per reported site, what blast2bed turns into the bed lines of the matrix."""
    with tempfile.TemporaryDirectory() as d:
        ref = os.path.join(d, "ref.fa"); q = os.path.join(d, "kmers.fa"); sam = os.path.join(d, "out.sam")
        write_fasta(ref, ref_records); write_fasta(q, kmer_records)
        subprocess.run(["bbmap.sh", "vslow=t", "ambiguous=all", "noheader=t", "secondary=t", "k=%d" % k, "perfectmode=t",
                        "threads=%d" % threads, "maxsites=2000000000", "outputunmapped=f", "ref=" + ref, "in=" + q, "outm=" + sam,
                        "path=" + d, "-Xmx%dg" % mem_gb], check=True, capture_output=True, cwd=d)
        hits = []
        with open(sam) as f:
            for line in f:
                if line.startswith("@") or not line.strip():
                    continue
                fields = line.split("\t")
                hits.append((fields[0], fields[2]))
    return sorted(hits)
