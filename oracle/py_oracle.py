"""Literal CPU restatement of polyCRACKER's hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``polycracker_b200``) never does.

PARITY STATUS: **parity unpinned by the reference for the external-tool stages**.  The reference
(``/root/reference``) ships no tests, no golden vectors and no test FASTA, and the arithmetic of
stages a2 (k-mer counting) and a4 (exact mapping) lives in un-vendored, un-pinned third-party
executables (BBTools ``kmercountexact.sh`` / ``bbmap.sh``, ``jellyfish``, ``bedtools``; see
SURVEY.md section 8c).  Those two stages are restated here from the tools' documented semantics and
from the reference's own call sites.  The Python-side stages (a3, a5, a6, a7) and the TF-IDF (a8) ARE
pinned: ``tests/golden/make_golden.py`` executes the reference's own function bodies
(polycracker.py:212-238, 255-281, 284-367) on this oracle's a2/a4 outputs and the committed fixtures
are compared against in ``tests/test_golden.py``; a8 is sklearn's ``TfidfTransformer`` itself.

Every function cites the reference lines it follows (paths relative to /root/reference/polycracker/).
Everything is deliberately slow, string-based and obvious.
"""
from collections import Counter

import numpy as np
import scipy.sparse as sps

_COMP = str.maketrans("ACGT", "TGCA")


def revcomp(kmer):
    return kmer.translate(_COMP)[::-1]


def canonical(kmer, mode="min"):
    """Strand-collapsed representative.  jellyfish -C prints the lexicographic minimum
    (polycracker.py:203); BBTools is believed to keep the maximum.  Counts are identical
    either way; only the label differs."""
    rc = revcomp(kmer)
    return min(kmer, rc) if mode == "min" else max(kmer, rc)


# --------------------------------------------------------------------------------------------- a1
def split_chunks(fai, chunk_size):
    """polycracker.py:129-156.  ``fai`` = [(name, length)].  Returns bed rows
    (chrom, start, end, name) in ``bedtools sort`` order (chrom bytewise, start numeric)."""
    rows = []
    for key, length in fai:
        length = int(length)
        if not length:                       # split(): "if inputStr: ... else ''"
            continue
        positions = list(range(0, length, chunk_size))        # np.arange(0, L, C)
        if length > chunk_size:
            for i in range(len(positions) - 1):
                s, e = positions[i], positions[i + 1]
                rows.append((key, s, e, "%s_%d_%d" % (key, s, e)))
        s = positions[-1]
        rows.append((key, s, length, "%s_%d_%d" % (key, s, length)))
    rows.sort(key=lambda r: (r[0].encode(), r[1]))            # BedTool(...).sort(), pc.py:154
    return rows


def chunk_records(sequences, bed_rows):
    """bedtools getfasta -name (pc.py:170): one FASTA record per bed row, case preserved."""
    seqs = dict(sequences)
    return [(name, seqs[chrom][s:e]) for chrom, s, e, name in bed_rows]


def find_scaffolds(bed_rows):
    """polycracker.py:284-287: Python-sorted 4th column of correspondence.bed."""
    return sorted(r[3] for r in bed_rows)


def filter_scaffolds(original_scaffolds, chunk_size, remove_non_chunk=1, min_chunk_threshold=0,
                     min_chunk_size=0, prefilter=0):
    """polycracker.py:322-327."""
    def clen(s):
        p = s.split("_")
        return abs(int(p[-1]) - int(p[-2]))
    if remove_non_chunk and prefilter == 0:
        return [s for s in original_scaffolds if clen(s) == chunk_size]
    if min_chunk_threshold and prefilter == 0:
        return [s for s in original_scaffolds if clen(s) >= min_chunk_size]
    return list(original_scaffolds)


# --------------------------------------------------------------------------------------------- a2
def count_kmers(records, k, mode="min"):
    """Canonical exact k-mer counts over every record independently (polycracker.py:197-203:
    kmercountexact.sh for k<=31, jellyfish -C for k>31).  No window spans two records; windows
    holding any non-ACGT character are skipped; lower case counts as upper case; a palindromic
    k-mer is counted once per window."""
    counts = Counter()
    for _, seq in records:
        seq = seq.upper()
        for p in range(len(seq) - k + 1):
            w = seq[p:p + k]
            if all(c in "ACGT" for c in w):
                counts[canonical(w, mode)] += 1
    return counts


def dump_kcount(counts, k):
    """Lines of the .kcount file: ``mincount=3`` for k<=31 (pc.py:199), everything for k>31
    (pc.py:203, no -L).  Returned sorted for determinism (the tools' dump order is unspecified)."""
    mincount = 3 if k <= 31 else 1
    return sorted((km, c) for km, c in counts.items() if c >= mincount)


# --------------------------------------------------------------------------------------------- a3
def kmer2fasta(kcount_lines, kmer_low_count=100, high_bool=0, kmer_high_count=2000000,
               sampling_sensitivity=1):
    """polycracker.py:220-238.  ``>=`` low and ``<=`` high, both inclusive; every
    ``sampling_sensitivity``-th survivor is kept.  Declared deviation: survivors are visited in
    sorted order (the reference visits them in the counter's unspecified dump order)."""
    if sampling_sensitivity < 1:
        sampling_sensitivity = 1
    out, count = [], 0
    for kmer, c in sorted(kcount_lines):
        keep = c >= kmer_low_count and (not high_bool or c <= kmer_high_count)
        if keep:
            if count % sampling_sensitivity == 0:
                out.append(kmer)
            count += 1
    return out


# --------------------------------------------------------------------------------------------- a4
def blast_kmers(kmers, records, mode="min"):
    """bbmap.sh perfectmode=t ambiguous=all secondary=t maxsites=2e9 (polycracker.py:252): one
    record per (query k-mer, site) where the k-mer or its reverse complement matches exactly.
    Returns SAM-like tuples (qname, flag, rname, pos1)."""
    by_len = {}
    for q in kmers:
        by_len.setdefault(len(q), set()).add(q)
    hits = []
    for name, seq in records:
        seq = seq.upper()
        for k, qset in by_len.items():
            for p in range(len(seq) - k + 1):
                w = seq[p:p + k]
                if w in qset:
                    hits.append((w, 0, name, p + 1))
                else:
                    rc = revcomp(w)
                    if rc in qset:                      # palindromes: w == rc, reported once
                        hits.append((rc, 16, name, p + 1))
    return hits


# --------------------------------------------------------------------------------------------- a5
def blast2bed(hits):
    """polycracker.py:268-272 + 280: chunk = RNAME.split('::')[0]; sort+merge(c=4,o=collapse) gives
    one line per chunk with >=1 hit: (chunk, 0, len, 'k1,k2,...')."""
    per_chunk = {}
    for q, _flag, rname, _pos in hits:
        l1 = rname.split("::")[0]
        per_chunk.setdefault(l1, []).append(q)
    merged = []
    for l1 in sorted(per_chunk, key=lambda s: s.encode()):
        p = l1.split("_")
        merged.append((l1, 0, int(p[-1]) - int(p[-2]), ",".join(per_chunk[l1])))
    return merged


# --------------------------------------------------------------------------------------------- a6/a7
def generate_kmer_matrix(merged_bed, original_scaffolds, kmers, chunk_size, remove_non_chunk=1,
                         min_chunk_threshold=0, min_chunk_size=0, prefilter=0, tfidf=1):
    """polycracker.py:320-359.  Returns (csc float32 matrix, scaffolds, sorted kmers)."""
    kmers = sorted(kmers)                                     # findKmerNames, pc.py:295
    scaffolds = filter_scaffolds(original_scaffolds, chunk_size, remove_non_chunk,
                                 min_chunk_threshold, min_chunk_size, prefilter)
    scaffold_idx = {s: i for i, s in enumerate(scaffolds)}
    kmer_idx = {km: i for i, km in enumerate(kmers)}
    data = sps.dok_matrix((len(scaffolds), len(kmers)), dtype=np.float32)
    for chunk, _s, _e, names in merged_bed:
        if chunk in scaffold_idx:
            for key, c in Counter(names.split(",")).items():
                if key in kmer_idx:
                    data[scaffold_idx[chunk], kmer_idx[key]] = c
    data = data.tocsc()
    if not tfidf and len(scaffolds):
        lens = np.array([[float(x) for x in s.split("_")[-2:]] for s in scaffolds])
        scale = np.abs(lens[:, 1] - lens[:, 0]) / 5000.
        data = data.tolil()
        for i in range(len(scaffolds)):
            data[i, :] = data[i, :] / scale[i]                # pc.py:357-359 (<=1 ulp)
        data = data.tocsc().astype(np.float32)
    return data, scaffolds, kmers


# --------------------------------------------------------------------------------------------- a8
def tfidf(matrix):
    """polycracker.py:395-396: sklearn TfidfTransformer defaults (smooth idf, L2 rows)."""
    from sklearn.feature_extraction.text import TfidfTransformer
    if matrix.shape[1] == 0 or matrix.shape[0] == 0:
        return sps.csr_matrix(matrix.shape, dtype=np.float32)
    return sps.csr_matrix(TfidfTransformer().fit_transform(matrix))


# --------------------------------------------------------------------------------------------- whole path
def run_path(sequences, chunk_size, k, kmer_low_count, high_bool=0, kmer_high_count=2000000,
             sampling_sensitivity=1, remove_non_chunk=1, min_chunk_threshold=0, min_chunk_size=0,
             mode="min"):
    """The reference's stage chain a1..a8 on in-memory data.  ``sequences`` = [(name, str)]."""
    bed = split_chunks([(n, len(s)) for n, s in sequences], chunk_size)
    records = chunk_records(sequences, bed)
    counts = count_kmers(records, k, mode)
    kcount = dump_kcount(counts, k)
    kmers = kmer2fasta(kcount, kmer_low_count, high_bool, kmer_high_count, sampling_sensitivity)
    hits = blast_kmers(kmers, records, mode)
    merged = blast2bed(hits)
    original_scaffolds = find_scaffolds(bed)
    mat, scaffolds, kmers = generate_kmer_matrix(merged, original_scaffolds, kmers, chunk_size,
                                                 remove_non_chunk, min_chunk_threshold,
                                                 min_chunk_size)
    return dict(bed=bed, counts=counts, kcount=kcount, kmers=kmers, scaffolds=scaffolds, merged=merged,
                original_scaffolds=original_scaffolds, matrix=mat, tfidf=tfidf(mat))


# ---- N2 diagnostics (polycracker.py:1366-1382, 1419-1496) ----------------------------------------------------------
def repeatmers_per_subsequence(merged_bed_lines):
    """awk '{print gsub(/,/,"")+1}' over blasted_merged.bed (polycracker.py:1374): k-mer names per line."""
    return [line.count(",") + 1 for line in merged_bed_lines]


def kcount_histogram(kcount_lines):
    """OrderedDict(sorted(Counter(second column).items())) (polycracker.py:1466-1469) -> [(count, n_kmers)]."""
    from collections import Counter
    c = Counter(int(line.split()[1]) for line in kcount_lines if line.strip())
    return sorted(c.items())
