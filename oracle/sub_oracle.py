"""CPU restatement of polyCRACKER's subgenome-extraction inner loop (scope row N1)  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` (and the golden generator under ``tests/golden/``) may import this module; the product never does.

PARITY STATUS.  ``compare_kmers`` (compareKmers, polycracker.py:735-783), the SAM -> bed3 -> coverage -> bedgraph
conversion (blast2bed3, 802-865), the union bedgraph (bed2unionBed, 868-879) and the binning rule
(kmerratio2scaffasta, 884-922) are PINNED: ``tests/golden/make_golden_sub.py`` executes the reference's own function
bodies for them and ``tests/test_golden_sub.py`` compares this module against the committed outputs.  The two
external-tool steps -- per-subgenome ``kmercountexact.sh`` / ``jellyfish`` (692-719) and ``bbmap.sh perfectmode=t``
(786-799) -- are the same un-vendored tools as hot-path stages a2 / a4 and stay "parity unpinned by the reference"
(restated in ``py_oracle.count_kmers`` / ``py_oracle.blast_kmers``).

Two flavours: literal string/dict code for small cases (``*_literal``) and numpy integer code for Mbp-sized cases.
Paths cited are relative to /root/reference/polycracker/.
"""
import numpy as np

from . import py_oracle


# ------------------------------------------------------------------------------------------------ literal (small)
def subgenome_dictionaries(group_records, kmer_lengths):
    """writeKmerCountSubgenome (polycracker.py:692-719) + kmercounttodict (722-733): per subgenome FASTA, the dumps of
    every k-mer length concatenated into one file, read back as one {kmer: count} dictionary.  The dump floor is
    ``mincount=3`` for k <= 31 and none for k > 31 (711-714), exactly as in the hot path (py_oracle.dump_kcount)."""
    dicts = []
    for records in group_records:
        d = {}
        for k in kmer_lengths:
            for kmer, c in py_oracle.dump_kcount(py_oracle.count_kmers(records, k), k):
                d[kmer] = c
        dicts.append(d)
    return dicts


def py2_div(a, b):
    """Python 2 ``/``: floor for int / int, true division as soon as one side is a float (pc.py:765, 778)."""
    if isinstance(a, int) and isinstance(b, int):
        return a // b
    return a / b


def compare_kmers(dicts, kmer_lengths, diff_kmer_threshold=20, default_kmercount_value=3., diff_sample_rate=1):
    """compareKmers (polycracker.py:735-783).  For every k-mer of subgenome g: the counts in the other dictionaries
    (``default_kmercount_value`` where absent), differential if ANY quotient exceeds the threshold.
    Returns per subgenome a list of (kmer, val1, [other values]) -- ordered by k-mer length (in ``kmer_lengths``
    order) then k-mer.  Declared deviation: with diff_sample_rate > 1 the reference keeps every n-th survivor in
    Python-2 dict order (unreproducible); here every n-th survivor of each (subgenome, length) list in sorted order."""
    if diff_sample_rate < 1:
        diff_sample_rate = 1
    out = []
    for g, d1 in enumerate(dicts):
        others = [d for o, d in enumerate(dicts) if o != g]
        mine = []
        for k in kmer_lengths:
            count = 0
            for kmer in sorted(km for km in d1 if len(km) == k):
                val1 = d1[kmer]
                values = [d.get(kmer, default_kmercount_value) for d in others]
                if any(py2_div(val1, val2) > diff_kmer_threshold for val2 in values):
                    if count % diff_sample_rate == 0:
                        mine.append((kmer, val1, values))
                    count += 1
        out.append(mine)
    return out


def higher_kmers_fasta(diff, sampled=False):
    """The text of ``<subgenome>.higher.kmers.fa`` (pc.py:767 sampled branch prints floats with str(), 781 with
    str(int(x)))."""
    fmt = (lambda x: str(x)) if sampled else (lambda x: str(int(x)))
    return "".join(">%s.%d.%s\n%s\n" % (kmer, val1, ".".join(fmt(v) for v in values), kmer)
                   for kmer, val1, values in diff)


def map_back_literal(diff_kmers, records):
    """writeBlast (786-799; bbmap perfectmode, every site, either strand) + blast2bed3 (802-865): each SAM record
    becomes the interval [POS, POS+1) on its chunk; ``bedtools coverage`` of the whole-chunk windows then reports the
    number of covered bases (field 6 of the 4-column windows, pc.py:861) = distinct POS values.
    Returns {chunk name: covered bases} with 0 for chunks without hits (coverage lists every window)."""
    hits = py_oracle.blast_kmers(diff_kmers, records)
    seen = {name: set() for name, _ in records}
    for _q, _flag, rname, pos1 in hits:
        seen[rname].add(pos1)
    return {name: len(s) for name, s in seen.items()}


def bin_chunks(chunk_names, coverage, absolute_threshold=10, ratio_threshold=2):
    """kmerratio2scaffasta (polycracker.py:908-922).  coverage[i][g] = union-bedgraph value of chunk i for subgenome g.
    Returns (per-subgenome chunk lists, ambiguous chunk list), in the order given."""
    n_groups = len(coverage[0]) if len(coverage) else 0
    out = [[] for _ in range(n_groups)]
    ambiguous = []
    for name, row in zip(chunk_names, coverage):
        x = [float(v) for v in row]
        amb = 1
        for i in range(len(x)):
            x_others = x[:i] + x[i + 1:]
            if all((x_i == 0 and x[i] > absolute_threshold) or (x_i > 0 and (x[i] / x_i) > ratio_threshold)
                   for x_i in x_others):
                out[i].append(name)
                amb = 0
        if amb:
            ambiguous.append(name)
    return out, ambiguous


# ------------------------------------------------------------------------------------------------ numpy (Mbp sizes)
_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def window_kmers(chunk, k):
    """canonical (min orientation) 2-bit k-mer of every window start of one chunk + validity, as uint64 / bool arrays
    (A=0 C=1 G=2 T=3, first base most significant -- the order in which ascending keys equal sorted strings)."""
    code = _CODE[np.asarray(chunk, np.uint8)]
    n = len(code) - k + 1
    if n <= 0:
        return np.zeros(0, np.uint64), np.zeros(0, bool)
    bad = (code > 3).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(bad)])
    valid = (csum[k:] - csum[:-k]) == 0
    c = (code & 3).astype(np.uint64)
    fwd = np.zeros(n, np.uint64)
    rc = np.zeros(n, np.uint64)
    for j in range(k):
        fwd = (fwd << np.uint64(2)) | c[j:j + n]
        rc |= (np.uint64(3) - c[j:j + n]) << np.uint64(2 * j)
    return np.minimum(fwd, rc), valid


def count_group(seq, chunk_begin, chunk_end, k, min_count):
    """sorted (keys, counts) with count >= min_count over the given chunks (no window spans two chunks)."""
    parts = []
    for b, e in zip(chunk_begin, chunk_end):
        keys, valid = window_kmers(seq[b:e], k)
        parts.append(keys[valid])
    allk = np.concatenate(parts) if parts else np.zeros(0, np.uint64)
    keys, counts = np.unique(allk, return_counts=True)
    keep = counts >= min_count
    return keys[keep], counts[keep].astype(np.int64)


def compare_arrays(dicts, diff_kmer_threshold, default_kmercount_value=3., diff_sample_rate=1):
    """compare_kmers on sorted (keys, counts) arrays of ONE k-mer length -> per subgenome (keys, counts[n, G])."""
    rate = max(1, int(diff_sample_rate))
    out = []
    for g, (keys, counts) in enumerate(dicts):
        passed = np.zeros(len(keys), bool)
        table = np.zeros((len(keys), len(dicts)), np.int64)
        for o, (okeys, ocounts) in enumerate(dicts):
            pos = np.searchsorted(okeys, keys)
            pos_c = np.minimum(pos, max(len(okeys) - 1, 0))
            found = (okeys[pos_c] == keys) if len(okeys) else np.zeros(len(keys), bool)
            val2 = np.where(found, ocounts[pos_c] if len(okeys) else 0, 0)
            table[:, o] = val2
            if o == g:
                continue
            q_int = counts // np.maximum(val2, 1)                      # Python 2 int / int
            q_flt = counts / float(default_kmercount_value)             # int / float
            passed |= np.where(found, q_int > diff_kmer_threshold, q_flt > diff_kmer_threshold)
        idx = np.flatnonzero(passed)[::rate]
        out.append((keys[idx], table[idx]))
    return out


def map_back(seq, chunk_begin, chunk_end, diff_by_k):
    """diff_by_k: {k: [sorted key array per subgenome]} -> int64[n_chunks, G] distinct matching window starts."""
    n_groups = len(next(iter(diff_by_k.values()))) if diff_by_k else 0
    cov = np.zeros((len(chunk_begin), n_groups), np.int64)
    for i, (b, e) in enumerate(zip(chunk_begin, chunk_end)):
        chunk = seq[b:e]
        hit = np.zeros((n_groups, max(len(chunk), 1)), bool)
        for k, sets in diff_by_k.items():
            keys, valid = window_kmers(chunk, k)
            for g, dk in enumerate(sets):
                if len(keys) and len(dk):
                    hit[g, :len(keys)] |= valid & np.isin(keys, dk)
        cov[i] = hit.sum(axis=1)
    return cov


# ------------------------------------------------------------------------------------------------ N4 (bio_hyp_class)
def kmer_genome_matrix_literal(group_records, k, min_count=3, ratio_threshold=20, default_kcount_value=1, max_val=100):
    """`bio_hyp_class` (utilities.py:177-284), the data path only, restated literally: per genome the counter's dump with
    ``mincount=min_count`` (utilities.py:219-222) read back as a dict (215, 223); union of the keys (228), LabelEncoder =
    sorted unique strings (230-231); genomes x k-mers int matrix with implicit zeros (233-236), transposed (252); kept rows:
    ``max / clip(min, default_kcount_value) >= ratio_threshold`` and ``max >= max_val`` (255-258), where ``/`` on two
    Python-2 numpy int arrays FLOORS and ``min`` over a sparse row sees the implicit zeros.
    PARITY UNPINNED: the function needs Python 2, pandas.SparseDataFrame and the external counters; nothing of it can be
    executed here, and the reference ships no test for it.  Returns (union size, kept k-mers sorted, counts[n][G])."""
    dicts = []
    for records in group_records:
        counts = py_oracle.count_kmers(records, k)
        dicts.append({km: c for km, c in counts.items() if c >= min_count})
    kmers = sorted(set().union(*[d.keys() for d in dicts]))
    kept, rows = [], []
    for km in kmers:
        row = [d.get(km, 0) for d in dicts]
        mx, mn = max(row), min(row)
        clipped = max(mn, default_kcount_value)
        if clipped > 0 and (mx // clipped) >= ratio_threshold and mx >= max_val:
            kept.append(km); rows.append(row)
    return len(kmers), kept, rows
