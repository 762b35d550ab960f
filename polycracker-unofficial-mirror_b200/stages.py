"""File-based stage functions: same names, arguments, file names and formats as the reference's click sub-commands
(polycracker/polycracker.py:113-367), so polycracker.nf and config_polyCRACKER.txt drive them unchanged.  Every
inter-stage hand-off is a file in the working directory, exactly like the reference; compute runs on the GPU through
the C ABI (binding.Context).  Where the reference shells out to an external tool the replacement is named in the
docstring.

Fast path between separately invoked stages: `blast_kmers` also drops a binary side-car (`<sam>.pcb200.npz`, the CSR the
GPU produced).  `blast2bed` and `generate_Kmer_Matrix` use it when it sits next to their input and fall back to parsing
the text artefacts otherwise (e.g. after Nextflow's collectFile concatenated several SAM files, polycracker.nf:286-290).
"""
import os
import re
import pickle

import numpy as np
import scipy.sparse as sps

from . import binding
from .binding import PolycrackerError
from .pipeline import ChunkLayout

_CTX = None


def context():
    """One lazily created GPU context per process (every reference stage is its own process anyway)."""
    global _CTX
    if _CTX is None:
        _CTX = binding.Context(int(os.environ.get("POLYCRACKER_B200_DEVICE", "0")))
    return _CTX


def create_path(path):
    os.makedirs(path, exist_ok=True)


def _split_name(fasta_path, fasta_file):
    return (fasta_path + fasta_file).split(".fa")[0] + "_split.fa"           # polycracker.py:170-171


def _write_records(path, names, seq, begins, ends):
    """`bedtools getfasta -name`: header = the bed name column, sequence on one line, case preserved."""
    with open(path, "wb") as f:
        for n, b, e in zip(names, begins.tolist(), ends.tolist()):
            f.write(b">" + n.encode() + b"\n")
            f.write(seq[b:e].tobytes())
            f.write(b"\n")


def _chunk_len(name):
    p = name.split("_")
    return int(p[-1]) - int(p[-2])


# --------------------------------------------------------------------------------------------------- a1
def splitFasta(fasta_file, fasta_path="./fasta_files/", chunk_size=75000, prefilter=0, min_chunk_size=0,
               remove_non_chunk=0, min_chunk_threshold=0, low_memory=0):
    """polycracker.py:113-171.  pyfaidx + `bedtools sort` + `bedtools getfasta -name` are replaced by host code;
    outputs: correspondence.bed, correspondenceOriginal.bed, <genome>_split.fa (+ the filtered-out files)."""
    if not os.path.exists(fasta_path):
        raise OSError("Specified fasta_file directory does not exist")
    seq, offsets, names = binding.read_fasta(fasta_path + fasta_file)
    with open(fasta_path + fasta_file + ".fai", "w") as f:                   # Fasta(...) side effect, pc.py:127
        for n, a, b in zip(names, offsets[:-1].tolist(), offsets[1:].tolist()):
            f.write("%s\t%d\n" % (n, b - a))
    lay = ChunkLayout(names, offsets, chunk_size)
    pos = {n: i for i, n in enumerate(lay.chunk_names)}
    rows = lay.bed_rows()

    def save(path, rws):
        with open(path, "w") as f:
            f.write("".join("%s\t%d\t%d\t%s\n" % r for r in rws))

    save("correspondence.bed", rows)
    save("correspondenceOriginal.bed", rows)
    kept = rows
    if prefilter and low_memory:
        if remove_non_chunk:
            kept = [r for r in rows if r[2] - r[1] == chunk_size]
            dropped = [r for r in rows if r[2] - r[1] != chunk_size]
        elif min_chunk_threshold:
            kept = [r for r in rows if r[2] - r[1] >= min_chunk_size]
            dropped = [r for r in rows if r[2] - r[1] < min_chunk_size]
        else:
            dropped = None
        if dropped is not None:
            save("correspondence.bed", kept)
            save("correspondence_filteredOut.bed", dropped)
            idx = np.array([pos[r[3]] for r in dropped], np.int64)
            _write_records(fasta_path + "filteredOut.fa", [r[3] for r in dropped], seq,
                           lay.chunk_begin_abs[idx], lay.chunk_end_abs[idx])
    idx = np.array([pos[r[3]] for r in kept], np.int64)
    out = _split_name(fasta_path, fasta_file)
    _write_records(out, [r[3] for r in kept], seq, lay.chunk_begin_abs[idx] if len(idx) else np.zeros(0, np.int64),
                   lay.chunk_end_abs[idx] if len(idx) else np.zeros(0, np.int64))
    return out


# --------------------------------------------------------------------------------------------------- a2
def writeKmerCount(fasta_path="./fasta_files/", kmercount_path="./kmercount_files", kmer_length="23,31", blast_mem="100"):
    """polycracker.py:174-209.  kmercountexact.sh (k <= 31, mincount=3) and jellyfish count -C / dump (k > 31) are
    replaced by K1-K3 on the GPU; each record of the split FASTA is an independent sequence.  k > 32 is refused."""
    kmer_lengths = [int(x) for x in str(kmer_length).split(",")]
    for k in kmer_lengths:
        if k < 1 or k > 32:
            raise PolycrackerError("kmer_length %d unsupported: the B200 path handles 1 <= k <= 32 (uint64 k-mers); "
                                   "run the reference's jellyfish path for longer k-mers" % k)
    create_path(kmercount_path)
    written = []
    for fasta_file in sorted(os.listdir(fasta_path)):
        if (fasta_file.endswith(".fa") or fasta_file.endswith(".fasta")) and "_split" in fasta_file:
            f = fasta_file.rstrip()
            print(f)
            out = kmercount_path + "/" + f[:f.rfind(".")] + ".kcount"
            open(out, "w").close()
            seq, offsets, _ = binding.read_fasta(os.path.join(fasta_path, fasta_file))
            ctx = context()
            ctx.load_sequences(seq, offsets[:-1], offsets[1:])
            for k in kmer_lengths:
                floor = 3 if k <= 31 else 1                                  # mincount=3 (pc.py:199) / dump all (pc.py:203)
                ctx.tune(keep_min=floor)                                     # rarer k-mers never leave the SM
                try:
                    ctx.count(k)
                finally:
                    ctx.tune(keep_min=1)
                keys, counts = ctx.dump_counts(floor)
                order = np.argsort(keys)
                binding.write_kcount(out, keys[order], counts[order], k, append=True)
            written.append(out)
    return written


# --------------------------------------------------------------------------------------------------- a3
def kmer2Fasta(kmercount_path="./kmercount_files/", kmer_low_count=100, high_bool=0, kmer_high_count=2000000,
               sampling_sensitivity=1):
    """polycracker.py:212-238, line for line: a text filter over every .kcount (>= low, <= high when high_bool, every
    sampling_sensitivity-th survivor in file order).  Host side by nature (text in, text out)."""
    if sampling_sensitivity < 1:
        sampling_sensitivity = 1
    count = 0
    for kmer in os.listdir(kmercount_path):
        if kmer.endswith(".kcount"):
            with open(kmercount_path + kmer, "r") as f, open(kmercount_path + kmer + ".fa", "w") as f2:
                for line in f:
                    fields = line.split()
                    if not fields:
                        continue
                    c = int(fields[-1])
                    if c >= kmer_low_count and (not high_bool or c <= kmer_high_count):
                        if count % sampling_sensitivity == 0:
                            f2.write(">%s\n%s\n" % (fields[0], fields[0]))
                        count += 1


# --------------------------------------------------------------------------------------------------- a4
def _read_kmer_fasta(path):
    with open(path) as f:
        return [line.strip() for line in f.readlines()[1::2] if line.strip()]


def blast_kmers(blast_mem="48", reference_genome=None, kmer_fasta=None, output_file="results.sam", kmer_length=13,
                perfect_mode=1, threads=5):
    """polycracker.py:241-252.  `bbmap.sh perfectmode=t ambiguous=all secondary=t maxsites=2e9` is replaced by K5/K6:
    every exact occurrence, either strand, of every query k-mer in every record of the reference genome.
    perfect_mode=0 (bbmap's heuristic mapping) cannot be reproduced and is refused.  kmer_length (bbmap's seed
    length), blast_mem and threads are accepted and ignored."""
    if not perfect_mode:
        raise PolycrackerError("perfect_mode=0 (bbmap heuristic mapping) is not reproducible on the GPU path: set "
                               "perfect_mode = 1 in config_polyCRACKER.txt (SURVEY.md F4)")
    kmers = _read_kmer_fasta(kmer_fasta)
    seq, offsets, names = binding.read_fasta(reference_genome)
    names = [n.split("::")[0] for n in names]
    row_len = np.array([_chunk_len(n) if n.count("_") >= 2 and n.split("_")[-1].isdigit() and n.split("_")[-2].isdigit()
                        else b - a for n, a, b in zip(names, offsets[:-1].tolist(), offsets[1:].tolist())], np.int64)
    ctx = context()
    ctx.load_sequences(seq, offsets[:-1], offsets[1:])
    by_k = {}
    for km in kmers:
        by_k.setdefault(len(km), []).append(km)
    open(output_file, "w").close()
    side = []
    for k in sorted(by_k):
        if k > 32:
            raise PolycrackerError("query k-mers of length %d unsupported (k <= 32)" % k)
        n_sel = ctx.set_selected(binding.kmer_encode(by_k[k], k), k)
        keys = ctx.selected(n_sel)
        nnz = ctx.build_matrix(np.arange(len(names), dtype=np.int32), len(names))
        indptr, indices, data = ctx.csr(len(names), nnz)
        binding.write_hit_lines(output_file, "sam", indptr, indices, data, keys, k, names, row_len, append=True)
        side.append((k, keys, indptr, indices, data))
    np.savez(output_file + ".pcb200.npz", n_groups=len(side), names=np.array(names), row_len=row_len,
             **{"%s_%d" % (nm, i): arr for i, g in enumerate(side)
                for nm, arr in zip(("k", "keys", "indptr", "indices", "data"), (np.array(g[0]),) + g[1:])})
    return output_file


def _load_sidecar(path):
    side = path + ".pcb200.npz"
    if not os.path.exists(side) or os.path.getmtime(side) + 1 < os.path.getmtime(path):
        return None
    z = np.load(side)
    groups = [tuple(z["%s_%d" % (nm, i)] for nm in ("k", "keys", "indptr", "indices", "data")) for i in range(int(z["n_groups"]))]
    return z["names"].tolist(), z["row_len"], groups


# --------------------------------------------------------------------------------------------------- a5
def blast2bed(blast_file="./blast_files/results.sam", bb=1, low_memory=0, output_bed_file="blasted.bed",
              output_merged_bed_file="blasted_merged.bed", external_call=False):
    """polycracker.py:255-281: SAM -> blasted.bed -> (bedtools sort | merge -c 4 -o collapse) blasted_merged.bed."""
    if external_call:
        with open(blast_file) as f, open(output_bed_file, "w") as f2:      # the awk one-liner, pc.py:265
            for line in f:
                p = line.split()
                if len(p) >= 3:
                    f2.write("%s\t0\t100\t%s\n" % (p[2], p[0]))
        side = None
    else:
        side = _load_sidecar(blast_file) if bb else None
        if side is not None:
            names, row_len, groups = side
            open(output_bed_file, "w").close()
            for k, keys, indptr, indices, data in groups:
                binding.write_hit_lines(output_bed_file, "bed", indptr, indices, data, keys, int(k), names, row_len, append=True)
        else:
            col = 2 if bb else 1
            with open(blast_file, "r") as f, open(output_bed_file, "w") as f2:
                for line in f:
                    if line.strip():
                        p = line.split("\t")
                        l1 = p[col].split("::")[0]
                        f2.write("\t".join([l1, "0", str(_chunk_len(l1)), p[0]]) + "\n")
    if low_memory == 0:
        if side is not None and len(side[2]) == 1:
            names, row_len, groups = side
            k, keys, indptr, indices, data = groups[0]
            order = sorted(range(len(names)), key=lambda i: names[i].encode())       # bedtools sort: chrom bytewise
            csr = sps.csr_matrix((data, indices, indptr), shape=(len(names), len(keys)))[order]
            binding.write_hit_lines(output_merged_bed_file, "merged", csr.indptr, csr.indices, csr.data, keys, int(k),
                                    [names[i] for i in order], row_len[order])
            # binary twin of the merged bed for generate_Kmer_Matrix (at Gbp scale the text is hundreds of GB of k-mer
            # names; the stage that reads it back only needs (chunk, k-mer, multiplicity))
            np.savez(output_merged_bed_file + ".pcb200.npz", k=np.array(int(k)), keys=keys, names=np.array([names[i] for i in order]),
                     indptr=csr.indptr.astype(np.int64), indices=csr.indices.astype(np.int32), data=csr.data.astype(np.float32))
        else:
            per_chunk = {}
            with open(output_bed_file) as f:
                for line in f:
                    p = line.rstrip("\n").split("\t")
                    if len(p) >= 4:
                        per_chunk.setdefault((p[0], p[1], p[2]), []).append(p[3])
            with open(output_merged_bed_file, "w") as f:
                for key in sorted(per_chunk, key=lambda t: (t[0].encode(), int(t[1]))):
                    f.write("\t".join(key) + "\t" + ",".join(per_chunk[key]) + "\n")


# --------------------------------------------------------------------------------------------------- a6 / a7
def findScaffolds():
    """polycracker.py:284-287."""
    with open("correspondence.bed", "r") as f:
        return sorted([line.split("\t")[-1].strip("\n") for line in f.readlines() if line])


def findKmerNames(kmercount_path, genome):
    """polycracker.py:290-296 (including its 'first matching file in os.listdir order' behaviour)."""
    for file in os.listdir(kmercount_path):
        if file.startswith(genome[:genome.find(".fa")]) and (file.endswith(".fa") or file.endswith(".fasta")):
            print(file)
            with open(kmercount_path + file, "r") as f:
                return sorted([line.strip("\n") for line in f.readlines()[1::2] if line])


def _cols_of(names_list, kmers, kmer_idx):
    """column ids of a list of k-mer names (vectorised through the uint64 codec when all share one length)."""
    if not names_list:
        return np.zeros(0, np.int32)
    lens = {len(n) for n in names_list}
    klens = {len(n) for n in kmers}
    if len(lens) == 1 and lens == klens and next(iter(lens)) <= 32:
        k = next(iter(lens))
        keys = binding.kmer_encode(kmers, k)
        q = binding.kmer_encode(names_list, k)
        order = np.argsort(keys)
        pos = np.searchsorted(keys[order], q)
        pos[pos >= len(keys)] = 0
        hit = (keys[order][pos] == q) & (q != np.uint64(0xFFFFFFFFFFFFFFFF))   # ~0 = "not a k-mer" (N, other letters)
        return np.where(hit, order[pos], -1).astype(np.int32)
    return np.array([kmer_idx.get(n, -1) for n in names_list], np.int32)


def generate_Kmer_Matrix(kmercount_path="./kmercount_files/", genome=None, chunk_size=75000, min_chunk_size=0,
                         remove_non_chunk=1, min_chunk_threshold=0, low_memory=0, prefilter=0,
                         fasta_path="./fasta_files/", genome_split_name=None, original_genome=None,
                         input_bed_file="blasted.bed", input_merged_bed_file="blasted_merged.bed",
                         output_sparse_matrix="clusteringMatrix.npz", output_kmer_file="kmers.p", tfidf=0):
    """polycracker.py:299-367.  The Counter -> dok_matrix -> tocsc work runs on the GPU (K6, pc_matrix_from_hits); the
    artefacts are the reference's: clusteringMatrix.npz (scipy CSC float32), rowNames.txt, colNames.txt, scaffolds.p,
    originalScaffolds.p, kmers.p (pickle protocol 2).  With tfidf=1 the GPU TF-IDF (K7) is written next to it as
    clusteringMatrix.tfidf.npz; clusteringMatrix.npz itself keeps the raw counts because transform_plot applies
    TfidfTransformer to it (polycracker.py:384-396)."""
    if prefilter and low_memory:
        seq, offsets, names = binding.read_fasta(fasta_path + original_genome)
        lay = ChunkLayout(names, offsets, chunk_size)
        pos = {n: i for i, n in enumerate(lay.chunk_names)}
        rows = [l.rstrip("\n").split("\t") for l in open("correspondenceOriginal.bed") if l.strip()]
        idx = np.array([pos[r[3]] for r in rows], np.int64)
        _write_records(fasta_path + genome_split_name, [r[3] for r in rows], seq, lay.chunk_begin_abs[idx], lay.chunk_end_abs[idx])
    kmers = findKmerNames(kmercount_path, genome) or []
    original_scaffolds = findScaffolds()
    if remove_non_chunk and prefilter == 0:
        scaffolds = [s for s in original_scaffolds if abs(_chunk_len(s)) == chunk_size]
    elif min_chunk_threshold and prefilter == 0:
        scaffolds = [s for s in original_scaffolds if abs(_chunk_len(s)) >= min_chunk_size]
    else:
        scaffolds = original_scaffolds
    scaffold_idx = {s: i for i, s in enumerate(scaffolds)}
    kmer_idx = {km: i for i, km in enumerate(kmers)}
    per_row = [[] for _ in scaffolds]
    side = None if low_memory else _load_merged_sidecar(input_merged_bed_file, kmers)
    if side is not None:
        # blast2bed left the merged bed's content in binary form: no text parsing, no Python lists of names
        col_of_key, names_s, indptr_s, indices_s, data_s = side
        row_of_name = {n: i for i, n in enumerate(names_s)}
        src = np.array([row_of_name.get(sc, -1) for sc in scaffolds], np.int64)
        csr_s = sps.csr_matrix((data_s, indices_s, indptr_s), shape=(len(names_s), len(col_of_key)))
        pick = csr_s[np.where(src >= 0, src, 0)]
        lens_r = np.diff(pick.indptr) * (src >= 0)
        mult = pick.data.astype(np.int64)
        valid = np.repeat(src >= 0, np.diff(pick.indptr)) & (col_of_key[pick.indices] >= 0)
        cols = np.repeat(col_of_key[pick.indices][valid], mult[valid]).astype(np.int32)
        row_id = np.repeat(np.repeat(np.arange(len(scaffolds)), np.diff(pick.indptr))[valid], mult[valid])
        del lens_r
        row_offsets = np.zeros(len(scaffolds) + 1, np.int64)
        np.cumsum(np.bincount(row_id, minlength=len(scaffolds)), out=row_offsets[1:])
        order = np.argsort(row_id, kind="stable")
        return _finish_matrix(scaffolds, original_scaffolds, kmers, row_offsets, cols[order], output_sparse_matrix,
                              output_kmer_file, tfidf)
    if low_memory:
        with open(input_bed_file) as f:
            for line in f:
                p = line.strip("\n").split("\t")
                if len(p) >= 4 and p[0] in scaffold_idx:
                    per_row[scaffold_idx[p[0]]].append(p[-1])
    else:
        with open(input_merged_bed_file) as f:
            for line in f:
                p = line.rstrip("\n").split("\t")
                if len(p) >= 4 and p[0] in scaffold_idx:
                    per_row[scaffold_idx[p[0]]] = p[-1].split(",")      # a later line for the same chunk overwrites
    flat = [n for row in per_row for n in row]
    cols = _cols_of(flat, kmers, kmer_idx)
    counts = np.array([len(r) for r in per_row], np.int64)
    row_of = np.repeat(np.arange(len(per_row)), counts)
    keep = cols >= 0                                                     # unknown names are skipped (pc.py:350-353)
    cols, row_of = cols[keep], row_of[keep]
    row_offsets = np.zeros(len(scaffolds) + 1, np.int64)
    np.cumsum(np.bincount(row_of, minlength=len(scaffolds)), out=row_offsets[1:])
    return _finish_matrix(scaffolds, original_scaffolds, kmers, row_offsets, cols, output_sparse_matrix, output_kmer_file, tfidf)


def _load_merged_sidecar(merged_bed, kmers):
    """(column id of every side-car k-mer or -1, row names, CSR) when blast2bed left a binary twin next to the merged bed that
    is at least as new as the text, and every column label has the side-car's k; else None (the text is parsed)."""
    side = merged_bed + ".pcb200.npz"
    if not (os.path.exists(side) and os.path.exists(merged_bed)) or os.path.getmtime(side) + 1 < os.path.getmtime(merged_bed):
        return None
    z = np.load(side)
    k = int(z["k"])
    if not kmers or {len(x) for x in kmers} != {k}:
        return None
    col_keys = binding.kmer_encode(kmers, k)                  # column j = kmers[j] (already string-sorted = key-sorted)
    order = np.argsort(col_keys)
    pos = np.searchsorted(col_keys[order], z["keys"])
    pos[pos >= len(col_keys)] = 0
    hit = col_keys[order][pos] == z["keys"]
    return np.where(hit, order[pos], -1).astype(np.int64), z["names"].tolist(), z["indptr"], z["indices"], z["data"]


def _finish_matrix(scaffolds, original_scaffolds, kmers, row_offsets, cols, output_sparse_matrix, output_kmer_file, tfidf):
    """K6 (+ K7) on the GPU from per-row hit lists, then the reference's artefacts (polycracker.py:355-367)."""
    ctx = context()
    nnz = ctx.matrix_from_hits(row_offsets, cols, len(kmers))
    indptr, indices, data = ctx.csr(len(scaffolds), nnz)
    csr = sps.csr_matrix((data, indices.astype(np.int32), indptr.astype(np.int32 if nnz < 2**31 else np.int64)),
                         shape=(len(scaffolds), len(kmers)))
    if tfidf and nnz:
        ctx.tfidf(len(scaffolds))
        tf = sps.csr_matrix((ctx.tfidf_values(nnz), csr.indices, csr.indptr), shape=csr.shape)
        sps.save_npz(os.path.splitext(output_sparse_matrix)[0] + ".tfidf.npz", tf)
    mat = csr.tocsc()
    if not tfidf and len(scaffolds):
        lens = np.array([[float(x) for x in s.split("_")[-2:]] for s in scaffolds])
        x = np.abs(lens[:, 1] - lens[:, 0]) / 5000.
        coo = mat.tocoo()
        vals = (coo.data.astype(np.float64) * (1.0 / x[coo.row])).astype(np.float32)      # pc.py:357-359
        mat = sps.csc_matrix((vals, (coo.row, coo.col)), shape=mat.shape)
    sps.save_npz(output_sparse_matrix, mat)
    with open("rowNames.txt", "w") as f:
        f.write("\n".join("\t".join([str(i), scaffolds[i]]) for i in range(len(scaffolds))))
    with open("colNames.txt", "w") as f:
        f.write("\n".join("\t".join([str(i), kmers[i]]) for i in range(len(kmers))))
    pickle.dump(scaffolds, open("scaffolds.p", "wb"), protocol=2)
    pickle.dump(original_scaffolds, open("originalScaffolds.p", "wb"), protocol=2)
    pickle.dump(kmers, open(output_kmer_file, "wb"), protocol=2)
    return mat


# --------------------------------------------------------------------------------------------------- N4 external counts
def hipmer_output_to_kcount(hipmer_input="test.txt", kcount_output="test.final.kcount", run_on_dir=False):
    """polycracker/format.py:123-133: `cat files | awk '{OFS="\t"; sum=0; for (i=2;i<=7;i++) sum+=$i; if (sum>=3) print $1, sum}'`.
    HipMer's k-mer count files carry the k-mer and six count columns; the .kcount line is the k-mer and their sum, kept
    when the sum is at least 3.  awk semantics: fields split on runs of blanks, a missing or non-numeric field counts 0
    (its leading number, if any), sums print as integers when integral."""
    files = [os.path.join(hipmer_input, f) for f in os.listdir(hipmer_input)] if run_on_dir else [hipmer_input]

    def num(tok):
        m = re.match(r"[ \t]*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", tok)
        return float(m.group(0)) if m else 0.0

    with open(kcount_output, "w") as out:
        for path in files:
            with open(path) as f:
                for line in f:
                    p = line.split()
                    total = sum(num(x) for x in p[1:7])
                    if total >= 3:
                        out.write("%s\t%s\n" % (p[0] if p else "", ("%d" % total) if total == int(total) else ("%.6g" % total)))
    return kcount_output


def read_kcount(path):
    """A .kcount text file (`KMER<blank or tab>COUNT` per line, polycracker.py:199-208) -> (k, uint64 canonical keys,
    uint32 counts).  Lines whose k-mer has another length than the first line's, or letters other than ACGT, are skipped
    like the reference's downstream mapping would never place them; both strands of one k-mer are merged (counts add)."""
    names, counts = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) >= 2:
                names.append(p[0]); counts.append(int(float(p[-1])))
    if not names:
        return 0, np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    k = len(names[0])
    keep = [i for i, n in enumerate(names) if len(n) == k]
    keys = binding.kmer_encode([names[i] for i in keep], k)
    cnt = np.asarray(counts, np.int64)[keep]
    ok = keys != np.uint64(0xFFFFFFFFFFFFFFFF)
    keys, cnt = keys[ok], cnt[ok]
    canon = np.array([binding.host_canonical(int(x), k) for x in keys.tolist()], np.uint64) if len(keys) < 4096 else _canonical_np(keys, k)
    uniq, inv = np.unique(canon, return_inverse=True)
    total = np.bincount(inv, weights=cnt.astype(np.float64), minlength=len(uniq)).astype(np.int64)
    return k, uniq, np.minimum(total, 0xFFFFFFFF).astype(np.uint32)


def _canonical_np(keys, k):
    """strand collapse of uint64 k-mers in numpy (min of the k-mer and its reverse complement)"""
    x = keys.astype(np.uint64)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF), (16, 0x0000FFFF0000FFFF)):
        x = ((x >> np.uint64(sh)) & np.uint64(m)) | ((x & np.uint64(m)) << np.uint64(sh))
    x = (x >> np.uint64(32)) | (x << np.uint64(32))
    rc = (~x) >> np.uint64(64 - 2 * k)
    return np.minimum(keys, rc)


def load_kcount(ctx, path):
    """Install the k-mers of a .kcount file (HipMer via hipmer_output_to_kcount, jellyfish dump, kmercountexact) as the
    context's count result: select / build_matrix / tfidf then run on externally counted k-mers (pc_load_counts)."""
    k, keys, counts = read_kcount(path)
    if k == 0:
        raise PolycrackerError("%s holds no k-mers" % path)
    ctx.load_counts(keys, counts, k)
    return k, len(keys)


# --------------------------------------------------------------------------------------------------- N2 diagnostics
def repeatmers_per_subsequence(merged_bed="blasted_merged.bed"):
    """The data behind `number_repeatmers_per_subsequence` (polycracker.py:1374: awk '{print gsub(/,/,"")+1}'): for every
    line of the merged bed the number of comma separated k-mer names in it (awk counts commas on the whole line)."""
    out = []
    with open(merged_bed) as f:
        for line in f:
            out.append(line.count(",") + 1)
    return np.asarray(out, np.int64)


def number_repeatmers_per_subsequence(merged_bed="blasted_merged.bed", out_file="kmers_per_subsequence.png", kde=False):
    """polycracker.py:1366-1382.  The per-chunk numbers are what `Context.row_hits` returns for the in-memory path; here
    they come from the text artefact like in the reference.  Writes `<out_file>.tsv` (value, frequency) always and the
    PNG when matplotlib is importable (the histogram is a plot of exactly that table)."""
    kcount = repeatmers_per_subsequence(merged_bed)
    if not out_file.endswith(".png"):
        out_file += ".png"                                   # pc.py:1380-1381
    values, freq = np.unique(kcount, return_counts=True)
    with open(out_file + ".tsv", "w") as f:
        f.write("\n".join("%d\t%d" % (v, n) for v, n in zip(values.tolist(), freq.tolist())))
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure()
        plt.hist(kcount, bins=min(100, max(1, len(values))), density=bool(kde))
        plt.title("Histogram of Repeat-mers in Chunked Genome Fragment")
        plt.xlabel("Number of Repeat-mers in Chunked Genome Fragment")
        plt.ylabel("Frequency/Density")
        plt.savefig(out_file, dpi=300)
    except ImportError:
        pass
    return kcount


def kcount_histogram(kcount_file):
    """The data behind `kcount_hist` (polycracker.py:1466-1469): OrderedDict(sorted(Counter(second column).items()))
    as two arrays (count value, number of k-mers).  `Context.count_histogram` is the same table straight from the GPU
    count, without the text file."""
    vals = []
    with open(kcount_file) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2 and parts[1]:
                vals.append(int(parts[1]))
    values, freq = np.unique(np.asarray(vals, np.int64), return_counts=True)
    return values, freq.astype(np.int64)


def kcount_hist(kcount_directory, out_file="kmer_histogram.png", log=False, kcount_range="3,10000", low_count=0.,
                kmer_lengths="", sample_speed=1):
    """polycracker.py:1419-1496: one count histogram per `<k>.kcount` in the directory.  Writes `<out_file>.tsv` (file,
    count, n_kmers) and, when matplotlib and scipy's smoothing are available, the PNG."""
    wanted = None
    if kmer_lengths:
        wanted = set(int(x) for x in kmer_lengths.split(":")[0].split(","))
    rows, tables = [], {}
    for name in sorted(os.listdir(kcount_directory), reverse=True):            # pc.py:1450
        if name.endswith(".kcount") and (wanted is None or int(name.split(".")[0]) in wanted):
            values, freq = kcount_histogram(os.path.join(kcount_directory, name))
            tables[name] = (values, freq)
            rows += ["%s\t%d\t%d" % (name, v, n) for v, n in zip(values.tolist(), freq.tolist())]
    with open(out_file + ".tsv", "w") as f:
        f.write("\n".join(rows))
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        lo, hi = (int(x) for x in kcount_range.split(","))
        plt.figure(figsize=(7, 7))
        for name, (values, freq) in tables.items():
            x = np.log(values) if log else values
            m = (values >= lo) & (values <= hi)
            plt.plot(x[m][::abs(int(sample_speed)) or 1], freq[m][::abs(int(sample_speed)) or 1],
                     label="kmer length = %s" % name.split(".")[0])
        if low_count:
            plt.axvline(x=low_count, label="Kmer-Count Cutoff Value = %d" % low_count)
        plt.title("KmerCount Histogram")
        plt.xlabel("Total Kmer Counts in Genome in log(b10)x" if log else "Total Kmer Counts in Genome")
        plt.ylabel("Frequency")
        plt.legend()
        plt.savefig(out_file, dpi=300)
    except ImportError:
        pass
    return tables
