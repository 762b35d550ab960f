"""Scope row N1: the inner loop of subgenome extraction (polycracker/polycracker.py:676-1020) on the GPU.

In the reference every bootstrap iteration shells out to kmercountexact.sh once per subgenome and k-mer length
(writeKmerCountSubgenome, 692-719), joins the dumps in Python dictionaries (compareKmers, 735-783), maps the
differential k-mers back with bbmap.sh (writeBlast, 786-799), and pushes the SAM through bedtools coverage /
unionbedg (blast2bed3 802-865, bed2unionBed 868-879) before binning the chunks (kmerratio2scaffasta, 884-966).
Here the genome stays packed in HBM: each subgenome's chunks are re-counted with K1-K3, the dictionaries are joined
by one kernel per subgenome, and one pass over the packed genome yields the per-chunk coverage of every subgenome.

`extract_pass` is the in-memory API; `subgenomeExtraction` mirrors the click command (same arguments, same files).
"""
import os

import numpy as np

from . import binding
from .binding import PolycrackerError


def dump_floor(k):
    """kmercountexact.sh mincount=3 for k <= 31, jellyfish dump (everything) above (polycracker.py:711-714)."""
    return 3 if k <= 31 else 1


def parse_kmer_lengths(kmer_length):
    ks = [int(x) for x in str(kmer_length).split(",")]
    for k in ks:
        if k < 1 or k > 32:
            raise PolycrackerError("kmer_length %d unsupported: the B200 path handles 1 <= k <= 32 (uint64 k-mers)" % k)
    if len(set(ks)) != len(ks):
        raise PolycrackerError("kmer_length lists a length twice: %s" % kmer_length)
    return ks


def extract_pass(ctx, seq, rec_begin, rec_end, groups, kmer_lengths, diff_kmer_threshold=20,
                 default_kmercount_value=3., diff_sample_rate=1, map_seq=None, map_begin=None, map_end=None,
                 want_diff=True):
    """One iteration: count every subgenome, find the differential k-mers, map them back.

    seq: ASCII bases (numpy uint8 or a torch CUDA uint8 tensor -- resident input avoids one H2D per subgenome);
    rec_begin/rec_end: int64 offsets of every record (chunk) of the counted genome; groups: one integer index array per
    subgenome (records of that subgenome's model FASTA); map_*: the genome the k-mers are mapped back to (default: the
    counted one; the reference's `original_genome`).
    Returns (coverage int32[n_map_records, n_groups], diff) with diff[k][g] = (keys, counts[n, n_groups]).
    """
    rec_begin = np.ascontiguousarray(rec_begin, np.int64)
    rec_end = np.ascontiguousarray(rec_end, np.int64)
    n_groups = len(groups)
    ctx.sub_begin(n_groups)
    n_diff = {}
    for k in kmer_lengths:
        for g, idx in enumerate(groups):
            idx = np.sort(np.asarray(idx, np.int64))           # counting is order-free; the loader wants ascending offsets
            ctx.load_sequences(seq, rec_begin[idx], rec_end[idx])
            ctx.tune(keep_min=dump_floor(k))                    # the dictionary only ever holds what the counter dumps
            try:
                ctx.count(k)
            finally:
                ctx.tune(keep_min=1)
            ctx.sub_add_counts(g, dump_floor(k))
        n_diff[k] = ctx.sub_compare(diff_kmer_threshold, default_kmercount_value, diff_sample_rate)
    if map_seq is None:
        map_seq, map_begin, map_end = seq, rec_begin, rec_end
    ctx.load_sequences(map_seq, np.ascontiguousarray(map_begin, np.int64), np.ascontiguousarray(map_end, np.int64))
    coverage = ctx.sub_mapback(len(map_begin))
    diff = {}
    if want_diff:
        for k in kmer_lengths:
            diff[k] = [ctx.sub_diff(k, g) for g in range(n_groups)]
    return coverage, diff, n_diff


def kmer_genome_matrix(ctx, seq, rec_begin, rec_end, groups, k=23, min_count=3, ratio_threshold=20,
                       default_kcount_value=1, max_val=100):
    """Scope row N4, the data path of `bio_hyp_class` (utilities.py:177-284): count every genome (group of records) with the
    counter's ``mincount=min_count`` floor, form the k-mer x genome count matrix over the union of the dumped k-mers and keep
    the k-mers with max / clip(min, default) >= ratio_threshold (floored, Python 2) and max >= max_val.
    Returns (union size, keys uint64[n] ascending = kmers_union.p order, counts uint32[n, n_genomes] =
    union_kmer_matrix.npz)."""
    if len(groups) > binding.PC_MAX_GROUPS:
        raise binding.PolycrackerError("%d genomes: one session holds at most %d" % (len(groups), binding.PC_MAX_GROUPS))
    rec_begin = np.ascontiguousarray(rec_begin, np.int64)
    rec_end = np.ascontiguousarray(rec_end, np.int64)
    ctx.sub_begin(len(groups))
    for g, idx in enumerate(groups):
        idx = np.sort(np.asarray(idx, np.int64))
        ctx.load_sequences(seq, rec_begin[idx], rec_end[idx])
        ctx.tune(keep_min=max(1, int(min_count)))
        try:
            ctx.count(k)
        finally:
            ctx.tune(keep_min=1)
        ctx.sub_add_counts(g, max(1, int(min_count)))
    return ctx.sub_union(ratio_threshold, default_kcount_value, max_val)


def bin_records(coverage, absolute_threshold=10, ratio_threshold=2):
    """kmerratio2scaffasta's rule (polycracker.py:908-922), vectorised: record i goes to subgenome g when for EVERY
    other subgenome o either (x_o == 0 and x_g > absolute_threshold) or (x_o > 0 and x_g / x_o > ratio_threshold).
    Returns int array: the subgenome of each record, -1 = ambiguous.  (At most one subgenome can qualify when
    ratio_threshold >= 1; otherwise the record is listed under each, as in the reference -- see `bin_lists`.)"""
    member = bin_membership(coverage, absolute_threshold, ratio_threshold)
    out = np.full(len(coverage), -1, np.int64)
    for g in range(member.shape[1]):
        out[member[:, g] & (out < 0)] = g
    return out


def bin_membership(coverage, absolute_threshold=10, ratio_threshold=2):
    x = np.asarray(coverage, np.float64)
    n, n_groups = x.shape
    member = np.ones((n, n_groups), bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for g in range(n_groups):
            for o in range(n_groups):
                if o == g:
                    continue
                ok = np.where(x[:, o] == 0, x[:, g] > absolute_threshold,
                              (x[:, o] > 0) & (x[:, g] / np.where(x[:, o] == 0, 1.0, x[:, o]) > ratio_threshold))
                member[:, g] &= ok
    return member


# ------------------------------------------------------------------------------------------------ file-based mirror
def fai2bed(genome):
    """polycracker.py:676-689: <genome>.bed with one whole-record window per sequence; returns {name: [bed line]} and
    the parsed FASTA.  A missing <genome>.fai is written with the two columns the chain reads (name, length;
    pc.py:873); an existing one (pyfaidx / samtools) is left alone."""
    seq, offsets, names = binding.read_fasta(genome)
    out = {}
    if not os.path.exists(genome + ".fai"):
        with open(genome + ".fai", "w") as fai:
            for n, a, b in zip(names, offsets[:-1].tolist(), offsets[1:].tolist()):
                fai.write("%s\t%d\n" % (n, b - a))
    with open(genome + ".bed", "w") as bed:
        for n, a, b in zip(names, offsets[:-1].tolist(), offsets[1:].tolist()):
            line = "%s\t%s\t%d\t%s\n" % (n, "0", b - a, n)
            bed.write(line)
            out[n] = [line]
    return out, (seq, offsets, names)


def _write_fasta(path, names, seq, offsets, index):
    with open(path, "wb") as f:
        for n in names:
            i = index[n]
            f.write(b">" + n.encode() + b"\n" + seq[offsets[i]:offsets[i + 1]].tobytes() + b"\n")


def subgenomeExtraction(subgenome_folder, original_subgenome_path, fasta_path="./fasta_files/", genome_name=None,
                        original_genome=None, bb=1, bootstrap=0, iteration=0, kmer_length="23,31", run_final=0,
                        original=0, blast_mem="100", kmer_low_count=100, diff_kmer_threshold=20,
                        unionbed_threshold="10,2", diff_sample_rate=1, default_kmercount_value=3., search_length=13,
                        perfect_mode=1, ctx=None, write_fasta=True):
    """polycracker.py:970-1020 with the recursion through kmerratio2scaffasta (884-966) unrolled into a loop.

    Reads <subgenome_folder>/subgenome_*.txt (record names of `genome_name`), writes per iteration, under
    bootstrap_<i>/: model_subgenome_<g>.fa, kmercount_files/model_subgenome_<g>.higher.kmers.fa,
    model_subgenome_<g>.higher.kmers.sorted.cov.bedgraph, subgenomes.union.bedgraph,
    extractedSubgenomes/<genome>.subgenome<A..>.fasta + ambiguousScaffolds.fasta, the next iteration's
    bootstrap_<i+1>/subgenome_<g>.txt and finally <original_subgenome_path>/finalResults/subgenome_<g>.txt.
    Not written: the per-subgenome .kcount dumps, the SAM / bed3 / .cov intermediates (GB-sized text that only the
    next function of the chain reads) and reformat.sh's *_wrapped.fasta copies.
    bb=0 (blastn) and perfect_mode=0 (bbmap heuristics) are refused; search_length and blast_mem are ignored.
    Returns the per-iteration summaries."""
    if not bb or not perfect_mode:
        raise PolycrackerError("only bb=1 with perfect_mode=1 (exact matches) is reproducible on the GPU path")
    from .stages import context, create_path
    if ctx is None:
        ctx = context()
    kmer_lengths = parse_kmer_lengths(kmer_length)
    try:
        absolute_threshold, ratio_threshold = tuple(map(int, str(unionbed_threshold).split(",")))
    except Exception:
        absolute_threshold, ratio_threshold = 10, 2                # polycracker.py:888-891
    if not fasta_path.endswith("/"):
        fasta_path += "/"
    bed_dict, (seq, offsets, names) = fai2bed(fasta_path + genome_name)
    index = {n: i for i, n in enumerate(names)}
    if original_genome == genome_name:
        map_seq, map_off, map_names = seq, offsets, names
    else:
        _, (map_seq, map_off, map_names) = fai2bed(fasta_path + original_genome)
    # the records that get extracted: `genome_name`, or `original_genome` when original=1 (pc.py:892-894)
    if original:
        ex_seq, ex_off, ex_index = map_seq, map_off, {n: i for i, n in enumerate(map_names)}
        ex_prefix = original_genome
    else:
        ex_seq, ex_off, ex_index = seq, offsets, index
        ex_prefix = genome_name
    ex_prefix = ex_prefix[:ex_prefix.rfind(".")]
    try:
        import torch
        dev = torch.device("cuda", ctx.device)
        seq_in = torch.from_numpy(seq).to(dev)
        map_in = seq_in if map_seq is seq else torch.from_numpy(map_seq).to(dev)
    except ImportError:                                           # torch is only plumbing: host arrays work too
        seq_in, map_in = seq, map_seq
    summaries = []
    while True:
        if subgenome_folder.endswith("/"):
            subgenome_folder = subgenome_folder[:-1]
        if not (bootstrap >= iteration or run_final == 1):
            break
        create_path(subgenome_folder + "/kmercount_files/")
        # subgenome_*.txt -> model_subgenome_*.fa (pc.py:1003-1010); empty files are skipped like in the reference
        groups, labels = [], []
        for fn in sorted(os.listdir(subgenome_folder)):
            path = subgenome_folder + "/" + fn
            if fn.endswith(".txt") and os.stat(path).st_size:
                members = [l.strip("\n") for l in open(path) if l.strip("\n")]
                with open(path.replace(".txt", ".bed"), "w") as f:
                    for m in members:
                        f.write(bed_dict[m][0])
                if write_fasta:
                    _write_fasta(subgenome_folder + "/model_" + fn.replace(".txt", ".fa"), members, seq, offsets, index)
                groups.append(np.array([index[m] for m in members], np.int64))
                labels.append("model_" + fn[:-len(".txt")])
        if not groups:
            raise PolycrackerError("no non-empty subgenome_*.txt in %s" % subgenome_folder)
        coverage, diff, n_diff = extract_pass(ctx, seq_in, offsets[:-1], offsets[1:], groups, kmer_lengths,
                                              diff_kmer_threshold, default_kmercount_value, diff_sample_rate,
                                              map_in, map_off[:-1], map_off[1:])
        n_groups = len(groups)
        # compareKmers' output files (pc.py:767 / 781)
        sampled = max(1, int(diff_sample_rate)) > 1
        for g, label in enumerate(labels):
            with open("%s/kmercount_files/%s.higher.kmers.fa" % (subgenome_folder, label), "w") as f:
                for k in kmer_lengths:
                    keys, counts = diff[k][g]
                    kmers = binding.kmer_decode(keys, k)
                    others = [o for o in range(n_groups) if o != g]
                    for kmer, row in zip(kmers, counts.tolist()):
                        vals = [row[o] if row[o] else float(default_kmercount_value) for o in others]
                        txt = ".".join(str(v) if sampled else str(int(v)) for v in vals)
                        f.write(">%s.%d.%s\n%s\n" % (kmer, row[g], txt, kmer))
        # blast2bed3's bedgraphs (genome order) and bed2unionBed's union (bedtools sort order)
        lens = np.diff(map_off)
        for g, label in enumerate(labels):
            with open("%s/%s.higher.kmers.sorted.cov.bedgraph" % (subgenome_folder, label), "w") as f:
                for n, ln, c in zip(map_names, lens.tolist(), coverage[:, g].tolist()):
                    f.write("%s\t%d\t%d\t%d\n" % (n, 0, ln, c))
        order = sorted(range(len(map_names)), key=lambda i: map_names[i].encode())
        with open(subgenome_folder + "/subgenomes.union.bedgraph", "w") as f:
            for i in order:
                f.write("%s\t0\t%d\t%s\n" % (map_names[i], lens[i], "\t".join(str(int(v)) for v in coverage[i])))
        # kmerratio2scaffasta (pc.py:884-966)
        member = bin_membership(coverage, absolute_threshold, ratio_threshold)
        scaffolds_out = [[map_names[i] for i in order if member[i, g]] for g in range(n_groups)]
        ambiguous = [map_names[i] for i in order if not member[i].any()]
        no_kill = all(len(s) > 0 for s in scaffolds_out)
        extract_path = subgenome_folder + "/extractedSubgenomes/"
        create_path(extract_path)
        if write_fasta:
            for g, members in enumerate(scaffolds_out):
                _write_fasta(extract_path + ex_prefix + ".subgenome" + chr(g + 65) + ".fasta", members, ex_seq, ex_off, ex_index)
            _write_fasta(extract_path + "ambiguousScaffolds.fasta", ambiguous, ex_seq, ex_off, ex_index)
        summaries.append(dict(iteration=iteration, folder=subgenome_folder, n_diff={k: v.tolist() for k, v in n_diff.items()},
                              binned=[len(s) for s in scaffolds_out], ambiguous=len(ambiguous),
                              timings=ctx.timings()))
        parent = subgenome_folder[:subgenome_folder.rfind("/")]
        if iteration < bootstrap:
            subgenome_folder = parent + "/bootstrap_%d" % (iteration + 1)
            create_path(subgenome_folder)
            for g, members in enumerate(scaffolds_out):
                if members:
                    with open(subgenome_folder + "/subgenome_%d.txt" % g, "w") as f:
                        f.write("\n".join(members))
        elif iteration == bootstrap:
            iteration += 1
            create_path(original_subgenome_path + "/finalResults")
            for g, members in enumerate(scaffolds_out):
                if members:
                    with open(original_subgenome_path + "/finalResults/subgenome_%d.txt" % g, "w") as f:
                        f.write("\n".join(members))
            run_final = 1
        if no_kill == 0:
            print("DEAD")
            break
        if run_final == 0:
            iteration += 1
            print("ITERATION")
            continue
        print("DONE")
        break
    return summaries
