"""In-memory hot path: split -> count -> select -> exact match -> CSR -> TF-IDF, all compute on the GPU through
the C ABI.  The file-based stage functions (stages.py / cli.py) and bench.py are thin callers of this module.

Host code here is integer bookkeeping only (chunk names, row order, row filter: polycracker.py:129-156, 284-287,
322-327); it never touches a base or a k-mer.
"""
import numpy as np

from . import binding


class ChunkLayout:
    """a1 + a6: chunk geometry, names, bed order, row universe (polycracker.py:129-156, 284-287)."""

    def __init__(self, names, seq_offsets, chunk_size):
        self.seq_names = list(names)
        self.seq_offsets = np.ascontiguousarray(seq_offsets, dtype=np.int64)
        self.chunk_size = int(chunk_size)
        seq_len = np.diff(self.seq_offsets)
        self.chunk_seq, self.chunk_start, self.chunk_end = binding.chunk_table(seq_len, self.chunk_size)
        base = self.seq_offsets[self.chunk_seq] if len(self.chunk_seq) else np.zeros(0, np.int64)
        self.chunk_begin_abs = base + self.chunk_start
        self.chunk_end_abs = base + self.chunk_end
        self.chunk_names = ["%s_%d_%d" % (self.seq_names[s], a, b)
                            for s, a, b in zip(self.chunk_seq.tolist(), self.chunk_start.tolist(),
                                               self.chunk_end.tolist())]
        self._rows_cache = {}

    @property
    def n_chunks(self):
        return len(self.chunk_names)

    def bed_rows(self):
        """correspondence.bed rows in `bedtools sort` order (chrom bytewise, start numeric; pc.py:154)."""
        rows = [(self.seq_names[s], a, b, n) for s, a, b, n in
                zip(self.chunk_seq.tolist(), self.chunk_start.tolist(), self.chunk_end.tolist(), self.chunk_names)]
        rows.sort(key=lambda r: (r[0].encode(), r[1]))
        return rows

    def original_scaffolds(self):
        return sorted(self.chunk_names)                       # findScaffolds, pc.py:284-287

    def rows(self, remove_non_chunk=1, min_chunk_threshold=0, min_chunk_size=0, prefilter=0, keep_names=None):
        """Row universe after the filter of polycracker.py:322-327 -> (scaffolds, chunk_row int32).
        keep_names (optional set) restricts rows further (a correspondence.bed that was pre-filtered)."""
        key = (bool(remove_non_chunk), bool(min_chunk_threshold), int(min_chunk_size), int(prefilter),
               None if keep_names is None else frozenset(keep_names))
        if key in self._rows_cache:
            return self._rows_cache[key]
        clen = np.abs(self.chunk_end - self.chunk_start)
        if remove_non_chunk and prefilter == 0:
            keep = clen == self.chunk_size
        elif min_chunk_threshold and prefilter == 0:
            keep = clen >= int(min_chunk_size)
        else:
            keep = np.ones(self.n_chunks, bool)
        idx = [i for i in np.flatnonzero(keep).tolist() if keep_names is None or self.chunk_names[i] in keep_names]
        idx.sort(key=lambda i: self.chunk_names[i])
        chunk_row = np.full(self.n_chunks, -1, np.int32)
        chunk_row[idx] = np.arange(len(idx), dtype=np.int32)
        self._rows_cache[key] = ([self.chunk_names[i] for i in idx], chunk_row)
        return self._rows_cache[key]


def rebuild_counts(data16, escapes=None):
    """float32 counts (what clusteringMatrix.npz holds, polycracker.py:355) from the compact uint16 hand-off; entries
    >= 65535 come from the escape list (positions, values)."""
    data = np.asarray(data16).view(np.uint16).astype(np.float32)
    if escapes is not None and len(escapes[0]):
        data[np.asarray(escapes[0])] = np.asarray(escapes[1], dtype=np.float32)
    return data


def rebuild_indices(indptr, dcol8, col_escapes, packed_counts=False):
    """int32 column ids from the one-byte gaps of pc_csr_get8: an escape (first entry of a row, gap >= 255) carries its
    absolute column, every other entry adds its gap to the entry before it."""
    d = np.asarray(dcol8)
    d = d.view(np.uint16 if d.dtype.itemsize == 2 else np.uint8).astype(np.int64)      # one- or two-byte gaps
    if packed_counts:                                           # pc_csr_get_g12c4: (gap << 4) | count
        d >>= 4
    n = len(d)
    if n == 0:
        return np.zeros(0, np.int32)
    pos, val = np.asarray(col_escapes[0], np.int64), np.asarray(col_escapes[1], np.int64)
    is_esc = np.zeros(n, bool)
    is_esc[pos] = True
    d[pos] = 0
    c = np.cumsum(d)
    absval = np.zeros(n, np.int64)
    absval[pos] = val
    last = np.maximum.accumulate(np.where(is_esc, np.arange(n), 0))      # entry 0 starts a row, so it is an escape
    return (c - c[last] + absval[last]).astype(np.int32)


def rebuild_counts8(cnt8, cnt_escapes, packed=None):
    if cnt8 is None:                                            # pc_csr_get_g12c4: the low 4 bits of the packed uint16
        data = (np.asarray(packed).view(np.uint16) & 15).astype(np.float32)
    else:
        data = np.asarray(cnt8).view(np.uint8).astype(np.float32)
    if cnt_escapes is not None and len(cnt_escapes[0]):
        data[np.asarray(cnt_escapes[0])] = np.asarray(cnt_escapes[1], dtype=np.float32)
    return data


def rebuild_tfidf(indptr, indices, data, idf, row_norm):
    """TF-IDF values from the two vectors they are a function of, with the kernel's own arithmetic (k_tfidf: one float32
    multiply, the float32 product divided by the row norm in double and rounded once; rows with norm 0 untouched) --
    bit-identical to pc_tfidf_get, from 4 * n_selected + 8 * n_rows bytes instead of 4 * nnz over PCIe."""
    indptr = np.asarray(indptr); indices = np.asarray(indices)
    x = np.asarray(data, dtype=np.float32) * np.asarray(idf, dtype=np.float32)[indices]          # float32 product
    nrm = np.repeat(np.asarray(row_norm, dtype=np.float64), np.diff(indptr))
    out = x.copy()
    nz = nrm > 0
    out[nz] = (x[nz].astype(np.float64) / nrm[nz]).astype(np.float32)
    return out


class HotPathResult:
    """Result of one pass.  The matrix values can arrive compact (data16 + escapes, idf + row_norm: 6 bytes per entry over
    PCIe instead of 12); `data` and `tfidf` then rebuild the float32 arrays on first use, bit-identical to the full
    hand-off.  Views into a StepPipeline arena are valid until the pipeline's next step() -- copy what must live longer."""

    def __init__(self, layout, k, scaffolds, kmer_keys, indptr, indices, data, tfidf, df=None, stats=None, timings=None,
                 data16=None, escapes=None, idf=None, row_norm=None, dcol8=None, cnt8=None, escapes8=None):
        self.layout, self.k, self.scaffolds, self.kmer_keys = layout, k, scaffolds, kmer_keys
        self.indptr, self._indices = indptr, indices          # CSR of raw counts (numpy, or torch CUDA tensors)
        self._data, self._tfidf = data, tfidf
        self.df = df
        self.stats = {} if stats is None else stats
        self.timings = {} if timings is None else timings
        self.data16, self.escapes, self.idf, self.row_norm = data16, escapes, idf, row_norm
        self.dcol8, self.cnt8, self.escapes8 = dcol8, cnt8, escapes8      # 2-byte hand-off (pc_csr_get8)

    @property
    def indices(self):
        if self._indices is None and self.dcol8 is not None:
            # (a two-byte dcol8 without a cnt8 array is the packed form of pc_csr_get_g12c4: gap << 4 | count)
            self._indices = rebuild_indices(self.indptr, self.dcol8, self.escapes8[0], packed_counts=self.cnt8 is None)
        return self._indices

    @property
    def data(self):
        if self._data is None and self.data16 is not None:
            self._data = rebuild_counts(self.data16, self.escapes)
        if self._data is None and (self.cnt8 is not None or self.dcol8 is not None):
            self._data = rebuild_counts8(self.cnt8, self.escapes8[1], packed=self.dcol8)
        return self._data

    @property
    def tfidf(self):
        if self._tfidf is None and self.idf is not None and self.row_norm is not None and self.indptr is not None:
            self._tfidf = rebuild_tfidf(self.indptr, self.indices, self.data, self.idf, self.row_norm)
        return self._tfidf

    @property
    def shape(self):
        return (len(self.scaffolds), len(self.kmer_keys))

    def kmers(self, orientation="min"):
        """Column labels.  'max' relabels every k-mer by its reverse complement (BBTools orientation);
        note the reference would then also sort by those labels -- see stages.relabel_max."""
        names = binding.kmer_decode(self.kmer_keys, self.k)
        if orientation == "max":
            comp = str.maketrans("ACGT", "TGCA")
            names = [n.translate(comp)[::-1] for n in names]
        return names

    def csr_matrix(self, values="counts"):
        import scipy.sparse as sps
        d = self.data if values == "counts" else self.tfidf
        idt = np.int32 if (len(self.indices) < 2**31 - 1) else np.int64
        return sps.csr_matrix((np.asarray(d), np.asarray(self.indices).astype(idt, copy=False),
                               np.asarray(self.indptr).astype(idt, copy=False)), shape=self.shape)


class HostArena:
    """Reusable page-locked host buffers for the D2H of results (a pageable numpy target would make the copy
    synchronous and several times slower).  Falls back to plain numpy when torch/CUDA is unavailable."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, n, dtype):
        dtype = np.dtype(dtype)
        need = max(int(n), 1) * dtype.itemsize
        buf = self._bufs.get(name)
        if buf is None or buf.nbytes < need:
            try:
                import torch
                t = torch.empty(int(need * 1.25) + 64, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
                buf = t.numpy()                            # the view keeps the pinned tensor alive
            except Exception:
                buf = np.empty(int(need * 1.25) + 64, np.uint8)
            self._bufs[name] = buf
        return buf[:int(n) * dtype.itemsize].view(dtype)


def effective_low_count(k, kmer_low_count):
    """kmer2Fasta only ever sees what writeKmerCount dumped: `mincount=3` for k <= 31 (polycracker.py:199),
    everything for k > 31 (jellyfish dump without -L, polycracker.py:203)."""
    return max(int(kmer_low_count), 3 if k <= 31 else 1)


def dump_floor(k):
    """Lowest count the reference's counter ever writes: kmercountexact `mincount=3` for k <= 31 (polycracker.py:199),
    jellyfish dump without -L for k > 31 (polycracker.py:203).  The shared-memory count keeps only k-mers that reach
    it (pc_tune keep_min), so the 2/3 of a genome's k-mers that occur once or twice never leave the SM."""
    return 3 if k <= 31 else 1


def run_hot_path(ctx, seq, seq_offsets, names, chunk_size=75000, k=23, kmer_low_count=100, high_bool=0,
                 kmer_high_count=2000000, sampling_sensitivity=1, remove_non_chunk=1, min_chunk_threshold=0,
                 min_chunk_size=0, prefilter=0, tfidf=True, table_capacity=0, fetch=True, layout=None, arena=None,
                 fuse_select=False, low_memory=False, compact=True):
    """One pass of the hot path on one GPU.  fuse_select: record threshold crossings during the count instead of
    re-scanning the table (pc_watch_threshold); measured slower on B200 (returning atomics cost more than the scan).

    seq: newline-free ASCII bases, numpy uint8 (pageable or pinned) or a torch uint8 tensor (CPU pinned or CUDA).
    Returns a HotPathResult; with fetch=False the CSR stays on the device and only sizes are returned.
    """
    if layout is None:
        layout = ChunkLayout(names, seq_offsets, chunk_size)
    scaffolds, chunk_row = layout.rows(remove_non_chunk, min_chunk_threshold, min_chunk_size, prefilter)
    ctx.load_sequences(seq, layout.chunk_begin_abs, layout.chunk_end_abs)
    ctx.watch_threshold(effective_low_count(k, kmer_low_count) if (fuse_select and not high_bool) else 0)
    ctx.tune(keep_min=dump_floor(k))
    try:
        ctx.count(k, table_capacity)
    finally:
        ctx.tune(keep_min=1)
    stats = ctx.count_info()
    if hasattr(ctx, "count_detail"):                      # how the count is held (before low_memory releases it)
        stats["count_detail"] = ctx.count_detail()
        stats["count_flags"] = ctx.count_flags()
    if low_memory:                                        # Gbp-scale genomes on one GPU: see pc_release
        ctx.release(ctx.REL_BUCKETS | ctx.REL_ASCII)
    n_sel = ctx.select(effective_low_count(k, kmer_low_count), kmer_high_count, bool(high_bool),
                       sampling_sensitivity)
    if low_memory:
        ctx.release(ctx.REL_TABLE)
    nnz = ctx.build_matrix(chunk_row, len(scaffolds))
    if low_memory:
        ctx.release(ctx.REL_HITS | ctx.REL_STREAM)          # scratch of K5/K6; the CSR is separate
    if tfidf:
        ctx.tfidf(len(scaffolds))
    stats.update(n_selected=n_sel, nnz=nnz, n_rows=len(scaffolds), n_chunks=layout.n_chunks)
    if not fetch:
        return HotPathResult(layout, k, scaffolds, np.empty(0, np.uint64), None, None, None, None, None, stats,
                             ctx.timings())
    return fetch_result(ctx, layout, k, scaffolds, n_sel, nnz, tfidf, stats, arena, compact=compact)


def fetch_result(ctx, layout, k, scaffolds, n_sel, nnz, tfidf, stats, arena=None, df=None, compact=True):
    """D2H of the repeat set and the matrix (into page-locked buffers when an arena is given).  compact=True (default): one
    byte per column gap and one per count plus escape lists (pc_csr_get8), and idf + row norms instead of the tf-idf array
    -- 2 bytes per entry instead of 12; compact=16: int32 columns + uint16 counts (6 bytes).  The int32 / float32 arrays are
    rebuilt on the host on first use (HotPathResult.indices / .data / .tfidf), bit-identical.  compact=False reads the
    arrays themselves (pc_csr_get / pc_tfidf_get)."""
    n_rows = len(scaffolds)
    compact = compact if hasattr(ctx, "csr8_into") else False   # (test stand-ins of the context only know the plain getters)
    if not hasattr(ctx, "csr_into"):
        keys = ctx.selected(n_sel)
        indptr, indices, data = ctx.csr(n_rows, nnz)
        return HotPathResult(layout, k, scaffolds, keys, indptr, indices, data, ctx.tfidf_values(nnz) if tfidf else None,
                             ctx.df(n_sel) if df is None and hasattr(ctx, "df") else df, stats, ctx.timings())

    def buf(name, n, dtype):
        return np.empty(n, dtype) if arena is None else arena.get(name, n, dtype)

    keys = buf("keys", n_sel, np.uint64)
    indptr = buf("indptr", n_rows + 1, np.int64)
    indices = buf("indices", nnz, np.int32) if compact in (False, 16) else None
    if n_sel:
        ctx.selected_into(keys)
    if df is None:
        df = buf("df", n_sel, np.int32)
        if n_sel:
            ctx.df_into(df)
    if not compact:
        data = buf("data", nnz, np.float32)
        ctx.csr_into(indptr, indices if nnz else None, data if nnz else None)
        tf = None
        if tfidf:
            tf = buf("tfidf", nnz, np.float32)
            if nnz:
                ctx.tfidf_into(tf)
        return HotPathResult(layout, k, scaffolds, keys, indptr, indices, data, tf, df, stats, ctx.timings())
    idf = row_norm = None
    if tfidf:
        idf = buf("idf", n_sel, np.float32)
        row_norm = buf("row_norm", n_rows, np.float64)
        ctx.tfidf_factors_into(idf if n_sel else None, row_norm if n_rows else None)
    if compact == 16:
        data16 = buf("data16", nnz, np.uint16)
        n_esc = ctx.csr16_into(indptr, indices if nnz else None, data16 if nnz else None)
        escapes = ctx.csr_escapes(n_esc) if n_esc else None
        return HotPathResult(layout, k, scaffolds, keys, indptr, indices, None, None, df, stats, ctx.timings(),
                             data16=data16, escapes=escapes, idf=idf, row_norm=row_norm)
    gdt = ctx.gap_dtype(n_sel, n_rows, nnz) if hasattr(ctx, "gap_dtype") else np.uint8
    wide = gdt == np.uint16                                     # gap and count then share one uint16 (pc_csr_get_g12c4)
    dcol8, cnt8 = buf("dcol16" if wide else "dcol8", nnz, gdt), (None if wide else buf("cnt8", nnz, np.uint8))
    n_col, n_cnt = ctx.csr8_into(indptr, dcol8 if (nnz or wide) else None, cnt8 if nnz else None)
    return HotPathResult(layout, k, scaffolds, keys, indptr, None, None, None, df, stats, ctx.timings(), idf=idf,
                         row_norm=row_norm, dcol8=dcol8, cnt8=cnt8, escapes8=ctx.csr8_escapes(n_col, n_cnt))
