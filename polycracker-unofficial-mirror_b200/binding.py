"""ctypes binding of ``include/polycracker_b200.h`` -- the only way Python reaches the CUDA kernels.

There is no CPU fallback: if ``libpolycracker_b200.so`` is missing the import of this module raises, and every
compute call raises :class:`PolycrackerError` when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

PC_HOST, PC_DEVICE = 0, 1
PC_N_STAGES = 24
STAGE_NAMES = ["h2d", "pack", "table_init", "count", "select", "sort", "rep_table", "probe", "row_sort", "emit",
               "tfidf", "d2h", "export", "merge", "part_hist", "part_scatter", "sub_dict", "sub_compare", "sub_mapback",
               "part2", "count_fallback", "gram", "expand", "spare23"]
PC_MAX_GROUPS, PC_MAX_SUB_K = 8, 8


class PolycrackerError(RuntimeError):
    pass


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_size", C.c_int64), ("n_subgenomes", C.c_int32),
                ("n_chrom", C.c_int32), ("divergence", C.c_double), ("repeat_fraction", C.c_double),
                ("n_shared_families", C.c_int32), ("n_specific_families", C.c_int32),
                ("min_family_len", C.c_int32), ("max_family_len", C.c_int32), ("max_copy_mutation", C.c_double),
                ("n_fraction", C.c_double), ("min_n_run", C.c_int32), ("max_n_run", C.c_int32),
                ("softmask_fraction", C.c_double)]


def _load():
    path = _build.LIB
    # on the dev box (no GPU, the read-only reference tree is mounted) a library older than its sources is rebuilt
    # in-tree; on a GPU box the prebuilt .so travels with the snapshot and is used as is (file times do not survive the
    # copy, and N ranks must not race to rebuild it)
    dev_box = os.path.isdir("/root/reference") or os.environ.get("PC_DEV_REBUILD") == "1"
    if not os.path.exists(path) or (dev_box and _build.is_stale()):
        try:
            _build.build()
        except Exception as exc:  # pragma: no cover
            if not os.path.exists(path):
                raise ImportError("libpolycracker_b200.so is missing and could not be built: %s "
                                  "(the CUDA extension is mandatory, there is no CPU fallback)" % exc)
    return C.CDLL(path)


lib = _load()

_vp, _i64, _i32, _u32, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64
_P = C.POINTER
_SIGS = {
    "pc_version": (C.c_int, []),
    "pc_last_error": (C.c_char_p, []),
    "pc_device_count": (C.c_int, []),
    "pc_create": (C.c_int, [C.c_int, _P(_vp)]),
    "pc_destroy": (C.c_int, [_vp]),
    "pc_set_stream": (C.c_int, [_vp, _vp]),
    "pc_sync": (C.c_int, [_vp]),
    "pc_chunk_count": (_i64, [_vp, _i64, _i64]),
    "pc_chunk_table": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "pc_load_sequences": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _i64]),
    "pc_ascii_buffer": (C.c_int, [_vp, _i64, _P(_vp)]),
    "pc_packed_info": (C.c_int, [_vp, _P(_i64), _P(_i64)]),
    "pc_packed_get": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "pc_count": (C.c_int, [_vp, C.c_int, _i64]),
    "pc_count_info": (C.c_int, [_vp, _vp]),
    "pc_load_counts": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, C.c_int]),
    "pc_count_detail": (C.c_int, [_vp, _vp]),
    "pc_count_flags": (C.c_int, [_vp, _vp]),
    "pc_copy_async": (C.c_int, [_vp, _vp, _vp, _i64, _vp, C.c_int]),
    "pc_read_back": (C.c_int, [_vp, _vp, _vp, _i64]),
    "pc_count_histogram": (C.c_int, [_vp, _u32, _u32, _vp, C.c_int]),
    "pc_standard_scale": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_scaled_get": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_row_hits": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_densify_panel": (C.c_int, [_vp, C.c_int, _i64, C.c_int, _vp, _i64]),
    "pc_sub_plan": (C.c_int, [_vp, _i64, C.c_int]),
    "pc_sub_hist": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "pc_sub_hist_set": (C.c_int, [_vp, C.c_int, _vp, _i64]),
    "pc_dump_counts": (C.c_int, [_vp, _u32, _u32, _vp, _vp, _i64, C.c_int, _P(_i64)]),
    "pc_partition": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _P(_vp)]),
    "pc_count_buckets": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _i64]),
    "pc_kbuf_export": (C.c_int, [_vp, _i64, _vp]),
    "pc_peer_open": (C.c_int, [_vp, _vp, _P(_vp)]),
    "pc_peer_close": (C.c_int, [_vp, _vp]),
    "pc_count_buckets_from": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _i64]),
    "pc_skm_plan": (C.c_int, [_vp, C.c_int, _i64, C.c_int, _vp]),
    "pc_skm_partition": (C.c_int, [_vp, C.c_int, _vp, _vp, _P(_vp), _P(_vp)]),
    "pc_skm_export": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pc_skm_count_from": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "pc_host_skm_plan": (C.c_int, [C.c_int, _i64, _P(_i32)]),
    "pc_host_skm_records": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _i64, _P(_i64)]),
    "pc_host_skm_expand": (C.c_int, [_vp, _i64, C.c_int, _vp, _vp, _i64, _P(_i64)]),
    "pc_export_by_owner": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp]),
    "pc_merge_counts": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _i64, C.c_int]),
    "pc_release": (C.c_int, [_vp, C.c_int]),
    "pc_sub_begin": (C.c_int, [_vp, C.c_int]),
    "pc_sub_add_counts": (C.c_int, [_vp, C.c_int, _u32, _P(_i64)]),
    "pc_sub_union": (C.c_int, [_vp, _i64, _i64, _i64, _P(_i64), _P(_i64)]),
    "pc_sub_union_get": (C.c_int, [_vp, _vp, _vp, C.c_int, _P(_i64)]),
    "pc_sub_compare": (C.c_int, [_vp, C.c_double, C.c_double, _i64, _vp]),
    "pc_sub_diff_get": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _P(_i64)]),
    "pc_sub_mapback": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_select": (C.c_int, [_vp, _u32, _u32, C.c_int, _i64, _P(_i64)]),
    "pc_watch_threshold": (C.c_int, [_vp, _u32]),
    "pc_selected_get": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_set_selected": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _P(_i64)]),
    "pc_build_matrix": (C.c_int, [_vp, _vp, _i64, _P(_i64)]),
    "pc_matrix_from_hits": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _P(_i64)]),
    "pc_csr_get": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "pc_df_get": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_csr_get16": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _P(_i64)]),
    "pc_csr_escapes_get": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "pc_csr_get8": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp]),
    "pc_csr_get_gap16": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp]),
    "pc_csr_get_g12c4": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "pc_csr_escapes8_get": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int]),
    "pc_tfidf_factors_get": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "pc_tfidf": (C.c_int, [_vp, _i64, _vp, C.c_int]),
    "pc_tfidf_get": (C.c_int, [_vp, _vp, C.c_int]),
    "pc_gram_accumulate": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int64, _vp, C.c_int64]),
    "pc_gram_finish": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64]),
    "pc_timings": (C.c_int, [_vp, _vp]),
    "pc_stage_timeline": (C.c_int, [_vp, _vp, _vp]),
    "pc_launch_count": (C.c_int, [_vp, _P(_i64)]),
    "pc_reset_counters": (C.c_int, [_vp]),
    "pc_tune": (C.c_int, [_vp, C.c_char_p, _i64]),
    "pc_host_pack_word": (C.c_int, [_vp, C.c_int, _P(_u64), _P(_u32)]),
    "pc_host_canonical": (C.c_int, [_u64, C.c_int, _P(_u64)]),
    "pc_host_pack_chunks": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _P(_i64)]),
    "pc_host_extract": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp]),
    "pc_kmer_decode": (C.c_int, [_vp, _i64, C.c_int, _vp]),
    "pc_kmer_encode": (C.c_int, [_vp, _i64, C.c_int, _vp]),
    "pc_fasta_scan": (C.c_int, [C.c_char_p, _P(_i64), _P(_i64), _P(_i64)]),
    "pc_fasta_read": (C.c_int, [C.c_char_p, _vp, _vp, _vp]),
    "pc_write_kcount": (C.c_int, [C.c_char_p, _vp, _vp, _i64, C.c_int, C.c_int]),
    "pc_write_kmer_fasta": (C.c_int, [C.c_char_p, _vp, _i64, C.c_int, C.c_int]),
    "pc_write_hit_lines": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, _vp, _i64, _vp, C.c_int, _vp, _vp, _vp, C.c_int]),
    "pc_synth_n_chrom": (_i64, [_P(SynthParams)]),
    "pc_synth_chrom_len": (C.c_int, [_P(SynthParams), _i64, _P(_i64), C.c_char_p]),
    "pc_synth_chrom": (C.c_int, [_P(SynthParams), _i64, _vp]),
}
EXPORTED = sorted(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype, _fn.argtypes = _res, _args


def last_error():
    return (lib.pc_last_error() or b"").decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise PolycrackerError("%s failed (%d): %s" % (what or "polycracker_b200 call", rc, last_error()))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def ptr(x):
    """(address, location) of a numpy array (host) or torch tensor (host or device)."""
    if x is None:
        return None, PC_HOST
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr(), (PC_DEVICE if x.is_cuda else PC_HOST)
    a = np.ascontiguousarray(x)
    if a is not x:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data, PC_HOST


def device_count():
    return int(lib.pc_device_count())


# ------------------------------------------------------------------------------------------------ host helpers
def chunk_table(seq_len, chunk_size):
    """a1 geometry (polycracker.py:129-147): returns (chunk_seq, chunk_start, chunk_end) int64 arrays."""
    seq_len = np.ascontiguousarray(seq_len, dtype=np.int64)
    n = lib.pc_chunk_count(seq_len.ctypes.data, len(seq_len), int(chunk_size))
    if n < 0:
        raise PolycrackerError("bad chunk_size")
    cs, st, en = (np.empty(n, np.int64) for _ in range(3))
    check(lib.pc_chunk_table(seq_len.ctypes.data, len(seq_len), int(chunk_size), cs.ctypes.data, st.ctypes.data,
                             en.ctypes.data), "pc_chunk_table")
    return cs, st, en


def kmer_decode(keys, k):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.empty(len(keys) * k, np.uint8)
    check(lib.pc_kmer_decode(keys.ctypes.data, len(keys), k, out.ctypes.data), "pc_kmer_decode")
    if len(keys) == 0:
        return []
    return out.view("S%d" % k).astype(str).tolist()


def kmer_encode(kmers, k):
    if len(kmers) == 0:
        return np.empty(0, np.uint64)
    buf = np.frombuffer("".join(kmers).encode(), dtype=np.uint8)
    if len(buf) != len(kmers) * k:
        raise ValueError("all k-mers must have length %d" % k)
    out = np.empty(len(kmers), np.uint64)
    check(lib.pc_kmer_encode(buf.ctypes.data, len(kmers), k, out.ctypes.data), "pc_kmer_encode")
    return out


def read_fasta(path):
    """Returns (seq uint8[n_bases] newline-free, offsets int64[n_seqs+1], names list[str])."""
    ns, nb, nn = _i64(), _i64(), _i64()
    check(lib.pc_fasta_scan(path.encode(), C.byref(ns), C.byref(nb), C.byref(nn)), "pc_fasta_scan(%s)" % path)
    seq = np.empty(nb.value, np.uint8)
    offsets = np.empty(ns.value + 1, np.int64)
    names = np.empty(max(nn.value, 1), np.uint8)
    check(lib.pc_fasta_read(path.encode(), seq.ctypes.data, offsets.ctypes.data, names.ctypes.data), "pc_fasta_read")
    name_list = names[:nn.value].tobytes().decode().split("\n")[:ns.value]
    return seq, offsets, name_list


def write_kcount(path, keys, counts, k, append=False):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    check(lib.pc_write_kcount(path.encode(), keys.ctypes.data, counts.ctypes.data, len(keys), k, int(append)),
          "pc_write_kcount")


def write_kmer_fasta(path, keys, k, rc_labels=False):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    check(lib.pc_write_kmer_fasta(path.encode(), keys.ctypes.data, len(keys), k, int(rc_labels)),
          "pc_write_kmer_fasta")


def write_hit_lines(path, mode, indptr, indices, data, keys, k, row_names, row_len, append=False):
    """mode: 'sam' | 'bed' | 'merged' (see pc_write_hit_lines)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    blob = "".join(row_names).encode()
    name_off = np.zeros(len(row_names) + 1, np.int64)
    name_off[1:] = np.cumsum([len(n.encode()) for n in row_names])
    row_len = np.ascontiguousarray(row_len, dtype=np.int64)
    blob_arr = np.frombuffer(blob + b"\0", dtype=np.uint8)
    check(lib.pc_write_hit_lines(path.encode(), {"sam": 0, "bed": 1, "merged": 2}[mode], indptr.ctypes.data,
                                 indices.ctypes.data, data.ctypes.data, len(indptr) - 1, keys.ctypes.data, int(k),
                                 blob_arr.ctypes.data, name_off.ctypes.data, row_len.ctypes.data, int(append)),
          "pc_write_hit_lines")


def host_pack_word(bases):
    b = np.frombuffer(bases if isinstance(bases, bytes) else bases.encode(), dtype=np.uint8)
    w, m = _u64(), _u32()
    check(lib.pc_host_pack_word(b.ctypes.data, len(b), C.byref(w), C.byref(m)), "pc_host_pack_word")
    return w.value, m.value


def host_pack_chunks(seq, chunk_begin, chunk_end):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cb = np.ascontiguousarray(chunk_begin, dtype=np.int64); ce = np.ascontiguousarray(chunk_end, dtype=np.int64)
    n = _i64()
    check(lib.pc_host_pack_chunks(seq.ctypes.data, cb.ctypes.data, ce.ctypes.data, len(cb), None, None, None, 0,
                                  C.byref(n)), "pc_host_pack_chunks")
    words = np.zeros(n.value + 1, np.uint64); nmask = np.zeros(n.value + 1, np.uint32)
    w0 = np.zeros(len(cb) + 1, np.int64)
    check(lib.pc_host_pack_chunks(seq.ctypes.data, cb.ctypes.data, ce.ctypes.data, len(cb), words.ctypes.data,
                                  nmask.ctypes.data, w0.ctypes.data, n.value + 1, C.byref(n)), "pc_host_pack_chunks")
    return words, nmask, w0


def host_extract(words, nmask, k):
    """words/nmask include the trailing guard word.  -> (keys uint64[n*32], valid bool[n*32])"""
    n = len(words) - 1
    keys = np.zeros(n * 32, np.uint64); valid = np.zeros(n * 32, np.uint8)
    check(lib.pc_host_extract(words.ctypes.data, nmask.ctypes.data, n, k, keys.ctypes.data, valid.ctypes.data),
          "pc_host_extract")
    return keys, valid.astype(bool)


def host_skm_m(k, n_sub_total):
    """minimizer length the super-k-mer count picks for k and a number of sub-buckets (0: k too short)"""
    m = _i32()
    check(lib.pc_host_skm_plan(int(k), int(n_sub_total), C.byref(m)), "pc_host_skm_plan")
    return m.value


def host_skm_records(words, nmask, k, m, nmax, pb, nb2):
    """Host run of the super-k-mer extraction (the device's own inline routines).  words/nmask include the trailing
    guard word.  -> uint64[n_records, 2] (hi, lo)"""
    n_words = len(words) - 1
    n = _i64()
    args = (words.ctypes.data, nmask.ctypes.data, n_words, int(k), int(m), int(nmax), int(pb), int(nb2))
    check(lib.pc_host_skm_records(*args, None, 0, C.byref(n)), "pc_host_skm_records")
    recs = np.zeros((n.value, 2), np.uint64)
    check(lib.pc_host_skm_records(*args, recs.ctypes.data, n.value, C.byref(n)), "pc_host_skm_records")
    return recs


def host_skm_expand(recs, k):
    """-> (canonical k-mers uint64[n], destination uint32[n]) of the records, in record order"""
    recs = np.ascontiguousarray(recs, dtype=np.uint64)
    n = _i64()
    check(lib.pc_host_skm_expand(recs.ctypes.data, len(recs), int(k), None, None, 0, C.byref(n)), "pc_host_skm_expand")
    keys = np.zeros(n.value, np.uint64); dest = np.zeros(n.value, np.uint32)
    check(lib.pc_host_skm_expand(recs.ctypes.data, len(recs), int(k), keys.ctypes.data, dest.ctypes.data, n.value,
                                 C.byref(n)), "pc_host_skm_expand")
    return keys, dest


def host_canonical(kmer, k):
    out = _u64()
    check(lib.pc_host_canonical(int(kmer), k, C.byref(out)), "pc_host_canonical")
    return out.value


# ------------------------------------------------------------------------------------------------ GPU context
class Context:
    """One per GPU.  Thin object wrapper over the pc_* entry points; no logic of its own."""

    def __init__(self, device=0):
        h = _vp()
        check(lib.pc_create(int(device), C.byref(h)), "pc_create")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib.pc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:            # interpreter shutdown: the library handle may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        """Issue all work of this context on `cuda_stream` (a raw cudaStream_t, e.g.
        torch.cuda.current_stream().cuda_stream).  torch reports the legacy default stream as handle 0, which the
        C ABI reads as "the context's own stream"; it is mapped to cudaStreamLegacy (0x1) so that kernels launched
        here are ordered with torch / NCCL work on that stream.  None restores the context's own stream."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1                 # cudaStreamLegacy
        check(lib.pc_set_stream(self._h, _vp(handle)), "pc_set_stream")

    def sync(self):
        check(lib.pc_sync(self._h), "pc_sync")

    # K1
    def load_sequences(self, ascii_buf, chunk_begin, chunk_end, n_bytes=None):
        p, loc = ptr(ascii_buf)
        nb = int(n_bytes if n_bytes is not None else (ascii_buf.numel() if _is_torch(ascii_buf) else ascii_buf.size))
        cb = np.ascontiguousarray(chunk_begin, dtype=np.int64)
        ce = np.ascontiguousarray(chunk_end, dtype=np.int64)
        check(lib.pc_load_sequences(self._h, p, nb, loc, cb.ctypes.data, ce.ctypes.data, len(cb)), "pc_load_sequences")

    def ascii_buffer(self, n_bytes):
        out = _vp()
        check(lib.pc_ascii_buffer(self._h, int(n_bytes), C.byref(out)), "pc_ascii_buffer")
        return out.value

    def load_from_ascii_buffer(self, dev_ptr, n_bytes, chunk_begin, chunk_end):
        cb = np.ascontiguousarray(chunk_begin, dtype=np.int64)
        ce = np.ascontiguousarray(chunk_end, dtype=np.int64)
        check(lib.pc_load_sequences(self._h, dev_ptr, int(n_bytes), PC_DEVICE, cb.ctypes.data, ce.ctypes.data, len(cb)),
              "pc_load_sequences")

    def packed(self):
        nw, nc = _i64(), _i64()
        check(lib.pc_packed_info(self._h, C.byref(nw), C.byref(nc)), "pc_packed_info")
        words = np.empty(nw.value, np.uint64); nmask = np.empty(nw.value, np.uint32)
        w0 = np.empty(nc.value + 1, np.int64)
        check(lib.pc_packed_get(self._h, words.ctypes.data, nmask.ctypes.data, w0.ctypes.data, PC_HOST), "pc_packed_get")
        return words, nmask, w0

    # K3
    def count(self, k, capacity=0):
        check(lib.pc_count(self._h, int(k), int(capacity)), "pc_count")

    def load_counts(self, keys, counts, k):
        """Install externally counted k-mers (uint64 keys, uint32 counts; numpy or torch) as the count result."""
        pk, loc = ptr(keys)
        pc_, loc2 = ptr(counts)
        if loc != loc2:
            raise ValueError("keys and counts must live in the same place")
        n = int(keys.numel() if _is_torch(keys) else len(keys))
        check(lib.pc_load_counts(self._h, pk if n else None, pc_ if n else None, n, int(k), loc), "pc_load_counts")

    def count_info(self):
        info = np.zeros(8, np.int64)
        check(lib.pc_count_info(self._h, info.ctypes.data), "pc_count_info")
        return dict(capacity=int(info[0]), distinct=int(info[1]), valid_windows=int(info[2]),
                    table_bytes=int(info[3]), max_probe=int(info[4]), k=int(info[5]), partition_bits=int(info[7]))

    def sub_plan(self, total, n_buckets):
        return int(lib.pc_sub_plan(self._h, int(total), int(n_buckets)))

    def sub_hist(self, pb, nb2, hist_t):
        """hist_t: torch CUDA int64 tensor with 2^pb * nb2 elements (filled with this rank's sub-bucket sizes)."""
        assert hist_t.is_cuda and hist_t.is_contiguous() and hist_t.numel() == (1 << pb) * nb2
        check(lib.pc_sub_hist(self._h, int(pb), int(nb2), hist_t.data_ptr()), "pc_sub_hist")

    def sub_hist_set(self, nb2, hist_t):
        """Stage the owner's summed histogram (torch CUDA int64, n_buckets * nb2) for the next count_buckets[_from];
        the tensor must stay alive until that call returns."""
        assert hist_t.is_cuda and hist_t.is_contiguous()
        check(lib.pc_sub_hist_set(self._h, int(nb2), hist_t.data_ptr(), hist_t.numel()), "pc_sub_hist_set")

    # N3 standard scaling
    def standard_scale(self, nnz, row_scale=None):
        """StandardScaler(with_mean=False) of the built matrix (after the optional length normalisation row_scale[r] =
        1 / (len_r / 5000.), float64): float32 values in CSR order.  See pc_standard_scale."""
        rs = None if row_scale is None else np.ascontiguousarray(row_scale, dtype=np.float64)
        check(lib.pc_standard_scale(self._h, None if rs is None else rs.ctypes.data, PC_HOST), "pc_standard_scale")
        out = np.empty(nnz, np.float32)
        if nnz:
            check(lib.pc_scaled_get(self._h, out.ctypes.data, PC_HOST), "pc_scaled_get")
        return out

    def densify_panel(self, which, col0, ncols, out_t):
        """Columns [col0, col0 + ncols) of the built matrix as dense float32 into a torch CUDA tensor [n_rows, >= ncols]
        (which: 'counts' | 'tfidf' | 'scaled'); asynchronous on the context stream."""
        assert out_t.is_cuda and out_t.dim() == 2 and out_t.stride(1) == 1 and out_t.shape[1] >= ncols
        check(lib.pc_densify_panel(self._h, {"counts": 0, "tfidf": 1, "scaled": 2}[which], int(col0), int(ncols),
                                   _vp(out_t.data_ptr()), int(out_t.stride(0))), "pc_densify_panel")

    def gram_accumulate(self, panel_t, ncols, k_t):
        """K += P P^T (tiles on and above the diagonal) on the tensor cores: panel_t = torch CUDA float32 [n_rows, >= ncols]
        (rounded to TF32 in place), k_t = torch CUDA float32 [n_rows, n_rows]; asynchronous on the context stream."""
        assert panel_t.is_cuda and k_t.is_cuda and panel_t.stride(1) == 1 and k_t.stride(1) == 1
        check(lib.pc_gram_accumulate(self._h, _vp(panel_t.data_ptr()), int(panel_t.shape[0]), int(ncols), int(panel_t.stride(0)),
                                     _vp(k_t.data_ptr()), int(k_t.stride(0))), "pc_gram_accumulate")

    def gram_finish(self, k_t):
        check(lib.pc_gram_finish(self._h, _vp(k_t.data_ptr()), int(k_t.shape[0]), int(k_t.stride(0))), "pc_gram_finish")

    # N2 diagnostics
    def count_histogram(self, min_count=3, max_count=10000):
        """(counts, n_kmers) like sorted(Counter(kcount column).items()) restricted to [min_count, max_count], plus the
        number of k-mers above max_count (kcount_hist, polycracker.py:1468-1469)."""
        h = np.zeros(max_count - min_count + 2, np.uint64)
        check(lib.pc_count_histogram(self._h, int(min_count), int(max_count), h.ctypes.data, PC_HOST), "pc_count_histogram")
        nz = np.flatnonzero(h[:-1])
        return (nz + min_count).astype(np.int64), h[:-1][nz].astype(np.int64), int(h[-1])

    def row_hits(self, n_rows):
        """Repeat-k-mer occurrences per matrix row (number_repeatmers_per_subsequence, polycracker.py:1374)."""
        out = np.zeros(n_rows, np.int64)
        if n_rows:
            check(lib.pc_row_hits(self._h, out.ctypes.data, PC_HOST), "pc_row_hits")
        return out

    def to_host(self, tensor):
        """numpy copy of a small CUDA tensor, read back through the context's mailbox (pc_read_back): ordered on the
        context stream, never queued behind a bulk copy on a copy engine."""
        import torch
        t = tensor.contiguous()
        np_dtype = {torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8, torch.float32: np.float32,
                    torch.float64: np.float64}[t.dtype]
        out = np.empty(tuple(t.shape), np_dtype)
        if t.numel():
            check(lib.pc_read_back(self._h, out.ctypes.data, _vp(t.data_ptr()), t.numel() * t.element_size()),
                  "pc_read_back")
        return out

    def copy_async(self, dst, src, stream=None, n_ctas=0):
        """dst / src: torch tensors (CUDA, or page-locked CPU) of equal byte size; the copy is made by SMs on `stream`
        (raw handle; None = the context stream).  See pc_copy_async."""
        nb = src.numel() * src.element_size()
        assert nb == dst.numel() * dst.element_size() and src.is_contiguous() and dst.is_contiguous()
        check(lib.pc_copy_async(self._h, _vp(dst.data_ptr()), _vp(src.data_ptr()), nb, _vp(int(stream or 0)), int(n_ctas)),
              "pc_copy_async")

    def count_detail(self):
        d = np.zeros(8, np.int64)
        check(lib.pc_count_detail(self._h, d.ctypes.data), "pc_count_detail")
        return dict(kind="list" if d[0] == 1 else "table", kept=int(d[1]), keep_min=int(d[2]), sub_buckets=int(d[3]),
                    sub_per_bucket=int(d[4]), overflowed=int(d[5]), skm_records=int(d[6]), skm_m=int(d[7]))

    def count_flags(self):
        f = np.zeros(4, np.int64)
        check(lib.pc_count_flags(self._h, f.ctypes.data), "pc_count_flags")
        return dict(could_wrap=bool(f[0]), largest_unit=int(f[1]), overflowed_sub_buckets=int(f[2]), slots_per_table=int(f[3]))

    def dump_counts(self, min_count=1, max_count=0xFFFFFFFF):
        n = _i64()
        check(lib.pc_dump_counts(self._h, min_count, max_count, None, None, 0, PC_HOST, C.byref(n)), "pc_dump_counts")
        keys = np.empty(n.value, np.uint64); counts = np.empty(n.value, np.uint32)
        if n.value:
            check(lib.pc_dump_counts(self._h, min_count, max_count, keys.ctypes.data, counts.ctypes.data, n.value,
                                     PC_HOST, C.byref(n)), "pc_dump_counts")
        return keys, counts

    # N1: subgenome extraction inner loop (polycracker.py:692-865)
    def sub_begin(self, n_groups):
        check(lib.pc_sub_begin(self._h, int(n_groups)), "pc_sub_begin")
        self._sub_groups = int(n_groups)

    def sub_add_counts(self, group, min_count):
        """the current count table becomes subgenome `group`'s k-mer:count dictionary -> number of entries"""
        n = _i64()
        check(lib.pc_sub_add_counts(self._h, int(group), int(min_count), C.byref(n)), "pc_sub_add_counts")
        return n.value

    def sub_compare(self, ratio_threshold, default_value, sampling=1):
        """compareKmers for the pending k -> number of differential k-mers per subgenome"""
        n = np.zeros(PC_MAX_GROUPS, np.int64)
        check(lib.pc_sub_compare(self._h, float(ratio_threshold), float(default_value), int(sampling), n.ctypes.data),
              "pc_sub_compare")
        return n[:self._sub_groups].copy()

    def sub_union(self, ratio_threshold=20, default_value=1, max_val=100):
        """bio_hyp_class's filter over the union of the pending dictionaries -> (union size, keys uint64[n] ascending,
        counts uint32[n, n_groups])"""
        nu, nk = _i64(), _i64()
        check(lib.pc_sub_union(self._h, int(ratio_threshold), int(default_value), int(max_val), C.byref(nu), C.byref(nk)),
              "pc_sub_union")
        keys = np.empty(nk.value, np.uint64); counts = np.empty((nk.value, self._sub_groups), np.uint32)
        if nk.value:
            n = _i64()
            check(lib.pc_sub_union_get(self._h, keys.ctypes.data, counts.ctypes.data, PC_HOST, C.byref(n)), "pc_sub_union_get")
        return nu.value, keys, counts

    def sub_diff(self, k, group):
        """-> (keys uint64[n] ascending, counts uint32[n, n_groups]; 0 = absent from that subgenome's dictionary)"""
        n = _i64()
        check(lib.pc_sub_diff_get(self._h, int(k), int(group), None, None, PC_HOST, C.byref(n)), "pc_sub_diff_get")
        keys = np.empty(n.value, np.uint64); counts = np.empty((n.value, self._sub_groups), np.uint32)
        if n.value:
            check(lib.pc_sub_diff_get(self._h, int(k), int(group), keys.ctypes.data, counts.ctypes.data, PC_HOST,
                                      C.byref(n)), "pc_sub_diff_get")
        return keys, counts

    def sub_mapback(self, n_chunks):
        """-> int32[n_chunks, n_groups]: distinct window starts per loaded chunk matched by each subgenome's
        differential k-mers"""
        out = np.zeros((int(n_chunks), self._sub_groups), np.int32)
        check(lib.pc_sub_mapback(self._h, out.ctypes.data, PC_HOST), "pc_sub_mapback")
        return out

    def partition(self, k, pb):
        """-> (bucket_off int64[2^pb + 1], device pointer of the bucket-sorted k-mers)"""
        off = np.zeros((1 << pb) + 1, np.int64)
        keys = _vp()
        check(lib.pc_partition(self._h, int(k), int(pb), off.ctypes.data, C.byref(keys)), "pc_partition")
        return off, keys.value

    def count_buckets(self, k, pb, keys, seg_begin, seg_len, capacity=0):
        """keys: torch CUDA tensor (int64 view of the k-mers) or a raw device pointer; seg_begin/seg_len:
        int64[n_buckets, n_seg]."""
        sb = np.ascontiguousarray(seg_begin, dtype=np.int64); sl = np.ascontiguousarray(seg_len, dtype=np.int64)
        assert sb.ndim == 2 and sb.shape == sl.shape
        p = keys if isinstance(keys, int) else keys.data_ptr()
        check(lib.pc_count_buckets(self._h, int(k), int(pb), sb.shape[0], p, sb.ctypes.data, sl.ctypes.data, sb.shape[1],
                                   int(capacity)), "pc_count_buckets")

    def kbuf_export(self, n_keys):
        """-> 64-byte CUDA IPC handle (numpy uint8[64]) of the bucket buffer, pinned at >= n_keys k-mers"""
        h = np.zeros(64, np.uint8)
        check(lib.pc_kbuf_export(self._h, int(n_keys), h.ctypes.data), "pc_kbuf_export")
        return h

    def peer_open(self, handle):
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        out = _vp()
        check(lib.pc_peer_open(self._h, h.ctypes.data, C.byref(out)), "pc_peer_open")
        return out.value

    def count_buckets_from(self, k, pb, seg_base, seg_begin, seg_len, capacity=0):
        """seg_base: uint64[n_buckets, n_seg] device base pointers (own or peer-mapped bucket buffers)."""
        sp = np.ascontiguousarray(seg_base, dtype=np.uint64)
        sb = np.ascontiguousarray(seg_begin, dtype=np.int64); sl = np.ascontiguousarray(seg_len, dtype=np.int64)
        assert sp.shape == sb.shape == sl.shape and sb.ndim == 2
        check(lib.pc_count_buckets_from(self._h, int(k), int(pb), sb.shape[0], sp.ctypes.data, sb.ctypes.data,
                                        sl.ctypes.data, sb.shape[1], int(capacity)), "pc_count_buckets_from")

    # super-k-mer variant of the multi-GPU count (pc_skm_*)
    def skm_plan(self, k, windows_global, world=1):
        """-> int32[8] plan shared by all ranks (plan[5] == 0: not usable for this k / part_mode)"""
        plan = np.zeros(8, np.int32)
        check(lib.pc_skm_plan(self._h, int(k), int(windows_global), int(world), plan.ctypes.data), "pc_skm_plan")
        return plan

    def skm_partition(self, k, plan):
        """-> (bucket_off int64[2^pb + 1] in records, device pointer of the bucket-sorted records, device pointer of the
        packed sub-bucket histogram uint64[2^pb * nb2])"""
        plan = np.ascontiguousarray(plan, dtype=np.int32)
        off = np.zeros((1 << int(plan[0])) + 1, np.int64)
        recs, hist = _vp(), _vp()
        check(lib.pc_skm_partition(self._h, int(k), plan.ctypes.data, off.ctypes.data, C.byref(recs), C.byref(hist)),
              "pc_skm_partition")
        return off, recs.value, hist.value

    def skm_export(self, k, plan):
        """-> 64-byte CUDA IPC handle of the record buffer, sized for the loaded chunks"""
        plan = np.ascontiguousarray(plan, dtype=np.int32)
        h = np.zeros(64, np.uint8)
        check(lib.pc_skm_export(self._h, int(k), plan.ctypes.data, h.ctypes.data), "pc_skm_export")
        return h

    def peer_close(self, dev_ptr):
        check(lib.pc_peer_close(self._h, _vp(int(dev_ptr))), "pc_peer_close")

    def skm_count_from(self, k, plan, seg_base, seg_begin, seg_len, hist, recs=None):
        """seg_base: uint64[n_buckets, n_seg] device pointers of the senders' record buffers (None: all segments are
        relative to `recs`); hist: torch CUDA int64 tensor / device pointer of the packed histogram of these buckets."""
        plan = np.ascontiguousarray(plan, dtype=np.int32)
        sb = np.ascontiguousarray(seg_begin, dtype=np.int64); sl = np.ascontiguousarray(seg_len, dtype=np.int64)
        assert sb.ndim == 2 and sb.shape == sl.shape
        sp = None
        if seg_base is not None:
            sp = np.ascontiguousarray(seg_base, dtype=np.uint64)
            assert sp.shape == sb.shape
        hp = hist if isinstance(hist, int) else hist.data_ptr()
        rp = recs if (recs is None or isinstance(recs, int)) else recs.data_ptr()
        check(lib.pc_skm_count_from(self._h, int(k), plan.ctypes.data, sb.shape[0], rp, None if sp is None else sp.ctypes.data,
                                    sb.ctypes.data, sl.ctypes.data, sb.shape[1], hp), "pc_skm_count_from")

    def export_by_owner(self, world, keys_t, counts_t):
        """keys_t (int64/uint64 view) and counts_t (int32) are torch CUDA tensors with >= distinct elements."""
        offsets = np.zeros(world + 1, np.int64)
        check(lib.pc_export_by_owner(self._h, int(world), keys_t.data_ptr(), counts_t.data_ptr(), keys_t.numel(),
                                     offsets.ctypes.data), "pc_export_by_owner")
        return offsets

    def merge_counts(self, keys_t, counts_t, n, k, capacity=0, fresh=True):
        check(lib.pc_merge_counts(self._h, keys_t.data_ptr() if n else None, counts_t.data_ptr() if n else None,
                                  int(n), int(k), int(capacity), int(bool(fresh))), "pc_merge_counts")

    # K4
    REL_ASCII, REL_BUCKETS, REL_TABLE, REL_STREAM, REL_HITS = 1, 2, 4, 8, 16

    def release(self, what):
        check(lib.pc_release(self._h, int(what)), "pc_release")

    def watch_threshold(self, low):
        check(lib.pc_watch_threshold(self._h, int(low)), "pc_watch_threshold")

    def select(self, low, high=2000000, use_high=False, sampling=1):
        n = _i64()
        check(lib.pc_select(self._h, int(low), int(high), int(bool(use_high)), int(sampling), C.byref(n)), "pc_select")
        return n.value

    def selected(self, n):
        keys = np.empty(n, np.uint64)
        if n:
            check(lib.pc_selected_get(self._h, keys.ctypes.data, PC_HOST), "pc_selected_get")
        return keys

    def selected_into(self, tensor):
        p, loc = ptr(tensor)
        check(lib.pc_selected_get(self._h, p, loc), "pc_selected_get")

    def set_selected(self, keys, k, n=None):
        p, loc = ptr(keys)
        n = int(n if n is not None else (keys.numel() if _is_torch(keys) else len(keys)))
        out = _i64()
        check(lib.pc_set_selected(self._h, p if n else None, n, int(k), loc, C.byref(out)), "pc_set_selected")
        return out.value

    # K5 + K6
    def build_matrix(self, chunk_row, n_rows):
        cr = np.ascontiguousarray(chunk_row, dtype=np.int32)
        nnz = _i64()
        check(lib.pc_build_matrix(self._h, cr.ctypes.data, int(n_rows), C.byref(nnz)), "pc_build_matrix")
        return nnz.value

    def matrix_from_hits(self, row_offsets, cols, n_cols):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int64)
        cl = np.ascontiguousarray(cols, dtype=np.int32)
        nnz = _i64()
        check(lib.pc_matrix_from_hits(self._h, ro.ctypes.data, cl.ctypes.data if len(cl) else None, len(ro) - 1,
                                      int(n_cols), C.byref(nnz)), "pc_matrix_from_hits")
        return nnz.value

    def csr(self, n_rows, nnz):
        indptr = np.empty(n_rows + 1, np.int64); indices = np.empty(nnz, np.int32); data = np.empty(nnz, np.float32)
        check(lib.pc_csr_get(self._h, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, PC_HOST), "pc_csr_get")
        return indptr, indices, data

    def csr_into(self, indptr=None, indices=None, data=None):
        locs = {ptr(x)[1] for x in (indptr, indices, data) if x is not None}
        if len(locs) > 1:
            raise ValueError("all CSR outputs must live in the same place")
        check(lib.pc_csr_get(self._h, ptr(indptr)[0], ptr(indices)[0], ptr(data)[0], locs.pop() if locs else PC_HOST),
              "pc_csr_get")

    def csr16_into(self, indptr=None, indices=None, data16=None):
        """Compact hand-off: counts as uint16 (any 2-byte container: numpy uint16 / torch int16).  -> number of entries
        >= 65535 (stored as 65535; fetch them with csr_escapes)."""
        locs = {ptr(x)[1] for x in (indptr, indices, data16) if x is not None}
        if len(locs) > 1:
            raise ValueError("all CSR outputs must live in the same place")
        n = _i64()
        check(lib.pc_csr_get16(self._h, ptr(indptr)[0], ptr(indices)[0], ptr(data16)[0], locs.pop() if locs else PC_HOST,
                               C.byref(n)), "pc_csr_get16")
        return n.value

    def csr8_into(self, indptr=None, dcol8=None, cnt8=None):
        """Densest hand-off (2 bytes per entry).  -> (column escapes, count escapes) counts; see csr8_escapes."""
        locs = {ptr(x)[1] for x in (indptr, dcol8, cnt8) if x is not None}
        if len(locs) > 1:
            raise ValueError("all CSR outputs must live in the same place")
        n = np.zeros(2, np.int64)
        # the gap width follows the buffer: one byte (pc_csr_get8) or two (pc_csr_get_gap16; see gap_dtype)
        wide = dcol8 is not None and (dcol8.element_size() if hasattr(dcol8, "element_size") else dcol8.dtype.itemsize) == 2
        loc = locs.pop() if locs else PC_HOST
        if wide and cnt8 is None:                              # gap and count share the uint16 (pc_csr_get_g12c4)
            check(lib.pc_csr_get_g12c4(self._h, ptr(indptr)[0], ptr(dcol8)[0], loc, n.ctypes.data), "pc_csr_get_g12c4")
            return int(n[0]), int(n[1])
        fn, name = (lib.pc_csr_get_gap16, "pc_csr_get_gap16") if wide else (lib.pc_csr_get8, "pc_csr_get8")
        check(fn(self._h, ptr(indptr)[0], ptr(dcol8)[0], ptr(cnt8)[0], loc, n.ctypes.data), name)
        return int(n[0]), int(n[1])

    @staticmethod
    def gap_dtype(n_selected, n_rows, nnz):
        """Width of the column gaps of the compact hand-off: one byte while the mean gap of a row (columns / entries per
        row) stays below ~100, two bytes for a large column universe (where half of the one-byte gaps would escape)."""
        return np.uint16 if nnz > 0 and n_selected * max(n_rows, 1) > 96 * nnz else np.uint8

    def csr8_escapes(self, n_col, n_cnt):
        """-> ((positions int64, absolute columns int32), (positions int64, counts float32))"""
        cp, cv = np.empty(n_col, np.int64), np.empty(n_col, np.int32)
        kp, kv = np.empty(n_cnt, np.int64), np.empty(n_cnt, np.float32)
        if n_col or n_cnt:
            check(lib.pc_csr_escapes8_get(self._h, cp.ctypes.data, cv.ctypes.data, kp.ctypes.data, kv.ctypes.data, PC_HOST),
                  "pc_csr_escapes8_get")
        return (cp, cv), (kp, kv)

    def csr8_escapes_into(self, col_pos, col_val, cnt_pos, cnt_val):
        """Escape lists of the last csr8_into into caller buffers (torch tensors on the device: a device-to-device copy on
        the context stream, so a pipeline can move them to the host with its other bulk copies)."""
        locs = {ptr(x)[1] for x in (col_pos, col_val, cnt_pos, cnt_val) if x is not None}
        if len(locs) > 1:
            raise ValueError("all escape buffers must live in the same place")
        check(lib.pc_csr_escapes8_get(self._h, ptr(col_pos)[0], ptr(col_val)[0], ptr(cnt_pos)[0], ptr(cnt_val)[0],
                                      locs.pop() if locs else PC_HOST), "pc_csr_escapes8_get")

    def csr_escapes(self, n):
        pos = np.empty(n, np.int64); val = np.empty(n, np.float32)
        if n:
            check(lib.pc_csr_escapes_get(self._h, pos.ctypes.data, val.ctypes.data, PC_HOST), "pc_csr_escapes_get")
        return pos, val

    def tfidf_factors_into(self, idf=None, row_norm=None):
        """idf float32[n_selected], row_norm float64[n_rows] (same location): everything a host needs to rebuild the
        tf-idf values bit for bit (pipeline.rebuild_tfidf)."""
        locs = {ptr(x)[1] for x in (idf, row_norm) if x is not None}
        if len(locs) > 1:
            raise ValueError("idf and row_norm must live in the same place")
        check(lib.pc_tfidf_factors_get(self._h, ptr(idf)[0], ptr(row_norm)[0], locs.pop() if locs else PC_HOST),
              "pc_tfidf_factors_get")

    def df(self, n_cols):
        out = np.empty(n_cols, np.int32)
        if n_cols:
            check(lib.pc_df_get(self._h, out.ctypes.data, PC_HOST), "pc_df_get")
        return out

    def df_into(self, tensor):
        p, loc = ptr(tensor)
        check(lib.pc_df_get(self._h, p, loc), "pc_df_get")

    # K7
    def tfidf(self, n_rows_global=0, df_global=None):
        p, loc = ptr(df_global)
        check(lib.pc_tfidf(self._h, int(n_rows_global), p, loc), "pc_tfidf")

    def tfidf_values(self, nnz):
        out = np.empty(nnz, np.float32)
        if nnz:
            check(lib.pc_tfidf_get(self._h, out.ctypes.data, PC_HOST), "pc_tfidf_get")
        return out

    def tfidf_into(self, tensor):
        p, loc = ptr(tensor)
        check(lib.pc_tfidf_get(self._h, p, loc), "pc_tfidf_get")

    # measurement
    def timings(self):
        ms = np.zeros(PC_N_STAGES, np.float32)
        check(lib.pc_timings(self._h, ms.ctypes.data), "pc_timings")
        return {name: float(ms[i]) for i, name in enumerate(STAGE_NAMES)}

    def stage_timeline(self):
        """-> [(stage, start_ms, end_ms)] of the last step in start order (device idle time = the holes in between)"""
        a = np.zeros(PC_N_STAGES, np.float32); b = np.zeros(PC_N_STAGES, np.float32)
        check(lib.pc_stage_timeline(self._h, a.ctypes.data, b.ctypes.data), "pc_stage_timeline")
        out = [(name, float(a[i]), float(b[i])) for i, name in enumerate(STAGE_NAMES) if a[i] >= 0]
        return sorted(out, key=lambda x: x[1])

    def launch_count(self):
        n = _i64()
        check(lib.pc_launch_count(self._h, C.byref(n)), "pc_launch_count")
        return n.value

    def tune(self, **kw):
        for key, value in kw.items():
            check(lib.pc_tune(self._h, key.encode(), int(value)), "pc_tune(%s)" % key)

    def reset_counters(self):
        check(lib.pc_reset_counters(self._h), "pc_reset_counters")
