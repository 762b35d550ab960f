"""The whole reference path on the CPU for bench.py's `--impl reference` arm and `cpu_baseline` leg -- TEST / MEASUREMENT
INFRASTRUCTURE ONLY (see oracle/oracle.c for the parity status).

Nothing here touches the product: the synthetic genome comes from oracle/libpcsynth.so (csrc/synth.cpp compiled as a
library of its own), the chunk geometry and the row order from the literal oracle (py_oracle.split_chunks /
find_scaffolds / filter_scaffolds = polycracker.py:129-156, 284-287, 322-327), counting / matching / TF-IDF from
oracle.c on all host cores.
"""
import ctypes as C
import importlib.util
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import c_oracle, py_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB = os.path.join(HERE, "libpcsynth.so")


def _config():
    """synth_config.py is a file of constants inside the product package; load it by path so that the package (and
    with it the CUDA library) is never imported here."""
    path = os.path.join(os.path.dirname(HERE), "polycracker-unofficial-mirror_b200", "synth_config.py")
    spec = importlib.util.spec_from_file_location("_pc_synth_config", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_size", C.c_int64), ("n_subgenomes", C.c_int32),
                ("n_chrom", C.c_int32), ("divergence", C.c_double), ("repeat_fraction", C.c_double),
                ("n_shared_families", C.c_int32), ("n_specific_families", C.c_int32),
                ("min_family_len", C.c_int32), ("max_family_len", C.c_int32), ("max_copy_mutation", C.c_double),
                ("n_fraction", C.c_double), ("min_n_run", C.c_int32), ("max_n_run", C.c_int32),
                ("softmask_fraction", C.c_double)]


_lib = None


def _synth():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB):
            c_oracle.build(force=True)
        _lib = C.CDLL(SYNTH_LIB)
        _lib.pc_synth_n_chrom.restype = C.c_int64
        _lib.pc_synth_n_chrom.argtypes = [C.POINTER(SynthParams)]
        _lib.pc_synth_chrom_len.argtypes = [C.POINTER(SynthParams), C.c_int64, C.POINTER(C.c_int64), C.c_char_p]
        _lib.pc_synth_chrom.argtypes = [C.POINTER(SynthParams), C.c_int64, C.c_void_p]
    return _lib


def workload(name, genome_size=0):
    """(cfg dict, SynthParams) of a named BASELINE workload; genome_size overrides the size (testing)."""
    cfgmod = _config()
    cfg = dict(cfgmod.CONFIGS[name])
    d = dict(cfgmod.DEFAULTS)
    d.update(n_subgenomes=cfg["n_subgenomes"])
    return cfg, SynthParams(seed=cfg["seed"], genome_size=int(genome_size or cfg["genome_size"]), **d)


def generate(params, max_bases=0, threads=8):
    """-> (seq uint8 newline-free, offsets int64[n + 1], names); max_bases > 0 keeps only the first chromosomes up to
    that many bases (a bounded sample of the same genome)."""
    lib = _synth()
    n = lib.pc_synth_n_chrom(C.byref(params))
    lay = []
    for c in range(n):
        ln = C.c_int64(); nm = C.create_string_buffer(64)
        if lib.pc_synth_chrom_len(C.byref(params), c, C.byref(ln), nm):
            raise ValueError("bad synth parameters")
        lay.append((nm.value.decode(), ln.value))
    chroms, total = [], 0
    for c, (_, ln) in enumerate(lay):
        if max_bases and total >= max_bases:
            break
        chroms.append(c); total += ln
    offsets = np.zeros(len(chroms) + 1, np.int64)
    offsets[1:] = np.cumsum([lay[c][1] for c in chroms])
    seq = np.empty(int(offsets[-1]), np.uint8)
    base = seq.ctypes.data

    def one(i):
        if lib.pc_synth_chrom(C.byref(params), chroms[i], C.c_void_p(base + int(offsets[i]))):
            raise ValueError("pc_synth_chrom failed")

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(one, range(len(chroms))))
    return seq, offsets, [lay[c][0] for c in chroms]


def run(seq, offsets, names, chunk_size, k, kmer_low_count, remove_non_chunk=1):
    """count -> select -> exact match -> CSR -> TF-IDF over one in-memory genome.  Returns a dict of the results."""
    fai = [(names[i], int(offsets[i + 1] - offsets[i])) for i in range(len(names))]
    bed = py_oracle.split_chunks(fai, chunk_size)                       # a1 (bedtools sort order)
    seq_start = {names[i]: int(offsets[i]) for i in range(len(names))}
    begin = np.array([seq_start[c] + s for c, s, e, _ in bed], np.int64)
    end = np.array([seq_start[c] + e for c, s, e, _ in bed], np.int64)
    order = np.argsort(begin, kind="stable")                            # oracle.c wants ascending, non-overlapping chunks
    begin, end = begin[order], end[order]
    chunk_names = [bed[i][3] for i in order.tolist()]
    keys, counts, n_windows = c_oracle.count(seq, begin, end, k)        # a2
    sel = c_oracle.select(keys, counts, k, kmer_low_count)              # a3
    scaffolds = py_oracle.filter_scaffolds(py_oracle.find_scaffolds(bed), chunk_size, remove_non_chunk, 0, 0)   # a6, a7 rows
    pos = {n: i for i, n in enumerate(chunk_names)}
    row_chunk = np.array([pos[s] for s in scaffolds], np.int64)
    indptr, indices, data = c_oracle.matrix(seq, begin, end, row_chunk, k, sel)       # a4, a5, a7
    tf = c_oracle.tfidf(indptr, indices, data, len(sel))                # a8
    return dict(keys=keys, counts=counts, n_windows=n_windows, sel=sel, scaffolds=scaffolds, indptr=indptr,
                indices=indices, data=data, tfidf=tf)
