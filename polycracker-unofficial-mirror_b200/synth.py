"""Deterministic synthetic polyploid genomes (bench + tests).  Thin wrapper over csrc/synth.cpp; the shapes of
BASELINE.json's configs are named here so bench.py, the tests and the CPU checker all see the same bytes."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .binding import SynthParams, check, lib

from .synth_config import CONFIGS, DEFAULTS  # noqa: F401  (re-exported)


def make_params(seed, genome_size, **kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return SynthParams(seed=seed, genome_size=genome_size, **d)


def layout(params):
    """[(name, length)] for every chromosome."""
    n = lib.pc_synth_n_chrom(C.byref(params))
    if n < 0:
        raise ValueError("bad synth parameters")
    out = []
    for c in range(n):
        ln = C.c_int64()
        name = C.create_string_buffer(64)
        check(lib.pc_synth_chrom_len(C.byref(params), c, C.byref(ln), name), "pc_synth_chrom_len")
        out.append((name.value.decode(), ln.value))
    return out


def generate(params, chroms=None, out=None, threads=4):
    """Returns (seq uint8 newline-free, offsets int64[n+1], names) for the chosen chromosomes (default: all).
    ``out`` may be a preallocated uint8 numpy array (e.g. a view of pinned memory)."""
    lay = layout(params)
    chroms = list(range(len(lay))) if chroms is None else list(chroms)
    lens = [lay[c][1] for c in chroms]
    offsets = np.zeros(len(chroms) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    if out is None:
        out = np.empty(total, np.uint8)
    assert out.dtype == np.uint8 and out.size >= total and out.flags.c_contiguous
    base = out.ctypes.data

    def one(i):
        check(lib.pc_synth_chrom(C.byref(params), chroms[i], C.c_void_p(base + int(offsets[i]))), "pc_synth_chrom")

    if threads > 1 and len(chroms) > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(chroms))))
    else:
        for i in range(len(chroms)):
            one(i)
    return out[:total], offsets, [lay[c][0] for c in chroms]


def small_params(seed=7, genome_size=200_000, n_subgenomes=2, n_chrom=2, **kw):
    """Test-sized genome: few short families so that repeats clear a threshold of ~10-30 at 1e5 bases."""
    d = dict(n_subgenomes=n_subgenomes, n_chrom=n_chrom, n_shared_families=3, n_specific_families=4,
             min_family_len=150, max_family_len=600, n_fraction=0.01, min_n_run=5, max_n_run=200,
             softmask_fraction=0.10)
    d.update(kw)
    return make_params(seed, genome_size, **d)
