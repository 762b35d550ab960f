"""Shared helpers for the tests: run the CPU checkers on the same inputs as the CUDA path."""
import numpy as np

from oracle import c_oracle, py_oracle
from polycracker_b200 import binding
from polycracker_b200.pipeline import ChunkLayout


def to_sequences(seq, offsets, names):
    s = seq.tobytes().decode("latin-1")
    return [(names[i], s[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(names))]


def from_sequences(sequences):
    names = [n for n, _ in sequences]
    lens = [len(s) for _, s in sequences]
    offsets = np.zeros(len(lens) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    seq = np.frombuffer("".join(s for _, s in sequences).encode("latin-1"), dtype=np.uint8).copy()
    return seq, offsets, names


def c_oracle_path(seq, offsets, names, chunk_size, k, low, high_bool=0, high=2000000, sampling=1,
                  remove_non_chunk=1, min_chunk_threshold=0, min_chunk_size=0):
    """Whole path through oracle.c; row order / filter from the literal oracle's functions."""
    lay = ChunkLayout(names, offsets, chunk_size)          # integer geometry only, checked in test_host.py
    keys, counts, n_windows = c_oracle.count(seq, lay.chunk_begin_abs, lay.chunk_end_abs, k)
    sel = c_oracle.select(keys, counts, k, low, high_bool, high, sampling)
    original = py_oracle.find_scaffolds([(None, None, None, n) for n in lay.chunk_names])
    scaffolds = py_oracle.filter_scaffolds(original, chunk_size, remove_non_chunk, min_chunk_threshold, min_chunk_size)
    pos = {n: i for i, n in enumerate(lay.chunk_names)}
    row_chunk = np.array([pos[s] for s in scaffolds], np.int64)
    indptr, indices, data = c_oracle.matrix(seq, lay.chunk_begin_abs, lay.chunk_end_abs, row_chunk, k, sel)
    tf = c_oracle.tfidf(indptr, indices, data, len(sel))
    return dict(keys=keys, counts=counts, n_windows=n_windows, sel=sel, scaffolds=scaffolds, indptr=indptr,
                indices=indices, data=data, tfidf=tf)


def assert_tfidf_close(got, want, rtol=1e-6):
    """north_star tolerance for the floating-point stage: 1e-6 relative (fp32)."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    assert got.shape == want.shape
    if got.size:
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert err.max() <= rtol, "tf-idf max relative error %.3g > %.1g" % (err.max(), rtol)
