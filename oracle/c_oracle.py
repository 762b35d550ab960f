"""ctypes wrapper of oracle/oracle.c (liboracle.so) -- TEST INFRASTRUCTURE ONLY.

PARITY STATUS: parity unpinned by the reference for a2/a4 (see oracle.c); pinned against py_oracle.py in
tests/test_oracle.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    """liboracle.so (this checker) and libpcsynth.so (the workload generator as a library of its own, for ref_path.py)."""
    src = os.path.join(HERE, "oracle.c")
    synth = os.path.join(HERE, "libpcsynth.so")
    if force or not os.path.exists(LIB) or not os.path.exists(synth) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "all"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        vp, i64 = C.c_void_p, C.c_int64
        _lib.orc_count.restype = i64
        _lib.orc_count.argtypes = [vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        _lib.orc_matrix.restype = i64
        _lib.orc_matrix.argtypes = [vp, vp, vp, vp, i64, C.c_int, vp, i64, vp, C.POINTER(vp), C.POINTER(vp)]
        _lib.orc_tfidf.restype = None
        _lib.orc_tfidf.argtypes = [vp, vp, vp, i64, i64, vp]
        _lib.orc_free.argtypes = [vp]
        _lib.orc_threads.restype = C.c_int
    return _lib


def threads():
    return int(lib().orc_threads())


def _take(ptr, n, dtype):
    if n == 0:
        out = np.empty(0, dtype)
    else:
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()
    lib().orc_free(ptr)
    return out


def count(seq, chunk_begin, chunk_end, k):
    """-> (keys uint64 ascending, counts uint64, n_windows)"""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cb = np.ascontiguousarray(chunk_begin, dtype=np.int64); ce = np.ascontiguousarray(chunk_end, dtype=np.int64)
    kp, cp, nw = C.c_void_p(), C.c_void_p(), C.c_int64()
    n = lib().orc_count(seq.ctypes.data, cb.ctypes.data, ce.ctypes.data, len(cb), k, C.byref(kp), C.byref(cp), C.byref(nw))
    if n < 0:
        raise ValueError("orc_count failed")
    return _take(kp, n, np.uint64), _take(cp, n, np.uint64), nw.value


def select(keys, counts, k, low, high_bool=0, high=2000000, sampling=1):
    """polycracker.py:220-238 on the sorted dump: mincount=3 pre-filter for k<=31 (pc.py:199), thresholds,
    every sampling-th survivor (sorted order, declared deviation)."""
    mincount = 3 if k <= 31 else 1
    keep = (counts >= mincount) & (counts >= low)
    if high_bool:
        keep &= counts <= high
    sel = keys[keep]
    return sel[::max(1, int(sampling))]


def matrix(seq, chunk_begin, chunk_end, row_chunk, k, sel_keys):
    """-> CSR (indptr int64, indices int32, data float32) with rows in row_chunk order"""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cb = np.ascontiguousarray(chunk_begin, dtype=np.int64); ce = np.ascontiguousarray(chunk_end, dtype=np.int64)
    rc = np.ascontiguousarray(row_chunk, dtype=np.int64)
    sel = np.ascontiguousarray(sel_keys, dtype=np.uint64)
    indptr = np.zeros(len(rc) + 1, np.int64)
    ip, dp = C.c_void_p(), C.c_void_p()
    nnz = lib().orc_matrix(seq.ctypes.data, cb.ctypes.data, ce.ctypes.data, rc.ctypes.data, len(rc), k, sel.ctypes.data,
                           len(sel), indptr.ctypes.data, C.byref(ip), C.byref(dp))
    return indptr, _take(ip, nnz, np.int32), _take(dp, nnz, np.float32)


def tfidf(indptr, indices, data, n_cols):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    out = np.zeros(len(data), np.float32)
    lib().orc_tfidf(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, len(indptr) - 1, n_cols, out.ctypes.data)
    return out
