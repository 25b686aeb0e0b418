"""ctypes front end of oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
CLI = os.path.join(_HERE, "_build", "smafa_oracle")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("subject", C.c_uint32), ("distance", C.c_uint32)]


class OraclePanic(Exception):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        l = C.CDLL(_LIB)
        l.orc_last_error.restype = C.c_char_p
        l.orc_query_encoded.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_size_t, C.c_long, C.c_long, C.c_long, C.c_int,
                                        C.POINTER(C.POINTER(Hit)), C.POINTER(C.c_size_t)]
        l.orc_cluster_encoded.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32,
                                          C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        l.orc_distances.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        l.orc_free.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def set_alphabet(alphabet):
    """0 = nucleotide (reference), 1 = protein (extension, parity unpinned).  Process-wide."""
    lib().orc_set_alphabet(int(alphabet))


def _opt(v):
    return -1 if v is None else int(v)


def distances(db, q):
    db = np.ascontiguousarray(db, dtype=np.uint64)
    q = np.ascontiguousarray(q, dtype=np.uint64)
    out = np.zeros(db.shape[0], dtype=np.uint64)
    lib().orc_distances(db.ctypes.data, db.shape[0], db.shape[1], q.ctypes.data, out.ctypes.data)
    return out.astype(np.int64)


def query(db, L, queries, q_len, m=None, k=None, r=None, threads=1):
    """-> uint32 array [n, 3] of (query, subject, distance) in the reference's print order."""
    db = np.ascontiguousarray(db, dtype=np.uint64)
    queries = np.ascontiguousarray(queries, dtype=np.uint64)
    W = db.shape[1] if db.ndim == 2 and db.shape[0] else (queries.shape[1] if queries.ndim == 2 else 0)
    hits = C.POINTER(Hit)()
    n = C.c_size_t(0)
    rc = lib().orc_query_encoded(db.ctypes.data, db.shape[0], W, L, queries.ctypes.data,
                                 queries.shape[0], q_len, _opt(m), _opt(k), _opt(r), threads,
                                 C.byref(hits), C.byref(n))
    if rc != 0:
        raise OraclePanic(lib().orc_last_error().decode())
    arr = np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint32)), shape=(n.value, 3)).copy() \
        if n.value else np.zeros((0, 3), dtype=np.uint32)
    lib().orc_free(hits)
    return arr


def cluster(enc, L, t):
    """-> (centroid_of int64 [n] with -1 for suppressed duplicates, n_centroids, n_comparisons)"""
    enc = np.ascontiguousarray(enc, dtype=np.uint64)
    n = enc.shape[0]
    cof = np.zeros(n, dtype=np.uint32)
    nc = C.c_size_t(0)
    cmp_ = C.c_uint64(0)
    rc = lib().orc_cluster_encoded(enc.ctypes.data, n, enc.shape[1] if n else 0, L, t,
                                   cof.ctypes.data, C.byref(nc), C.byref(cmp_))
    if rc != 0:
        raise OraclePanic(lib().orc_last_error().decode())
    out = cof.astype(np.int64)
    out[cof == 0xFFFFFFFF] = -1
    return out, nc.value, cmp_.value
