"""ctypes binding of include/smafa_b200.h and the host-side mirror of the reference's public
functions (smafa::makedb / query / cluster / count: reference src/lib.rs:137,198,378 and
src/cluster.rs:13).  There is deliberately no fallback: if the shared library is missing or no
sm_100 device is present, calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsmafa_b200.so")
CLI_PATH = os.path.join(_HERE, "bin", "smafa")

KERNEL_AUTO, KERNEL_POPC, KERNEL_MMA = 0, 1, 2
_KERNELS = {"auto": 0, "popc": 1, "mma": 2}

COMM_ID_BYTES = 128

# status codes that model reference panics (process exit 101)
_PANIC_CODES = {-1, -2, -3, -4, -21}


class SmafaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{status}: {message}")
        self.status = status
        self.message = message


class SmafaPanic(SmafaError):
    """The reference would have panicked (exit code 101) with this message."""


class SmafaCapacityError(SmafaError):
    """A caller-provided device buffer is too small; .needed = rows the answer has."""

    def __init__(self, needed, message):
        super().__init__("SMAFA_E_OOM", message)
        self.needed = needed


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("subject", C.c_uint32), ("distance", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("candidates", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("retries", C.c_uint32), ("kernel_used", C.c_uint32), ("scan_ms", C.c_float),
                ("total_ms", C.c_float), ("guess_bound", C.c_int32), ("rescanned", C.c_uint32),
                ("exchange_ms", C.c_float), ("union_degree", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib_path():
    return _LIB_PATH


def load_library():
    """Loads libsmafa_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: run `python -m smafa_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    l = C.CDLL(_LIB_PATH)
    vp, u64, u32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64
    l.smafa_abi_version.restype = C.c_int
    l.smafa_status_name.restype = C.c_char_p
    l.smafa_status_name.argtypes = [C.c_int]
    l.smafa_last_error.restype = C.c_char_p
    l.smafa_last_error.argtypes = [vp]
    l.smafa_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    l.smafa_ctx_destroy.argtypes = [vp]
    l.smafa_ctx_destroy.restype = None
    l.smafa_ctx_set_kernel.argtypes = [vp, C.c_int]
    l.smafa_ctx_set_candidate_capacity.argtypes = [vp, u64]
    l.smafa_ctx_set_alphabet.argtypes = [vp, C.c_int]
    l.smafa_db_upload.argtypes = [vp, vp, u64, u32, u64, C.POINTER(vp)]
    l.smafa_db_append.argtypes = [vp, vp, vp, u64]
    l.smafa_db_size.restype = u64
    l.smafa_db_size.argtypes = [vp]
    l.smafa_db_window_len.restype = u32
    l.smafa_db_window_len.argtypes = [vp]
    l.smafa_db_free.argtypes = [vp]
    l.smafa_db_free.restype = None
    l.smafa_distances.argtypes = [vp, vp, vp, u64, u32, vp]
    l.smafa_query.argtypes = [vp, vp, vp, u64, u32, i64, i64, C.POINTER(C.POINTER(Hit)), C.POINTER(u64),
                              C.POINTER(Stats)]
    l.smafa_query_dev.argtypes = [vp, vp, vp, u64, u32, i64, i64, vp, u64, C.POINTER(u64), vp, C.POINTER(Stats)]
    l.smafa_merge_dev.argtypes = [vp, vp, u64, i64, i64, C.POINTER(u64), vp]
    l.smafa_apply_limit_per_sequence.restype = u64
    l.smafa_apply_limit_per_sequence.argtypes = [vp, u64, vp, u32, u64, u32]
    l.smafa_cluster.argtypes = [vp, vp, u64, u32, u32, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(Stats)]
    l.smafa_free.argtypes = [vp]
    l.smafa_free.restype = None
    l.smafa_makedb_file.argtypes = [C.c_char_p, C.c_char_p]
    l.smafa_makedb_file_alphabet.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    l.smafa_query_file.argtypes = [vp, C.c_char_p, C.c_char_p, i64, i64, i64, C.c_int]
    l.smafa_cluster_file.argtypes = [vp, C.c_char_p, u32, C.c_int]
    l.smafa_count_files.argtypes = [C.POINTER(C.c_char_p), C.c_size_t, C.c_int]
    l.smafa_db_file_check.argtypes = [C.c_char_p]
    l.smafa_db_file_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]
    l.smafa_debug_mma_dump.argtypes = [vp, vp, vp, u64, u32, vp]
    l.smafa_debug_mma_peak.argtypes = [vp, u32, C.POINTER(C.c_double)]
    l.smafa_ctx_last_mma_k.argtypes = [vp]
    l.smafa_ctx_last_mma_k.restype = u32
    l.smafa_ctx_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int]
    l.smafa_ctx_device_count.argtypes = [vp]
    l.smafa_comm_unique_id.argtypes = [vp]
    l.smafa_ctx_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    l.smafa_ctx_comm_free.argtypes = [vp]
    l.smafa_ctx_comm_free.restype = None
    l.smafa_db_upload_shard.argtypes = [vp, vp, u64, u32, u64, u64, C.POINTER(vp)]
    l.smafa_db_upload_mapped.argtypes = [vp, vp, u64, u32, vp, u64, C.c_int, C.POINTER(vp)]
    l.smafa_group_order.argtypes = [vp, vp, u64, u32, vp, C.POINTER(u64)]
    l.smafa_query_sharded.argtypes = [vp, vp, vp, u64, u32, i64, i64, C.POINTER(C.POINTER(Hit)), C.POINTER(u64),
                                      C.POINTER(Stats)]
    l.smafa_query_sharded_dev.argtypes = [vp, vp, vp, u64, u32, i64, i64, vp, u64, C.POINTER(u64), vp, C.POINTER(Stats)]
    l.smafa_db_mma_k.restype = u32
    l.smafa_db_mma_k.argtypes = [vp]
    l.smafa_encode_symbol.restype = C.c_uint8
    l.smafa_encode_symbol.argtypes = [C.c_uint8]
    l.smafa_encode_window.argtypes = [C.c_char_p, C.c_size_t, vp, C.POINTER(C.c_size_t)]
    l.smafa_decode_window.argtypes = [vp, C.c_size_t, C.c_char_p]
    l.smafa_encode_symbol_alphabet.restype = C.c_uint8
    l.smafa_encode_symbol_alphabet.argtypes = [C.c_uint8, C.c_int]
    l.smafa_encode_window_alphabet.argtypes = [C.c_char_p, C.c_size_t, vp, C.POINTER(C.c_size_t), C.c_int]
    l.smafa_decode_window_alphabet.argtypes = [vp, C.c_size_t, C.c_char_p, C.c_int]
    _lib = l
    return l


def _opt(v):
    return -1 if v is None else int(v)


def _raise(status, ctx_handle=None):
    l = load_library()
    msg = l.smafa_last_error(ctx_handle).decode(errors="replace")
    name = l.smafa_status_name(status).decode()
    raise (SmafaPanic if status in _PANIC_CODES else SmafaError)(name, msg)


def _words(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2:
        raise ValueError("encoded windows must be a [n, W] uint64 array")
    return a


def comm_unique_id():
    """A fresh communicator id (rank 0 makes it, every rank of a sharded run gets a copy): bytes of COMM_ID_BYTES."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = load_library().smafa_comm_unique_id(buf)
    if rc:
        _raise(rc)
    return bytes(buf)


class Context:
    """One GPU (device = its number), or several GPUs of this process (device = a list: every db is row-sharded
    over them, smafa_ctx_create_multi).  kernel: 'auto' | 'popc' | 'mma'."""

    def __init__(self, device=0, kernel="auto"):
        self._l = load_library()
        self._h = C.c_void_p()
        kern = _KERNELS[kernel] if isinstance(kernel, str) else kernel
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self._l.smafa_ctx_create_multi(C.byref(self._h), devs, len(device), kern)
            self.devices = [int(d) for d in device]
            device = self.devices[0] if self.devices else 0
        else:
            rc = self._l.smafa_ctx_create(C.byref(self._h), int(device), kern)
            self.devices = [int(device)]
        if rc:
            _raise(rc)
        self.device = device

    # ---- sharded runs with one process per GPU (smafa_b200/dist.py drives this) ----
    def comm_init(self, comm_id, rank, world_size):
        """Joins the NCCL communicator `comm_id` (comm_unique_id() of rank 0).  Collective over the ranks."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        rc = self._l.smafa_ctx_comm_init(self._h, buf, int(rank), int(world_size))
        if rc:
            _raise(rc, self._h)

    def upload_shard(self, enc, L, subject_offset, total_rows):
        return Db(self, enc, L, subject_offset, total_rows=total_rows)

    def group_order(self, enc, L):
        """Similarity-grouped order of a whole db (smafa_group_order): -> (perm uint32 [D], clusters found; 0 = the db
        has too little structure and perm is the identity)."""
        e = _words(enc)
        perm = np.zeros(e.shape[0], dtype=np.uint32)
        nc = C.c_uint64(0)
        rc = self._l.smafa_group_order(self._h, e.ctypes.data, e.shape[0], int(L), perm.ctypes.data, C.byref(nc))
        if rc:
            _raise(rc, self._h)
        return perm, nc.value

    def upload_mapped(self, enc, L, subjects, total_rows, grouped=True):
        """Rows in the caller's order under explicit subject numbers (a shard of a db that was grouped as a whole)."""
        return Db(self, enc, L, 0, total_rows=total_rows, subjects=subjects, grouped=grouped)

    def query_sharded(self, db, q_enc, q_len, max_divergence=None, max_num_hits=None, return_stats=False):
        """Collective: every rank passes the same queries and receives the complete answer (uint32 [n, 3])."""
        q = _words(q_enc) if len(q_enc) else np.zeros((0, max(db.W, 1)), dtype=np.uint64)
        hits = C.POINTER(Hit)()
        n = C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_query_sharded(self._h, db.handle, q.ctypes.data, q.shape[0], int(q_len), _opt(max_divergence),
                                         _opt(max_num_hits), C.byref(hits), C.byref(n), C.byref(st))
        if rc:
            _raise(rc, self._h)
        arr = (np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint32)), shape=(n.value, 3)).copy()
               if n.value else np.zeros((0, 3), dtype=np.uint32))
        self._l.smafa_free(hits)
        return (arr, st.as_dict()) if return_stats else arr

    def query_sharded_ptr(self, db, q_host_ptr, Q, q_len, max_divergence=None, max_num_hits=None):
        """smafa_query_sharded on a raw host pointer (e.g. pinned memory) -> (rows as uint32 [n, 3], stats)."""
        hits = C.POINTER(Hit)()
        n = C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_query_sharded(self._h, db.handle, q_host_ptr, int(Q), int(q_len), _opt(max_divergence),
                                         _opt(max_num_hits), C.byref(hits), C.byref(n), C.byref(st))
        if rc:
            _raise(rc, self._h)
        arr = (np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint32)), shape=(n.value, 3)).copy()
               if n.value else np.zeros((0, 3), dtype=np.uint32))
        self._l.smafa_free(hits)
        return arr, st.as_dict()

    def query_sharded_dev(self, db, q_dev_ptr, Q, q_len, hits_dev_ptr, hits_capacity, max_divergence=None,
                          max_num_hits=None, stream=None):
        """Device-resident collective variant.  -> (n_hits, stats dict)."""
        n = C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_query_sharded_dev(self._h, db.handle, q_dev_ptr, int(Q), int(q_len), _opt(max_divergence),
                                             _opt(max_num_hits), hits_dev_ptr, int(hits_capacity), C.byref(n),
                                             stream, C.byref(st))
        if rc == -11 and n.value > hits_capacity:
            raise SmafaCapacityError(n.value, self._l.smafa_last_error(self._h).decode(errors="replace"))
        if rc:
            _raise(rc, self._h)
        return n.value, st.as_dict()

    @property
    def handle(self):
        return self._h

    def set_kernel(self, kernel):
        rc = self._l.smafa_ctx_set_kernel(self._h, _KERNELS[kernel] if isinstance(kernel, str) else kernel)
        if rc:
            _raise(rc, self._h)

    def set_alphabet(self, alphabet):
        """'nucleotide' (the reference) or 'protein' (extension; see include/smafa_b200.h): applies to the dbs
        uploaded afterwards, to cluster() input and to the file-level calls on this context."""
        a = {"nucleotide": 0, "protein": 1}[alphabet] if isinstance(alphabet, str) else int(alphabet)
        rc = self._l.smafa_ctx_set_alphabet(self._h, a)
        if rc:
            _raise(rc, self._h)

    def set_candidate_capacity(self, rows):
        self._l.smafa_ctx_set_candidate_capacity(self._h, int(rows))

    def upload(self, enc, L, subject_offset=0):
        return Db(self, enc, L, subject_offset)

    def close(self):
        if self._h:
            self._l.smafa_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- get_distances (reference src/lib.rs:71-89) ----
    def distances(self, db, q_enc, q_len):
        q = _words(q_enc) if len(q_enc) else np.zeros((0, max(db.W, 1)), dtype=np.uint64)
        out = np.zeros((q.shape[0], db.size), dtype=np.uint16)
        rc = self._l.smafa_distances(self._h, db.handle, q.ctypes.data, q.shape[0], int(q_len), out.ctypes.data)
        if rc:
            _raise(rc, self._h)
        return out

    # ---- query (reference src/lib.rs:238-314) ----
    def query(self, db, q_enc, q_len, max_divergence=None, max_num_hits=None, limit_per_sequence=None,
              return_stats=False):
        """-> uint32 [n, 3] rows (query, subject, distance) in the reference's print order."""
        q = _words(q_enc) if len(q_enc) else np.zeros((0, max(db.W, 1)), dtype=np.uint64)
        mode_b = max_num_hits is not None and max_num_hits != 1
        hits = C.POINTER(Hit)()
        n = C.c_uint64(0)
        st = Stats()
        nq = q.shape[0]
        if limit_per_sequence is not None and not mode_b and nq:
            nq = 1  # the reference panics on the first record (src/lib.rs:301-303), after its checks
        rc = self._l.smafa_query(self._h, db.handle, q.ctypes.data, nq, int(q_len), _opt(max_divergence),
                                 _opt(max_num_hits), C.byref(hits), C.byref(n), C.byref(st))
        if rc:
            _raise(rc, self._h)
        if limit_per_sequence is not None and not mode_b and q.shape[0]:
            self._l.smafa_free(hits)
            raise SmafaPanic("SMAFA_E_LIMIT_NEEDS_K", "limit_per_sequence is implemented unless max_num_hits > 1. "
                             "It can be implemented by analogy, just haven't gotten around to it.")
        cnt = n.value
        if limit_per_sequence is not None and cnt:
            if db.host_words is None:
                raise ValueError("limit_per_sequence needs the host copy of the db words")
            cnt = self._l.smafa_apply_limit_per_sequence(hits, cnt, db.host_words.ctypes.data, db.W,
                                                         db.subject_offset, int(limit_per_sequence))
        arr = (np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint32)), shape=(cnt, 3)).copy()
               if cnt else np.zeros((0, 3), dtype=np.uint32))
        self._l.smafa_free(hits)
        return (arr, st.as_dict()) if return_stats else arr

    def query_dev(self, db, q_dev_ptr, Q, q_len, hits_dev_ptr, hits_capacity, max_divergence=None,
                  max_num_hits=None, stream=None):
        """Device-resident variant.  -> (n_hits, stats dict)."""
        n = C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_query_dev(self._h, db.handle, q_dev_ptr, int(Q), int(q_len), _opt(max_divergence),
                                     _opt(max_num_hits), hits_dev_ptr, int(hits_capacity), C.byref(n),
                                     stream, C.byref(st))
        if rc == -11 and n.value > hits_capacity:
            raise SmafaCapacityError(n.value, self._l.smafa_last_error(self._h).decode(errors="replace"))
        if rc:
            _raise(rc, self._h)
        return n.value, st.as_dict()

    def query_ptr(self, db, q_host_ptr, Q, q_len, max_divergence=None, max_num_hits=None):
        """smafa_query on a raw host pointer (e.g. pinned memory) -> (rows as uint32 [n, 3], stats)."""
        hits = C.POINTER(Hit)()
        n = C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_query(self._h, db.handle, q_host_ptr, int(Q), int(q_len), _opt(max_divergence),
                                 _opt(max_num_hits), C.byref(hits), C.byref(n), C.byref(st))
        if rc:
            _raise(rc, self._h)
        arr = (np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint32)), shape=(n.value, 3)).copy()
               if n.value else np.zeros((0, 3), dtype=np.uint32))
        self._l.smafa_free(hits)
        return arr, st.as_dict()

    def merge_dev(self, cands_dev_ptr, n, max_divergence=None, max_num_hits=None, stream=None):
        out = C.c_uint64(0)
        rc = self._l.smafa_merge_dev(self._h, cands_dev_ptr, int(n), _opt(max_divergence), _opt(max_num_hits),
                                     C.byref(out), stream)
        if rc:
            _raise(rc, self._h)
        return out.value

    def mma_peak_tops(self, mmas_per_cta=20000):
        """Measured dense int8 tcgen05 rate of this GPU in TOP/s."""
        t = C.c_double(0)
        rc = self._l.smafa_debug_mma_peak(self._h, int(mmas_per_cta), C.byref(t))
        if rc:
            _raise(rc, self._h)
        return t.value

    @property
    def last_mma_k(self):
        """int8 contraction depth per window of the last tcgen05 scan (union-row operands: K / 2); 0 before any."""
        return int(self._l.smafa_ctx_last_mma_k(self._h))

    def debug_mma_dump(self, db, q_enc, bound):
        """Raw tcgen05 accumulators of the first db tile: int32 [128, 256] (see smafa_b200.h)."""
        q = _words(q_enc)
        out = np.zeros((128, 256), dtype=np.int32)
        rc = self._l.smafa_debug_mma_dump(self._h, db.handle, q.ctypes.data, q.shape[0], int(bound), out.ctypes.data)
        if rc:
            _raise(rc, self._h)
        return out

    # ---- cluster (reference src/cluster.rs:45-74) on de-duplicated encodings ----
    def cluster(self, enc, L, max_divergence, return_stats=False):
        e = _words(enc)
        cof = np.zeros(e.shape[0], dtype=np.uint32)
        nc, ncmp = C.c_uint64(0), C.c_uint64(0)
        st = Stats()
        rc = self._l.smafa_cluster(self._h, e.ctypes.data, e.shape[0], int(L), int(max_divergence), cof.ctypes.data,
                                   C.byref(nc), C.byref(ncmp), C.byref(st))
        if rc:
            _raise(rc, self._h)
        if return_stats:
            return cof, nc.value, ncmp.value, st.as_dict()
        return cof, nc.value, ncmp.value


class Db:
    """GPU-resident window set (the reference's WindowSet, src/lib.rs:54-60), optionally one
    row-shard of it (subject_offset = first global row)."""

    def __init__(self, ctx, enc, L, subject_offset=0, keep_host=True, total_rows=None, subjects=None, grouped=True):
        self.ctx = ctx
        self._l = ctx._l
        e = _words(enc) if len(enc) else np.zeros((0, max((L + 11) // 12, 1)), dtype=np.uint64)
        self.W = (L + 11) // 12
        self.L = L
        self.subject_offset = subject_offset
        self.host_words = e if keep_host else None
        self._h = C.c_void_p()
        if subjects is not None:
            sub = np.ascontiguousarray(subjects, dtype=np.uint32)
            assert sub.shape[0] == e.shape[0]
            rc = self._l.smafa_db_upload_mapped(ctx.handle, e.ctypes.data if e.shape[0] else None, e.shape[0], int(L),
                                                sub.ctypes.data, int(total_rows), 1 if grouped else 0, C.byref(self._h))
        elif total_rows is None:
            rc = self._l.smafa_db_upload(ctx.handle, e.ctypes.data if e.shape[0] else None, e.shape[0], int(L),
                                         int(subject_offset), C.byref(self._h))
        else:  # one shard of a row-sharded db (one process per GPU)
            rc = self._l.smafa_db_upload_shard(ctx.handle, e.ctypes.data if e.shape[0] else None, e.shape[0], int(L),
                                               int(subject_offset), int(total_rows), C.byref(self._h))
        if rc:
            _raise(rc, ctx.handle)

    @property
    def handle(self):
        return self._h

    @property
    def size(self):
        return self._l.smafa_db_size(self._h)

    def append(self, enc):
        """Adds windows at the end of the db (smafa_db_append; cluster grows its centroid db this way)."""
        e = _words(enc)
        rc = self._l.smafa_db_append(self.ctx.handle, self._h, e.ctypes.data if e.shape[0] else None, e.shape[0])
        if rc:
            _raise(rc, self.ctx.handle)
        if self.host_words is not None:
            self.host_words = np.concatenate([self.host_words, e])

    @property
    def mma_k(self):
        """Contraction depth of the tcgen05 operands (0: db not eligible for the MMA kernel)."""
        return int(self._l.smafa_db_mma_k(self._h))

    def close(self):
        if self._h:
            self._l.smafa_db_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- file-level mirror of the reference's public functions ---------------------------------

def load_db_file(db_path):
    """smafa::query's db load (reference src/lib.rs:206-218) -> (uint64 [n, W] words, window length or None)."""
    l = load_library()
    words = C.POINTER(C.c_uint64)()
    n, W, L = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
    rc = l.smafa_db_file_load(os.fsencode(db_path), C.byref(words), C.byref(n), C.byref(W), C.byref(L))
    if rc:
        _raise(rc)
    arr = (np.ctypeslib.as_array(words, shape=(n.value, W.value)).copy() if n.value * W.value
           else np.zeros((n.value, W.value), dtype=np.uint64))
    l.smafa_free(words)
    return arr, (L.value or None)


def makedb(subject_fasta, db_path, protein=False):
    """smafa::makedb (reference src/lib.rs:137-165).  Host only.  protein=True: the amino-acid extension."""
    rc = load_library().smafa_makedb_file_alphabet(os.fsencode(subject_fasta), os.fsencode(db_path), 1 if protein else 0)
    if rc:
        _raise(rc)


def query(ctx, db_path, query_fasta, max_divergence=None, max_num_hits=None, limit_per_sequence=None, out_fd=1):
    """smafa::query (reference src/lib.rs:198-325); TSV goes to out_fd."""
    rc = load_library().smafa_query_file(ctx.handle, os.fsencode(db_path), os.fsencode(query_fasta),
                                         _opt(max_divergence), _opt(max_num_hits), _opt(limit_per_sequence), out_fd)
    if rc:
        _raise(rc)


def cluster(ctx, input_fasta, max_divergence, out_fd=1):
    """smafa::cluster (reference src/cluster.rs:13-94)."""
    rc = load_library().smafa_cluster_file(ctx.handle, os.fsencode(input_fasta), int(max_divergence), out_fd)
    if rc:
        _raise(rc)


def count(paths, out_fd=1):
    """smafa::count (reference src/lib.rs:378-398).  Host only."""
    arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    rc = load_library().smafa_count_files(arr, len(paths), out_fd)
    if rc:
        _raise(rc)
