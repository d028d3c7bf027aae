"""ctypes binding of include/syzgy_b200.h (libsyzgy_b200.so, built in-tree).

This is the same C ABI the Go cgo shim binds.  There is no fallback of
any kind: if the shared library is missing or no B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsyzgy_b200.so")

EUCLIDEAN = 0
COSINE = 1
F_DEFAULT = 0
F_NO_FP64_VERIFY = 1
MISSING_DISTANCE = -1.0
MAX_K = 224
MAX_DIM = 16384
OPT_STREAMS = 1
OPT_TIMING = 2
OPT_MIN_CANDIDATE_MODE = 3
OPT_SCAN_WARPS = 4
OPT_SCAN_STAGES = 5
OPT_SCAN_TILE_CHUNKS = 6
OPT_DIGITS = 7
OPT_BATCH_TENSOR = 8
OPT_COMBINE = 9
OPT_BATCH_MIN_QUERIES = 10
OPT_GRAPHS = 11
OPT_TRACE_BUFFER = 12

# szg_filter_op opcodes (include/syzgy_b200.h SZG_FOP_*)
(FOP_COL, FOP_NUM, FOP_STR, FOP_BOOL, FOP_NULL, FOP_EQ, FOP_NE, FOP_LT, FOP_LE, FOP_GT, FOP_GE, FOP_AND, FOP_OR, FOP_NOT,
 FOP_IN, FOP_NOT_IN, FOP_CONTAINS, FOP_STARTS_WITH, FOP_ENDS_WITH, FOP_STR_TABLE, FOP_EXISTS, FOP_NOT_EXISTS) = range(1, 23)

# every symbol include/syzgy_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "szg_last_error", "szg_create", "szg_destroy", "szg_reserve", "szg_upsert", "szg_encode", "szg_remove", "szg_count",
    "szg_mask_create", "szg_mask_destroy", "szg_search_topk", "szg_search_batch", "szg_search_radius", "szg_result_count",
    "szg_result_fetch", "szg_result_free", "szg_rescore", "szg_search_topk_dev", "szg_search_batch_dev", "szg_merge_topk_dev",
    "szg_fill_synthetic", "szg_fetch_codes", "szg_get_stats", "szg_set_option", "szg_last_scan_times_ms",
    "szg_spanfile_open", "szg_spanfile_close", "szg_spanfile_get_info", "szg_spanfile_ids", "szg_spanfile_record",
    "szg_spanfile_load", "szg_meta_upsert", "szg_filter_mask", "szg_meta_dictionary_size", "szg_meta_dictionary_get",
    "szg_create_sharded", "szg_search_radius_batch", "szg_rescore_batch",
]


class SzgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"syzgy_b200 error {code}: {msg}")
        self.code = code


class SpanFileInfo(C.Structure):
    _fields_ = [
        ("file_bytes", C.c_uint64), ("records", C.c_uint64), ("spans_active", C.c_uint64), ("spans_free", C.c_uint64),
        ("spans_corrupt", C.c_uint64), ("free_bytes", C.c_uint64), ("foreign_records", C.c_uint64),
        ("next_sequence", C.c_uint32), ("has_header", C.c_int32), ("distance_method", C.c_int32),
        ("dimension_count", C.c_int32), ("quantization", C.c_int32), ("name", C.c_char * 256),
    ]


class MetaValue(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("str_len", C.c_uint32), ("num", C.c_double), ("str", C.c_void_p)]


class FilterOp(C.Structure):
    _fields_ = [("op", C.c_uint32), ("arg", C.c_uint32), ("num", C.c_double), ("str", C.c_void_p), ("str_len", C.c_uint32),
                ("table_len", C.c_uint32), ("table", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("escalations", C.c_uint64), ("uncertain_results", C.c_uint64),
        ("batch_queries", C.c_uint64),
        ("device_bytes", C.c_uint64), ("live_rows", C.c_uint64), ("slots", C.c_uint64),
        ("rowbytes", C.c_uint32), ("pitch", C.c_uint32), ("sm_count", C.c_uint32), ("scan_grid", C.c_uint32),
        ("scan_block", C.c_uint32), ("scan_stages", C.c_uint32), ("scan_tile_bytes", C.c_uint32),
        ("scan_smem_bytes", C.c_uint32), ("shards", C.c_uint32), ("combined_queries", C.c_uint64),
        ("graph_launches", C.c_uint64),
    ]


_lib = None


def load():
    """Loads the CUDA library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C syzgydb_b200/csrc` (or __graft_entry__.build()). "
            "syzgydb_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p, f64p, f32p = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_float))
    L.szg_last_error.restype = C.c_char_p
    L.szg_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.szg_create_sharded.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.szg_destroy.argtypes = [vp]
    L.szg_reserve.argtypes = [vp, C.c_uint64]
    L.szg_upsert.argtypes = [vp, u64p, u8p, C.c_uint64]
    L.szg_encode.argtypes = [vp, u64p, f64p, C.c_uint64, u8p, C.c_int]
    L.szg_remove.argtypes = [vp, u64p, C.c_uint64, u64p]
    L.szg_count.argtypes = [vp, u64p]
    L.szg_mask_create.argtypes = [vp, u64p, u8p, C.c_uint64, C.POINTER(C.c_int)]
    L.szg_mask_destroy.argtypes = [vp, C.c_int]
    L.szg_meta_upsert.argtypes = [vp, u64p, C.c_uint64, u8p, u32p, C.c_uint32, C.POINTER(MetaValue)]
    L.szg_filter_mask.argtypes = [vp, C.POINTER(FilterOp), C.c_uint32, C.POINTER(C.c_int)]
    L.szg_meta_dictionary_size.argtypes = [vp, u32p]
    L.szg_meta_dictionary_get.argtypes = [vp, C.c_uint32, C.POINTER(C.c_void_p), u32p]
    L.szg_search_topk.argtypes = [vp, f64p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, u64p, f64p, u32p, u64p]
    L.szg_search_batch.argtypes = L.szg_search_topk.argtypes
    L.szg_search_radius.argtypes = [vp, f64p, C.c_double, C.c_int, C.c_uint32, C.POINTER(vp), u64p]
    L.szg_search_radius_batch.argtypes = [vp, f64p, C.c_uint32, f64p, C.c_int, C.c_uint32, C.POINTER(vp), u64p]
    L.szg_rescore_batch.argtypes = [vp, f64p, C.c_uint32, u64p, u64p, f64p]
    L.szg_result_count.argtypes = [vp, u64p]
    L.szg_result_fetch.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, f64p]
    L.szg_result_free.argtypes = [vp]
    L.szg_result_free.restype = None
    L.szg_rescore.argtypes = [vp, f64p, u64p, C.c_uint64, f64p]
    L.szg_spanfile_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.szg_spanfile_close.argtypes = [vp]
    L.szg_spanfile_get_info.argtypes = [vp, C.POINTER(SpanFileInfo)]
    L.szg_spanfile_ids.argtypes = [vp, u64p, C.c_uint64, u64p]
    L.szg_spanfile_record.argtypes = [vp, C.c_uint64, C.POINTER(u8p), u64p, C.POINTER(u8p), u64p]
    L.szg_spanfile_load.argtypes = [vp, vp, u64p]
    L.szg_search_topk_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, vp, vp, vp, vp, vp]
    L.szg_search_batch_dev.argtypes = L.szg_search_topk_dev.argtypes
    L.szg_merge_topk_dev.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]
    L.szg_fill_synthetic.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.szg_fetch_codes.argtypes = [vp, u64p, C.c_uint64, u8p]
    L.szg_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.szg_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.szg_last_scan_times_ms.argtypes = [vp, f32p, C.c_uint32, u32p]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise SzgError(rc, load().szg_last_error().decode("utf-8", "replace"))


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def vector_size(quant: int, dims: int) -> int:
    """getVectorSize, collection.go:796-811"""
    if quant == 4:
        return (dims + 1) // 2
    if quant in (8, 16, 32, 64):
        return dims * (quant // 8)
    raise ValueError("Unsupported quantization level")


class Index:
    """GPU mirror of one collection (or of one row shard).  Thin, 1:1 over the C ABI."""

    def __init__(self, dim: int, quant: int, metric: int, device: int = 0, devices=None):
        """devices=[d0, d1, ...]: one handle over several GPUs (szg_create_sharded); results do not depend on the list."""
        self._L = load()
        self._h = C.c_void_p()
        if devices is not None:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(self._L.szg_create_sharded(dim, quant, metric, devs, len(devices), C.byref(self._h)))
            device = int(devices[0]) if len(devices) else 0
        else:
            _check(self._L.szg_create(dim, quant, metric, device, C.byref(self._h)))
        self.dim, self.quant, self.metric, self.device = dim, (quant or 64), metric, device
        self.devices = list(devices) if devices is not None else [device]
        self.rowbytes = vector_size(self.quant, dim)

    # -- lifecycle
    def close(self):
        if self._h:
            self._L.szg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- mutation
    def reserve(self, nrows: int):
        _check(self._L.szg_reserve(self._h, nrows))

    def upsert(self, ids, codes):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if codes.size != ids.size * self.rowbytes:
            raise ValueError(f"codes must hold {ids.size} x {self.rowbytes} bytes")
        _check(self._L.szg_upsert(self._h, _p(ids, C.c_uint64), _p(codes, C.c_uint8), ids.size))

    def encode(self, vectors, ids=None, upsert: bool = False) -> np.ndarray:
        """encodeDocument on the device (szg_encode): float64 vectors -> stream-1 bytes; upsert=True also mirrors
        them under `ids`."""
        v = np.ascontiguousarray(vectors, dtype=np.float64).reshape(-1, self.dim)
        out = np.zeros((v.shape[0], self.rowbytes), dtype=np.uint8)
        idp = None
        if upsert:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            if ids.size != v.shape[0]:
                raise ValueError("one id per vector")
            idp = _p(ids, C.c_uint64)
        _check(self._L.szg_encode(self._h, idp, _p(v, C.c_double), v.shape[0], _p(out, C.c_uint8), 1 if upsert else 0))
        return out

    def remove(self, ids) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        n = C.c_uint64(0)
        _check(self._L.szg_remove(self._h, _p(ids, C.c_uint64), ids.size, C.byref(n)))
        return n.value

    def count(self) -> int:
        n = C.c_uint64(0)
        _check(self._L.szg_count(self._h, C.byref(n)))
        return n.value

    def fill_synthetic(self, seed: int, row0: int, nrows: int):
        _check(self._L.szg_fill_synthetic(self._h, seed, row0, nrows))

    def fetch_codes(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.zeros((ids.size, self.rowbytes), dtype=np.uint8)
        _check(self._L.szg_fetch_codes(self._h, _p(ids, C.c_uint64), ids.size, _p(out, C.c_uint8)))
        return out

    def mask_create(self, ids, passed) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        passed = np.ascontiguousarray(passed, dtype=np.uint8)
        if ids.size != passed.size:
            raise ValueError("ids and pass must have the same length")
        mid = C.c_int(0)
        _check(self._L.szg_mask_create(self._h, _p(ids, C.c_uint64), _p(passed, C.c_uint8), ids.size, C.byref(mid)))
        return mid.value

    # -- metadata columns and device-side filters (szg_meta_upsert / szg_filter_mask)
    def meta_upsert(self, ids, doc_kinds, cols, values):
        """values[i][j] = (kind, value) of column cols[j] for ids[i] (kinds: syzgydb_b200.filter.MV_*)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        dk = np.ascontiguousarray(doc_kinds, dtype=np.uint8)
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        n, m = ids.size, cols.size
        arr = (MetaValue * max(n * m, 1))()
        keep = []
        for i in range(n):
            for j in range(m):
                kind, v = values[i][j]
                mv = arr[i * m + j]
                mv.kind = kind
                if kind == 4:  # string
                    b = v.encode("utf-8")
                    buf = C.create_string_buffer(b, len(b) + 1)
                    keep.append(buf)
                    mv.str = C.cast(buf, C.c_void_p)
                    mv.str_len = len(b)
                elif kind in (2, 3):  # bool, number
                    mv.num = float(v)
        _check(self._L.szg_meta_upsert(self._h, _p(ids, C.c_uint64), n, _p(dk, C.c_uint8), _p(cols, C.c_uint32), m, arr))

    def filter_mask(self, program) -> int:
        """Runs a postfix filter program (list of dicts: op, arg, num, str, table) on the device; returns a mask id."""
        ops = (FilterOp * max(len(program), 1))()
        keep = []
        for i, o in enumerate(program):
            ops[i].op = o["op"]
            ops[i].arg = o.get("arg", 0)
            ops[i].num = o.get("num", 0.0)
            if "str" in o:
                b = o["str"].encode("utf-8")
                buf = C.create_string_buffer(b, len(b) + 1)
                keep.append(buf)
                ops[i].str = C.cast(buf, C.c_void_p)
                ops[i].str_len = len(b)
            if "table" in o:
                t = bytes(o["table"])
                buf = C.create_string_buffer(t, len(t) + 1)
                keep.append(buf)
                ops[i].table = C.cast(buf, C.c_void_p)
                ops[i].table_len = len(t)
        mid = C.c_int(0)
        _check(self._L.szg_filter_mask(self._h, ops, len(program), C.byref(mid)))
        return mid.value

    def meta_dictionary(self):
        """The string dictionary of the metadata columns, by code."""
        n = C.c_uint32(0)
        _check(self._L.szg_meta_dictionary_size(self._h, C.byref(n)))
        out = []
        for code in range(n.value):
            sp, ln = C.c_void_p(), C.c_uint32(0)
            _check(self._L.szg_meta_dictionary_get(self._h, code, C.byref(sp), C.byref(ln)))
            out.append(C.string_at(sp.value, ln.value).decode("utf-8", "replace") if ln.value else "")
        return out

    def mask_destroy(self, mask_id: int):
        _check(self._L.szg_mask_destroy(self._h, mask_id))

    # -- search (host buffers)
    def search_topk(self, queries, k: int, mask_id: int = -1, flags: int = 0):
        """Returns (ids [nq,k] uint64, dist [nq,k] float64, n [nq] uint32, scanned)."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ValueError(f"query has {q.shape[1]} dimensions, collection has {self.dim}")  # appendix B-12
        nq = q.shape[0]
        kk = max(int(k), 1)
        ids = np.zeros((nq, kk), dtype=np.uint64)
        dist = np.zeros((nq, kk), dtype=np.float64)
        n = np.zeros(nq, dtype=np.uint32)
        scanned = C.c_uint64(0)
        _check(self._L.szg_search_topk(self._h, _p(q, C.c_double), nq, int(k), mask_id, flags, _p(ids, C.c_uint64),
                                       _p(dist, C.c_double), _p(n, C.c_uint32), C.byref(scanned)))
        return ids, dist, n, scanned.value

    def search_batch(self, queries, k: int, mask_id: int = -1, flags: int = 0):
        """Batched form of search_topk (tensor-core contraction for 8-bit collections).  Same return value."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ValueError(f"query has {q.shape[1]} dimensions, collection has {self.dim}")
        nq = q.shape[0]
        kk = max(int(k), 1)
        ids = np.zeros((nq, kk), dtype=np.uint64)
        dist = np.zeros((nq, kk), dtype=np.float64)
        n = np.zeros(nq, dtype=np.uint32)
        scanned = C.c_uint64(0)
        _check(self._L.szg_search_batch(self._h, _p(q, C.c_double), nq, int(k), mask_id, flags, _p(ids, C.c_uint64),
                                        _p(dist, C.c_double), _p(n, C.c_uint32), C.byref(scanned)))
        return ids, dist, n, scanned.value

    def search_radius(self, query, radius: float, mask_id: int = -1, flags: int = 0):
        """Returns (ids, dist, scanned), ascending distance."""
        q = np.ascontiguousarray(query, dtype=np.float64).reshape(-1)
        if q.size != self.dim:
            raise ValueError(f"query has {q.size} dimensions, collection has {self.dim}")
        res = C.c_void_p()
        scanned = C.c_uint64(0)
        _check(self._L.szg_search_radius(self._h, _p(q, C.c_double), float(radius), mask_id, flags, C.byref(res),
                                         C.byref(scanned)))
        try:
            n = C.c_uint64(0)
            _check(self._L.szg_result_count(res, C.byref(n)))
            ids = np.zeros(n.value, dtype=np.uint64)
            dist = np.zeros(n.value, dtype=np.float64)
            if n.value:
                _check(self._L.szg_result_fetch(res, 0, n.value, _p(ids, C.c_uint64), _p(dist, C.c_double)))
        finally:
            self._L.szg_result_free(res)
        return ids, dist, scanned.value

    def search_radius_batch(self, queries, radii, mask_id: int = -1, flags: int = 0):
        """nq radius searches in one call; returns a list of (ids, dist) per query and `scanned`."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        r = np.ascontiguousarray(radii, dtype=np.float64).reshape(-1)
        if r.size != q.shape[0]:
            raise ValueError("one radius per query")
        nq = q.shape[0]
        res = (C.c_void_p * max(nq, 1))()
        scanned = C.c_uint64(0)
        _check(self._L.szg_search_radius_batch(self._h, _p(q, C.c_double), nq, _p(r, C.c_double), mask_id, flags, res,
                                               C.byref(scanned)))
        out = []
        try:
            for i in range(nq):
                n = C.c_uint64(0)
                _check(self._L.szg_result_count(res[i], C.byref(n)))
                ids = np.zeros(n.value, dtype=np.uint64)
                dist = np.zeros(n.value, dtype=np.float64)
                if n.value:
                    _check(self._L.szg_result_fetch(res[i], 0, n.value, _p(ids, C.c_uint64), _p(dist, C.c_double)))
                out.append((ids, dist))
        finally:
            for i in range(nq):
                if res[i]:
                    self._L.szg_result_free(res[i])
        return out, scanned.value

    def rescore_batch(self, queries, id_lists):
        """Candidate lists of several queries in one call (szg_rescore_batch); returns one distance array per list."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        if len(id_lists) != q.shape[0]:
            raise ValueError("one candidate list per query")
        off = np.zeros(len(id_lists) + 1, dtype=np.uint64)
        for i, l in enumerate(id_lists):
            off[i + 1] = off[i] + len(l)
        ids = (np.concatenate([np.asarray(l, dtype=np.uint64) for l in id_lists]) if len(id_lists) else np.zeros(0, np.uint64))
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.zeros(ids.size, dtype=np.float64)
        _check(self._L.szg_rescore_batch(self._h, _p(q, C.c_double), q.shape[0], _p(ids, C.c_uint64), _p(off, C.c_uint64),
                                         _p(out, C.c_double)))
        return [out[int(off[i]):int(off[i + 1])] for i in range(len(id_lists))]

    def rescore(self, query, ids) -> np.ndarray:
        q = np.ascontiguousarray(query, dtype=np.float64).reshape(-1)
        if q.size != self.dim:
            raise ValueError(f"query has {q.size} dimensions, collection has {self.dim}")
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.zeros(ids.size, dtype=np.float64)
        _check(self._L.szg_rescore(self._h, _p(q, C.c_double), _p(ids, C.c_uint64), ids.size, _p(out, C.c_double)))
        return out

    # -- search (device-resident; pointers are raw device addresses, e.g. torch tensor.data_ptr())
    def search_topk_dev(self, d_queries: int, nq: int, k: int, d_out_ids: int, d_out_dist: int, d_out_n: int,
                        stream: int = 0, mask_id: int = -1, flags: int = 0, d_out_flags: int = 0):
        _check(self._L.szg_search_topk_dev(self._h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist,
                                           d_out_n, d_out_flags or None, stream))

    def search_batch_dev(self, d_queries: int, nq: int, k: int, d_out_ids: int, d_out_dist: int, d_out_n: int,
                         stream: int = 0, mask_id: int = -1, flags: int = 0, d_out_flags: int = 0):
        _check(self._L.szg_search_batch_dev(self._h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist,
                                            d_out_n, d_out_flags or None, stream))

    def merge_topk_dev(self, d_g_ids: int, d_g_dist: int, d_g_n: int, nranks: int, nq: int, k: int, d_out_ids: int,
                       d_out_dist: int, d_out_n: int, stream: int = 0, rank_stride_bytes: int = 0, d_g_flags: int = 0,
                       d_out_flags: int = 0):
        _check(self._L.szg_merge_topk_dev(self._h, d_g_ids, d_g_dist, d_g_n, d_g_flags or None, rank_stride_bytes,
                                          nranks, nq, k, d_out_ids, d_out_dist, d_out_n, d_out_flags or None, stream))

    # -- introspection
    def stats(self) -> dict:
        s = Stats()
        _check(self._L.szg_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_ if f != "reserved"}

    def set_option(self, opt: int, value: int):
        _check(self._L.szg_set_option(self._h, opt, value))

    def last_scan_times_ms(self, cap: int = 4096) -> np.ndarray:
        out = np.zeros(cap, dtype=np.float32)
        n = C.c_uint32(0)
        _check(self._L.szg_last_scan_times_ms(self._h, _p(out, C.c_float), cap, C.byref(n)))
        return out[:n.value].copy()


class SpanFile:
    """A collection's .dat file, read directly (include/syzgy_b200.h, "span file -> mirror").  Needs no GPU until load()."""

    def __init__(self, path: str):
        self._L = load()
        self._h = C.c_void_p()
        _check(self._L.szg_spanfile_open(os.fsencode(path), C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.szg_spanfile_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def info(self) -> dict:
        i = SpanFileInfo()
        _check(self._L.szg_spanfile_get_info(self._h, C.byref(i)))
        d = {f: getattr(i, f) for f, _ in SpanFileInfo._fields_}
        d["name"] = d["name"].decode("utf-8", "replace")
        return d

    def ids(self) -> np.ndarray:
        n = C.c_uint64(0)
        _check(self._L.szg_spanfile_ids(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint64)
        _check(self._L.szg_spanfile_ids(self._h, _p(out, C.c_uint64), n.value, C.byref(n)))
        return out

    def record(self, doc_id: int):
        """(vector bytes or None, metadata bytes or None); KeyError like ReadRecord's "record not found"."""
        vp_, mp_ = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
        vl, ml = C.c_uint64(0), C.c_uint64(0)
        rc = self._L.szg_spanfile_record(self._h, int(doc_id), C.byref(vp_), C.byref(vl), C.byref(mp_), C.byref(ml))
        if rc == -4:
            raise KeyError(doc_id)
        _check(rc)
        vec = C.string_at(vp_, vl.value) if vp_ else None
        meta = C.string_at(mp_, ml.value) if mp_ else None
        return vec, meta

    def open_index(self, device: int = 0) -> "Index":
        """szg_create with the header's options + szg_spanfile_load: the GPU mirror of the collection."""
        i = self.info()
        if not i["has_header"]:
            raise SzgError(-1, "span file has no collection header")
        ix = Index(i["dimension_count"], i["quantization"], i["distance_method"], device)
        try:
            self.load_into(ix)
        except Exception:
            ix.close()
            raise
        return ix

    def load_into(self, ix: "Index") -> int:
        n = C.c_uint64(0)
        _check(self._L.szg_spanfile_load(self._h, ix._h, C.byref(n)))
        return n.value
