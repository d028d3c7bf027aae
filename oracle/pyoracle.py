"""ctypes binding of the CPU oracle (oracle/syzgy_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (syzgydb_b200/) never imports it.
Parity pinning status: see the header of syzgy_oracle.c ("parity unpinned" except the
euclidean KAT and the behavioural properties of the reference's own tests).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsyzgy_oracle.so")

EUCLIDEAN = 0
COSINE = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "syzgy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u64p, i64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_double))
        L.orc_quantize.restype = C.c_uint64
        L.orc_quantize.argtypes = [C.c_double, C.c_int]
        L.orc_dequantize.restype = C.c_double
        L.orc_dequantize.argtypes = [C.c_uint64, C.c_int]
        L.orc_vector_size.restype = C.c_int64
        L.orc_vector_size.argtypes = [C.c_int, C.c_int64]
        L.orc_encode.argtypes = [f64p, C.c_int64, C.c_int, u8p]
        L.orc_decode.argtypes = [u8p, C.c_int64, C.c_int, f64p]
        L.orc_euclidean.restype = C.c_double
        L.orc_euclidean.argtypes = [f64p, f64p, C.c_int64]
        L.orc_angular.restype = C.c_double
        for fn in (L.orc_go_acos, L.orc_go_asin, L.orc_go_atan):
            fn.restype = C.c_double
            fn.argtypes = [C.c_double]
        L.orc_libm_acos_mode.argtypes = [C.c_int]
        L.orc_angular.argtypes = [f64p, f64p, C.c_int64]
        L.orc_lex_order.argtypes = [u64p, C.c_int64, i64p]
        L.orc_search_exact.restype = C.c_int64
        L.orc_search_exact.argtypes = [u8p, u64p, C.c_int64, C.c_int64, C.c_int, C.c_int, f64p, C.c_int64,
                                       C.c_double, u8p, i64p, C.c_int, u64p, f64p, C.c_int64, f64p]
        L.orc_spans_build.restype = C.c_void_p
        L.orc_spans_build.argtypes = [u8p, u64p, C.c_int64, C.c_int64, C.c_int64]
        L.orc_spans_free.argtypes = [C.c_void_p]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [u8p, C.c_int64]
        L.orc_search_exact_spans.restype = C.c_int64
        L.orc_search_exact_spans.argtypes = [C.c_void_p, u8p, u64p, C.c_int64, C.c_int64, C.c_int, C.c_int, f64p, C.c_int64,
                                             C.c_double, u8p, i64p, u64p, f64p, C.c_int64, f64p]
        L.orc_replay.restype = C.c_int64
        L.orc_replay.argtypes = [u8p, u64p, C.c_int64, C.c_int64, C.c_int, C.c_int, f64p, C.c_int64,
                                 C.c_double, u8p, i64p, C.c_int64, u64p, f64p, C.c_int64, i64p]
        L.orc_row_distances.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, f64p, i64p, C.c_int64, f64p]
        L.orc_rand_u64.restype = C.c_uint64
        L.orc_rand_u64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_synth_rows.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int, u8p]
        L.orc_synth_queries.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, f64p]
        L.orc_lsh_new.restype = C.c_void_p
        L.orc_lsh_new.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.orc_lsh_add.argtypes = [C.c_void_p, C.c_int64, f64p]
        L.orc_lsh_free.argtypes = [C.c_void_p]
        L.orc_search_lsh.restype = C.c_int64
        L.orc_search_lsh.argtypes = [C.c_void_p, u64p, C.c_int64, f64p, C.c_int64, C.c_double, u8p, u64p,
                                     f64p, C.c_int64, f64p, i64p, C.c_int64, i64p]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def quantize(v: float, bits: int) -> int:
    return int(lib().orc_quantize(float(v), bits))


def dequantize(q: int, bits: int) -> float:
    return float(lib().orc_dequantize(int(q), bits))


def vector_size(bits: int, dims: int) -> int:
    return int(lib().orc_vector_size(bits, dims))


def encode(vec, bits: int) -> np.ndarray:
    v = np.ascontiguousarray(vec, dtype=np.float64)
    out = np.zeros(vector_size(bits, v.size), dtype=np.uint8)
    lib().orc_encode(_p(v, C.c_double), v.size, bits, _p(out, C.c_uint8))
    return out


def encode_rows(vecs, bits: int) -> np.ndarray:
    vecs = np.ascontiguousarray(vecs, dtype=np.float64)
    return np.stack([encode(v, bits) for v in vecs]) if len(vecs) else np.zeros(
        (0, vector_size(bits, vecs.shape[1])), np.uint8)


def decode(data, dims: int, bits: int) -> np.ndarray:
    d = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(dims, dtype=np.float64)
    lib().orc_decode(_p(d, C.c_uint8), dims, bits, _p(out, C.c_double))
    return out


def euclidean(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().orc_euclidean(_p(a, C.c_double), _p(b, C.c_double), a.size))


def angular(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().orc_angular(_p(a, C.c_double), _p(b, C.c_double), a.size))


def lex_order(ids) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    perm = np.zeros(ids.size, dtype=np.int64)
    lib().orc_lex_order(_p(ids, C.c_uint64), ids.size, _p(perm, C.c_int64))
    return perm


class Spans:
    """In-memory span-file image of a row-major code matrix (orc_spans_build): what the faithful CPU variant of the scan
    reads every record from -- decimal-string key, index lookup, parseSpan, CRC-32 over the whole span."""

    def __init__(self, codes, ids, meta_len=24):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.ids = np.ascontiguousarray(ids, dtype=np.uint64)
        rb = self.codes.size // max(self.ids.size, 1)
        self._h = lib().orc_spans_build(_p(self.codes, C.c_uint8), _p(self.ids, C.c_uint64), self.ids.size, rb, meta_len)

    def close(self):
        if self._h:
            lib().orc_spans_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def crc32(data: bytes) -> int:
    d = np.frombuffer(data, dtype=np.uint8)
    return int(lib().orc_crc32(_p(np.ascontiguousarray(d), C.c_uint8), d.size))


def search_exact(codes, ids, dims, bits, metric, query, k=0, radius=0.0, passmask=None, order="lex",
                 faithful=False, out_cap=None, spans=None):
    """Search(Precision="exact").  Returns (ids, dists, percent_searched).  spans: a Spans image of the same rows -> every
    record goes through the restated getDocument (the faithful CPU variant)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    n = ids.size
    q = np.ascontiguousarray(query, dtype=np.float64)
    if isinstance(order, str):
        order = lex_order(ids) if order == "lex" else None
    elif order is not None:
        order = np.ascontiguousarray(order, dtype=np.int64)
    pm = None if passmask is None else np.ascontiguousarray(passmask, dtype=np.uint8)
    cap = out_cap if out_cap is not None else (n if radius > 0 else max(int(k), 0))
    cap = max(cap, 1)
    oi = np.zeros(cap, dtype=np.uint64)
    od = np.zeros(cap, dtype=np.float64)
    pct = C.c_double(0)
    if spans is not None:
        m = lib().orc_search_exact_spans(spans._h, _p(codes, C.c_uint8), _p(ids, C.c_uint64), n, dims, bits, metric,
                                         _p(q, C.c_double), int(k), float(radius), _p(pm, C.c_uint8), _p(order, C.c_int64),
                                         _p(oi, C.c_uint64), _p(od, C.c_double), cap, C.byref(pct))
    else:
        m = lib().orc_search_exact(_p(codes, C.c_uint8), _p(ids, C.c_uint64), n, dims, bits, metric,
                                   _p(q, C.c_double), int(k), float(radius), _p(pm, C.c_uint8),
                                   _p(order, C.c_int64), int(faithful), _p(oi, C.c_uint64), _p(od, C.c_double),
                                   cap, C.byref(pct))
    m = min(int(m), cap)
    return oi[:m].copy(), od[:m].copy(), pct.value


def replay(codes, ids, dims, bits, metric, query, visit, k=0, radius=0.0, passmask=None):
    """`consider` replayed over an explicit visit sequence.  Returns (ids, dists, points_searched)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    q = np.ascontiguousarray(query, dtype=np.float64)
    visit = np.ascontiguousarray(visit, dtype=np.int64)
    pm = None if passmask is None else np.ascontiguousarray(passmask, dtype=np.uint8)
    cap = max(visit.size, 1)
    oi = np.zeros(cap, dtype=np.uint64)
    od = np.zeros(cap, dtype=np.float64)
    ps = C.c_int64(0)
    m = lib().orc_replay(_p(codes, C.c_uint8), _p(ids, C.c_uint64), ids.size, dims, bits, metric,
                         _p(q, C.c_double), int(k), float(radius), _p(pm, C.c_uint8), _p(visit, C.c_int64),
                         visit.size, _p(oi, C.c_uint64), _p(od, C.c_double), cap, C.byref(ps))
    return oi[:m].copy(), od[:m].copy(), ps.value


def row_distances(codes, dims, bits, metric, query, rows) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    q = np.ascontiguousarray(query, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.zeros(rows.size, dtype=np.float64)
    lib().orc_row_distances(_p(codes, C.c_uint8), dims, bits, metric, _p(q, C.c_double),
                            _p(rows, C.c_int64), rows.size, _p(out, C.c_double))
    return out


def synth_rows(seed: int, row0: int, nrows: int, dims: int, bits: int) -> np.ndarray:
    rb = vector_size(bits, dims)
    out = np.zeros((nrows, rb), dtype=np.uint8)
    lib().orc_synth_rows(seed, row0, nrows, dims, bits, _p(out, C.c_uint8))
    return out


def synth_queries(seed: int, q0: int, nq: int, dims: int) -> np.ndarray:
    out = np.zeros((nq, dims), dtype=np.float64)
    lib().orc_synth_queries(seed, q0, nq, dims, _p(out, C.c_double))
    return out


class LshTree:
    """Restated lshTree (lshtree.go) over a row-major code matrix."""

    def __init__(self, codes, dims, bits, metric, threshold=100, ntrees=5, seed=1):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)  # kept alive: the C side borrows it
        self.dims, self.bits, self.metric = dims, bits, metric
        self._t = lib().orc_lsh_new(_p(self.codes, C.c_uint8), dims, bits, metric, threshold, ntrees, seed)

    def add(self, row: int, vec):
        v = np.ascontiguousarray(vec, dtype=np.float64)
        lib().orc_lsh_add(self._t, row, _p(v, C.c_double))

    def add_all_decoded(self, nrows: int):
        """NewCollection's reload loop (collection.go:298-311): decoded vectors, rows in order."""
        for r in range(nrows):
            self.add(r, decode(self.codes[r], self.dims, self.bits))

    def search(self, ids, query, k=0, radius=0.0, passmask=None):
        """Returns (ids, dists, percent_searched, visit_rows)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        n = ids.size
        q = np.ascontiguousarray(query, dtype=np.float64)
        pm = None if passmask is None else np.ascontiguousarray(passmask, dtype=np.uint8)
        cap = max(n, 1)
        oi = np.zeros(cap, dtype=np.uint64)
        od = np.zeros(cap, dtype=np.float64)
        visit = np.zeros(cap, dtype=np.int64)
        pct = C.c_double(0)
        nv = C.c_int64(0)
        m = lib().orc_search_lsh(self._t, _p(ids, C.c_uint64), n, _p(q, C.c_double), int(k), float(radius),
                                 _p(pm, C.c_uint8), _p(oi, C.c_uint64), _p(od, C.c_double), cap,
                                 C.byref(pct), _p(visit, C.c_int64), cap, C.byref(nv))
        return oi[:m].copy(), od[:m].copy(), pct.value, visit[:nv.value].copy()

    def close(self):
        if self._t:
            lib().orc_lsh_free(self._t)
            self._t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def go_acos(x: float) -> float:
    """Go's math.Acos restated (oracle/syzgy_oracle.c orc_go_acos)."""
    return float(lib().orc_go_acos(float(x)))


def go_atan(x: float) -> float:
    return float(lib().orc_go_atan(float(x)))


def libm_acos_mode(on: bool):
    """Switches the oracle's angular distance between Go's Acos restatement (default) and libm acos."""
    lib().orc_libm_acos_mode(1 if on else 0)
