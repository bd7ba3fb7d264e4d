"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle.

`Oracle()` loads oracle/liboracle.so (our C restatement, oracle/hist_oracle.c + graph_oracle.c).
`Reference()` loads oracle/_ref/libscema_ref.so (the unmodified reference header behind
oracle/ref_harness.cc); it raises FileNotFoundError when oracle/_ref has not been built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (scema_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile oracle/liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []))


class _PairsMixin:
    """Shared wrappers: both libraries export the same shapes under different prefixes."""

    _prefix = ""

    def _bind_common(self):
        p = self._prefix
        L = self.lib
        f = getattr(L, p + "splinify_batch")
        f.argtypes = [_f64p, _u64p, C.c_uint64, C.c_uint32, _f64p, C.c_int]
        f = getattr(L, p + "compare_l2")
        f.restype = C.c_double
        f.argtypes = [_f64p, _f64p, C.c_uint32]
        f = getattr(L, p + "all_pairs")
        f.restype = C.c_uint64
        f.argtypes = [_f64p, C.c_uint64, C.c_uint32, C.c_double, C.c_uint64, C.c_uint64, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        f = getattr(L, p + "format_double")
        f.restype = C.c_int
        f.argtypes = [C.c_double, C.c_char_p, C.c_int]
        f = getattr(L, p + "max_threads")
        f.restype = C.c_int

    def max_threads(self):
        return int(getattr(self.lib, self._prefix + "max_threads")())

    def splinify_batch(self, steps, offsets, P, nthreads=0):
        steps = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, 6)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.empty((n, 6 * P), dtype=np.float64)
        rc = getattr(self.lib, self._prefix + "splinify_batch")(steps, offsets, n, P, out, nthreads)
        if rc:
            raise ValueError("splinify: a history has fewer than 3 steps (reference exits)")
        return out

    def splinify(self, steps, P):
        steps = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, 6)
        return self.splinify_batch(steps, np.array([0, len(steps)], dtype=np.uint64), P, 1)[0]

    def compare_l2(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert a.shape == b.shape
        return float(getattr(self.lib, self._prefix + "compare_l2")(a, b, a.size))

    def all_pairs(self, rows, thr, row_begin=0, row_end=None, nthreads=0, count_only=False):
        """-> (ei, ej, ed, pairs) sorted by (i,j); with count_only -> (n_edges, pairs)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        n, k = rows.shape
        row_end = n if row_end is None else row_end
        pairs = C.c_uint64(0)
        fn = getattr(self.lib, self._prefix + "all_pairs")
        found = fn(rows, n, k, thr, row_begin, row_end, nthreads, None, None, None, 0, C.byref(pairs))
        if count_only:
            return int(found), int(pairs.value)
        ei = np.empty(found, dtype=np.uint32)
        ej = np.empty(found, dtype=np.uint32)
        ed = np.empty(found, dtype=np.float64)
        if found:
            fn(rows, n, k, thr, row_begin, row_end, nthreads, ei.ctypes.data, ej.ctypes.data,
               ed.ctypes.data, found, C.byref(pairs))
        return ei, ej, ed, int(pairs.value)

    def format_double(self, v):
        buf = C.create_string_buffer(64)
        getattr(self.lib, self._prefix + "format_double")(float(v), buf, 64)
        return buf.value.decode()


class Oracle(_PairsMixin):
    _prefix = "oracle_"

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self._bind_common()
        L = self.lib
        L.oracle_splinify_batch.restype = C.c_int
        L.oracle_check_edges.restype = C.c_uint64
        L.oracle_check_edges.argtypes = [_f64p, C.c_uint32, C.c_double, _u32p, _u32p, _f64p,
                                         C.c_uint64, C.c_int]
        L.oracle_write_similar_files.restype = C.c_int
        L.oracle_write_similar_files.argtypes = [_u32p, C.c_uint64, _u32p, _u32p, _f64p, C.c_uint64,
                                                 C.c_char_p]
        L.oracle_reduce_graph.restype = C.c_int
        L.oracle_reduce_graph.argtypes = [_u32p, _u32p, C.c_void_p, C.c_uint64, C.c_uint32, _u32p,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]

    def check_edges(self, rows, thr, ei, ej, ed, nthreads=0):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        return int(self.lib.oracle_check_edges(rows, rows.shape[1], thr,
                                               np.ascontiguousarray(ei, dtype=np.uint32),
                                               np.ascontiguousarray(ej, dtype=np.uint32),
                                               np.ascontiguousarray(ed, dtype=np.float64),
                                               len(ei), nthreads))

    def write_similar_files(self, ids, ei, ej, ed, pattern):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        return int(self.lib.oracle_write_similar_files(
            ids, len(ids), np.ascontiguousarray(ei, dtype=np.uint32),
            np.ascontiguousarray(ej, dtype=np.uint32), np.ascontiguousarray(ed, dtype=np.float64),
            len(ei), pattern.encode()))

    def reduce_graph(self, eu, ev, num_gps, dist=None):
        """add_edge call sequence (eu[k], ev[k]) -> (mapping, iterations, neighbours_removed)."""
        eu = np.ascontiguousarray(eu, dtype=np.uint32)
        ev = np.ascontiguousarray(ev, dtype=np.uint32)
        mapping = np.empty(num_gps, dtype=np.uint32)
        it = C.c_uint64(0)
        nr = C.c_uint64(0)
        dptr = None
        if dist is not None:
            dist = np.ascontiguousarray(dist, dtype=np.float64)
            dptr = dist.ctypes.data
        rc = self.lib.oracle_reduce_graph(eu, ev, dptr, len(eu), num_gps, mapping, C.byref(it),
                                          C.byref(nr))
        if rc:
            raise ValueError({2: "IndexError: ID >= num_gps", 3: "ZeroDivisionError: dist == 0"}[rc])
        return mapping, int(it.value), int(nr.value)


class Reference(_PairsMixin):
    """The unmodified reference header (oracle/_ref/libscema_ref.so)."""

    _prefix = "ref_"

    def __init__(self, o0=False):
        path = os.path.join(REF_DIR, "libscema_ref_O0.so" if o0 else "libscema_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._bind_common()
        self.lib.ref_splinify_batch.restype = None
        self.lib.ref_pipeline.restype = None
        self.lib.ref_pipeline.argtypes = [_f64p, _u64p, _u32p, C.c_uint64, C.c_uint32, C.c_double,
                                          C.c_char_p, C.c_void_p]
        self.lib.ref_from_file.restype = C.c_uint64
        self.lib.ref_from_file.argtypes = [C.c_char_p, _f64p, C.c_uint64]

    def from_file(self, path, cap_steps=1 << 16):
        """Strain6D::from_file on one file -> [L, 6] array of the steps it stored."""
        buf = np.empty((cap_steps, 6), dtype=np.float64)
        n = int(self.lib.ref_from_file(os.fsencode(path), buf, cap_steps))
        assert n <= cap_steps
        return buf[:n].copy()

    def pipeline(self, steps, offsets, ids, P, thr, pattern=None, want_spline=False):
        steps = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, 6)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        n = len(ids)
        sp = np.empty((n, 6 * P), dtype=np.float64) if want_spline else None
        self.lib.ref_pipeline(steps, offsets, ids, n, P, thr,
                              pattern.encode() if pattern else None,
                              sp.ctypes.data if want_spline else None)
        return sp


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "libscema_ref.so"))


def ref_binary(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None
