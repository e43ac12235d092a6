"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product package (treensearch_b200/) never does.

* ``OraclePort``  -> oracle/libtns_oracle.so   (restated criterion, oracle/tns_oracle.c)
* ``Reference``   -> oracle/_ref/libtns_ref.so (the unmodified reference, oracle/ref_shim.cpp)

Both expose the same small surface: add_point_set / set_search_radius / set_active_search /
set_symmetric_search / run(mode) / csr(set_i, set_j) -> (offsets int64[n+1], indices int32[K], ascending lists).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libtns_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libtns_ref.so")

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


def build(quiet: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile (the reference part only if /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


REF_TESTS_BIN = os.path.join(_HERE, "_ref", "ref_tests_on_b200")
REF_STRESS_BIN = os.path.join(_HERE, "_ref", "ref_stress_on_b200")                # the reference's dynamic_emitter_stress_test on the CUDA engine
REF_TESTS_NATIVE_BIN = os.path.join(_HERE, "_ref", "ref_tests_native")      # the same program built with the reference's own library


def build_ref_tests(quiet: bool = True) -> None:
    """Compile the reference's own tests against include/TreeNSearch + libtnsb.so (needs /root/reference and a built library)."""
    subprocess.run(["make", "-C", _HERE, "ref-tests"], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _f32(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None and a.size > 0 else typ()


class _Base:
    _prefix = ""
    _lib = None

    def __init__(self):
        self._keep = []          # borrowed arrays must outlive the handle (the reference stores raw pointers)
        self._n = []
        self.h = C.c_void_p(self._fn("create", C.c_void_p)())

    @classmethod
    def _fn(cls, name, restype=None, *argtypes):
        f = getattr(cls._lib, cls._prefix + name)
        f.restype = restype
        if argtypes:
            f.argtypes = list(argtypes)
        return f

    def close(self):
        if self.h:
            self._fn("destroy", None, C.c_void_p)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_point_set(self, pts, radii=None):
        pts = _f32(pts).reshape(-1, 3)
        radii = _f32(radii)
        self._keep.append((pts, radii))
        self._n.append(pts.shape[0])
        return self._fn("add_point_set", C.c_int, C.c_void_p, _fp, _fp, C.c_int, C.c_int)(
            self.h, _ptr(pts, _fp), _ptr(radii, _fp), pts.shape[0], int(radii is not None))

    def resize_point_set(self, s, pts, radii=None):
        pts = _f32(pts).reshape(-1, 3)
        radii = _f32(radii)
        self._keep.append((pts, radii))
        self._n[s] = pts.shape[0]
        self._fn("resize_point_set", None, C.c_void_p, C.c_int, _fp, _fp, C.c_int, C.c_int)(
            self.h, s, _ptr(pts, _fp), _ptr(radii, _fp), pts.shape[0], int(radii is not None))

    def set_search_radius(self, r):
        self._fn("set_search_radius", None, C.c_void_p, C.c_float)(self.h, float(r))

    def set_active_search(self, i, j, active=True):
        self._fn("set_active_search", None, C.c_void_p, C.c_int, C.c_int, C.c_int)(self.h, i, j, int(active))

    def set_symmetric_search(self, b):
        self._sym = bool(b)
        self._fn("set_symmetric_search", None, C.c_void_p, C.c_int)(self.h, int(b))

    def pair_total(self, si, sj):
        return int(self._fn("pair_total", C.c_int64, C.c_void_p, C.c_int, C.c_int)(self.h, si, sj))


class OraclePort(_Base):
    """Restated oracle.  run(mode): 0 = all pairs, 1 = uniform grid candidates."""
    _prefix = "tnso_"

    def __init__(self):
        if OraclePort._lib is None:
            if not os.path.exists(PORT_SO):
                build()
            OraclePort._lib = C.CDLL(PORT_SO)
        super().__init__()

    def run(self, mode=1):
        rc = self._fn("run", C.c_int, C.c_void_p, C.c_int)(self.h, mode)
        if rc != 0:
            raise RuntimeError(f"oracle port: invalid configuration ({rc})")

    def csr(self, si, sj):
        k = self.pair_total(si, sj)
        off = np.empty(self._n[si] + 1, dtype=np.int64)
        idx = np.empty(max(k, 0), dtype=np.int32)
        rc = self._fn("pair_export", C.c_int, C.c_void_p, C.c_int, C.c_int, _i64p, _i32p)(
            self.h, si, sj, off.ctypes.data_as(_i64p), _ptr(idx, _i32p))
        if rc != 0:
            raise RuntimeError("oracle port: pair not computed")
        return off, idx


def morton3d_64(x, y, z) -> int:
    if OraclePort._lib is None:
        OraclePort()
    f = OraclePort._lib.tnso_morton3d_64
    f.restype = C.c_uint64
    f.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return int(f(int(x), int(y), int(z)))


def list_digests(ids: np.ndarray, pos: np.ndarray, cnt: np.ndarray) -> np.ndarray:
    """Order-independent 64-bit digest of each neighbour list (ids[pos[i] : pos[i]+cnt[i]])."""
    if OraclePort._lib is None:
        OraclePort()
    f = OraclePort._lib.tnso_list_digests
    f.restype = None
    f.argtypes = [_i32p, _i64p, _i32p, C.c_int64, _u64p]
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    out = np.empty(pos.shape[0], dtype=np.uint64)
    f(_ptr(ids, _i32p), _ptr(pos, _i64p), _ptr(cnt, _i32p), pos.shape[0], out.ctypes.data_as(_u64p))
    return out


def csr_digests(off: np.ndarray, idx: np.ndarray) -> np.ndarray:
    cnt = np.diff(off).astype(np.int32)
    return list_digests(idx, off[:-1], cnt)


def reference_available() -> bool:
    return os.path.exists(REF_SO)


class Reference(_Base):
    """The unmodified reference.  run(mode): 0 = run() (AVX2), 1 = run_scalar(), 2 = BruteforceNSearch."""
    _prefix = "tnsref_"

    def __init__(self):
        if Reference._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle` where /root/reference exists")
            Reference._lib = C.CDLL(REF_SO)
        self._sym = True
        super().__init__()

    @classmethod
    def n_threads(cls):
        if cls._lib is None:
            cls._lib = C.CDLL(REF_SO)
        return int(cls._fn("n_threads", C.c_int)())

    def add_point_set_f64(self, pts, radii=None):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        radii = None if radii is None else np.ascontiguousarray(radii, dtype=np.float64)
        self._keep.append((pts, radii))
        self._n.append(pts.shape[0])
        return self._fn("add_point_set_f64", C.c_int, C.c_void_p, _dp, _dp, C.c_int, C.c_int)(
            self.h, _ptr(pts, _dp), _ptr(radii, _dp), pts.shape[0], int(radii is not None))

    def set_n_threads(self, n):
        self._fn("set_n_threads", None, C.c_void_p, C.c_int)(self.h, int(n))

    def run(self, mode=0):
        if mode == 2:
            self._fn("run_bruteforce", C.c_int, C.c_void_p, C.c_int)(self.h, int(self._sym))
        else:
            self._fn("run", C.c_int, C.c_void_p, C.c_int)(self.h, mode)

    def csr(self, si, sj, sort_lists=True):
        k = self.pair_total(si, sj)
        off = np.empty(self._n[si] + 1, dtype=np.int64)
        idx = np.empty(max(k, 0), dtype=np.int32)
        self.n_unsorted_lists = int(self._fn("pair_export", C.c_int64, C.c_void_p, C.c_int, C.c_int, _i64p, _i32p, C.c_int)(
            self.h, si, sj, off.ctypes.data_as(_i64p), _ptr(idx, _i32p), int(sort_lists)))
        return off, idx

    def prepare_zsort(self):
        self._fn("prepare_zsort", None, C.c_void_p)(self.h)

    def zsort_order(self, s):
        out = np.empty(self._n[s], dtype=np.int32)
        self._fn("get_zsort_order", None, C.c_void_p, C.c_int, _i32p)(self.h, s, _ptr(out, _i32p))
        return out

    def apply_zsort(self, s, arr, stride):
        assert arr.dtype == np.float32 and arr.flags.c_contiguous
        self._fn("apply_zsort_f32", None, C.c_void_p, C.c_int, _fp, C.c_int)(self.h, s, _ptr(arr, _fp), stride)

    def neighborlist_n_bytes(self):
        return int(self._fn("neighborlist_n_bytes", C.c_uint64, C.c_void_p)(self.h))

    def time_runs(self, reps, mode=0):
        ms = np.zeros(reps, dtype=np.float64)
        self._fn("time_runs", None, C.c_void_p, C.c_int, C.c_int, _dp)(self.h, mode, reps, ms.ctypes.data_as(_dp))
        return ms
