"""Python mirror of the reference's public class ``tns::TreeNSearch`` (TreeNSearch/source/TreeNSearch.h:28-335).

Same method names, argument meaning and error behaviour as the reference; every call forwards through the C ABI of
libtnsb.so (include/tnsb.h).  Where the reference prints a message and calls ``exit(-1)`` (TreeNSearch.cpp:22-25,
:366-392) this mirror raises :class:`TreeNSearchError` carrying the same message.

Point / radii arrays are *borrowed*, like in the reference (TreeNSearch.cpp:35-41): they are re-read by every ``run()``,
so in-place updates of the numpy array / torch tensor are seen by the next run.  Arrays may be numpy (host) arrays or
torch tensors on the CPU (pinned or not) or on the context's GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class TreeNSearchError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _as_buffer(a, dtypes, what):
    """Returns (pointer, n_elements, is_f64, keepalive) of a contiguous float32/float64 numpy array or torch tensor."""
    if a is None:
        return None, 0, False, None
    if _is_torch(a):
        import torch
        if a.dtype not in (torch.float32, torch.float64):
            raise TypeError(f"{what}: torch tensor must be float32 or float64")
        if not a.is_contiguous():
            raise ValueError(f"{what}: tensor must be contiguous (the engine borrows the memory)")
        return a.data_ptr(), a.numel(), a.dtype == torch.float64, a
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        raise TypeError(f"{what}: array must be float32 or float64 (got {a.dtype})")
    if not a.flags.c_contiguous:
        raise ValueError(f"{what}: array must be C-contiguous (the engine borrows the memory)")
    return a.ctypes.data, a.size, a.dtype == np.float64, a


class _EngineView(np.ndarray):
    """numpy view of engine-owned memory that holds a reference to its engine (so the buffers outlive the view)."""
    _owner = None

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class NeighborList:
    """Handle to one neighbour list (reference: TreeNSearch/source/NeighborList.h:8-39)."""
    __slots__ = ("_view",)

    def __init__(self, view):
        self._view = view            # numpy int32 view of [j0 .. j(n-1)]

    def size(self):
        return int(self._view.shape[0])

    def __len__(self):
        return int(self._view.shape[0])

    def __getitem__(self, i):
        return int(self._view[i])

    def get_ptr(self):
        return self._view

    def __iter__(self):
        return iter(self._view.tolist())


class TreeNSearch:
    def __init__(self, device: int = -1):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.tnsb_create(C.byref(h), int(device))
        if rc != L.TNSB_OK:
            raise TreeNSearchError(rc, self._lib.tnsb_last_error(None).decode())
        self._h = h
        self._keep = {}              # set id -> (points, radii) keep-alive of the borrowed arrays
        self._n_threads = -1
        self._views = {}
        self._query_limit = -1

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.tnsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise TreeNSearchError(rc, self._lib.tnsb_last_error(self._h).decode())
        return rc

    # ------------------------------------------------------------------ main interface (TreeNSearch.h:50-225)
    def add_point_set(self, points, radii=None, n_points=None, variable_radius=None):
        """add_point_set(points[, radii]) -> set id.  float32 or float64, xyzxyz layout, shape (n,3) or (3n,)."""
        if variable_radius is None:
            variable_radius = radii is not None
        p, cnt, f64, keep_p = _as_buffer(points, None, "points")
        r, cnt_r, f64_r, keep_r = _as_buffer(radii, None, "radii")
        n = cnt // 3 if n_points is None else int(n_points)
        if radii is not None and f64_r != f64:
            raise TypeError("points and radii must have the same dtype")
        if radii is not None and cnt_r < n:
            raise ValueError("radii array shorter than the point set")
        fn = self._lib.tnsb_add_point_set_f64 if f64 else self._lib.tnsb_add_point_set_f32
        s = self._check(fn(self._h, p, r, n, int(bool(variable_radius))))
        self._keep[s] = (keep_p, keep_r)
        return s

    def resize_point_set(self, set_id, points, radii=None, n_points=None, variable_radius=None):
        if variable_radius is None:
            variable_radius = radii is not None
        p, cnt, f64, keep_p = _as_buffer(points, None, "points")
        r, cnt_r, f64_r, keep_r = _as_buffer(radii, None, "radii")
        n = cnt // 3 if n_points is None else int(n_points)
        fn = self._lib.tnsb_resize_point_set_f64 if f64 else self._lib.tnsb_resize_point_set_f32
        self._check(fn(self._h, int(set_id), p, r, n, int(bool(variable_radius))))
        old = self._keep.get(set_id, (None, None))
        self._keep[set_id] = (keep_p, keep_r if variable_radius else old[1])

    def set_search_radius(self, search_radius):
        self._check(self._lib.tnsb_set_search_radius(self._h, float(np.float32(search_radius))))

    def set_cell_size(self, cell_size):
        self._check(self._lib.tnsb_set_cell_size(self._h, float(np.float32(cell_size))))

    def run(self):
        self._views.clear()
        self._check(self._lib.tnsb_run(self._h))

    def run_scalar(self):
        """The reference's scalar twin of run() (TreeNSearch.cpp:150-160); here the same CUDA path."""
        self.run()

    def _pair(self, set_i, set_j):
        key = (set_i, set_j)
        v = self._views.get(key)
        if v is None:
            rag = C.POINTER(C.c_int32)()
            pos = C.POINTER(C.c_int64)()
            n_ints = C.c_int64()
            self._check(self._lib.tnsb_get_neighborlists(self._h, set_i, set_j, C.byref(rag), C.byref(pos), C.byref(n_ints)))
            n_i = self.get_n_points_in_set(set_i)
            if self._query_limit >= 0:
                n_i = min(n_i, self._query_limit)      # find-only (halo) points have no list
            ragged = np.ctypeslib.as_array(rag, shape=(max(n_ints.value, 1),))[: n_ints.value] if n_ints.value > 0 else np.zeros(0, np.int32)
            list_pos = np.ctypeslib.as_array(pos, shape=(max(n_i, 1),))[:n_i] if n_i > 0 else np.zeros(0, np.int64)
            # zero-copy views of engine-owned pinned memory: they keep the engine alive, are read-only, and -- exactly like the
            # reference's NeighborList pointers (TreeNSearch.cpp:393) -- are invalidated by the next run() / close()
            ragged, list_pos = ragged.view(_EngineView), list_pos.view(_EngineView)
            ragged._owner = list_pos._owner = self
            ragged.flags.writeable = False
            list_pos.flags.writeable = False
            v = (ragged, list_pos)
            self._views[key] = v
        return v

    def get_neighborlist(self, set_i, set_j, point_i) -> NeighborList:
        ragged, list_pos = self._pair(set_i, set_j)
        p = int(list_pos[point_i])
        n = int(ragged[p])
        return NeighborList(ragged[p + 1: p + 1 + n])

    def for_each_neighbor(self, set_i, set_j, i, f):
        for j in self.get_neighborlist(set_i, set_j, i):
            f(j)

    def prepare_zsort(self):
        self._check(self._lib.tnsb_prepare_zsort(self._h))

    def get_zsort_order(self, set_i) -> np.ndarray:
        ptr = C.POINTER(C.c_int32)()
        n = C.c_int()
        self._check(self._lib.tnsb_get_zsort_order(self._h, int(set_i), C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def apply_zsort(self, set_i, data, stride=1):
        """In-place gather data[new] = data[old] (TreeNSearch.h:443-481).  numpy arrays are gathered on the host;
        torch CUDA tensors (any dtype whose rows are a multiple of 4 bytes) on the device."""
        if _is_torch(data) and data.is_cuda:
            self.apply_zsort_device(set_i, [data], [stride])
            return
        order = self.get_zsort_order(set_i)
        if _is_torch(data):
            data = data.numpy()
        n = self.get_n_points_in_set(set_i)
        flat = data.reshape(-1)
        rows = flat[: n * stride].reshape(n, stride)
        rows[...] = rows[order.astype(np.int64)]

    def apply_zsort_device(self, set_i, arrays, strides=None, out=None):
        """Fused device gather of several per-point arrays of set_i in ONE launch (positions, velocities, ...):
        arrays[k][new] = arrays[k][old] in place, or into out[k] when given.  arrays: contiguous torch CUDA tensors;
        strides[k] = elements per point (default: inferred from the tensor shape)."""
        n = self.get_n_points_in_set(set_i)
        k = len(arrays)
        src = (C.c_void_p * k)()
        dst = (C.c_void_p * k)()
        row = (C.c_int * k)()
        for i, a in enumerate(arrays):
            if not (_is_torch(a) and a.is_cuda and a.is_contiguous()):
                raise TypeError("apply_zsort_device needs contiguous torch CUDA tensors")
            stride = int(strides[i]) if strides is not None else (a.numel() // max(n, 1))
            if a.numel() < n * stride:
                raise ValueError("array shorter than n_points * stride")
            src[i] = a.data_ptr()
            o = a if out is None else out[i]
            if o is not a and not (_is_torch(o) and o.is_cuda and o.is_contiguous() and o.numel() >= n * stride and o.element_size() == a.element_size()):
                raise TypeError("out[k] must be a contiguous CUDA tensor shaped like arrays[k]")
            dst[i] = o.data_ptr()
            row[i] = stride * a.element_size()
        self._check(self._lib.tnsb_apply_zsort_device(self._h, int(set_i), k, src, dst, row))

    def set_symmetric_search(self, activate):
        self._check(self._lib.tnsb_set_symmetric_search(self._h, int(bool(activate))))

    # ------------------------------------------------------------------ secondary methods (TreeNSearch.h:233-246)
    def print_state(self):
        s = self.stats()
        print("\n ================ OPTIONS ================ ")
        print(f"n_threads: {self._n_threads}")
        print("\n ================ GRID ================ ")
        print("World AABB float")
        print(list(s["domain_bottom"]))
        print(list(s["domain_top"]))
        print(f"cell_size: {s['cell_size']}")
        print(f"# cells: {s['n_cells']}")
        print("\n ================ NEIGHBORLISTS ================ ")
        print("Active searches: ")
        ns = self.get_n_sets()
        for i in range(ns):
            for j in range(ns):
                if self.is_search_active(i, j):
                    print(f"\tset_{i} -> set_{j}")
        print(f"Total memory (MB): {self.get_neighborlist_n_bytes() / 1024.0 / 1024.0}")
        print("\n ================ PER SET DATA ================ ")
        for i in range(ns):
            print(f"\n ---------------- set_{i} ---------------- ")
            print(f"# points: {self.get_n_points_in_set(i)}")
            for j in range(ns):
                if self.is_search_active(i, j):
                    out = (C.c_int64 * 3)()
                    if self._lib.tnsb_get_pair_neighbor_stats(self._h, i, j, out) == 0:
                        n = max(self.get_n_points_in_set(i), 1)
                        print(f"n_neighbors set_{i} -> set_{j} [min, max, avg]: [{out[0]}, {out[1]}, {out[2] / n}]")

    def get_neighborlist_n_bytes(self):
        return int(self._lib.tnsb_get_neighborlist_n_bytes(self._h))

    # ------------------------------------------------------------------ setters and getters (TreeNSearch.h:256-334)
    def set_all_searches(self, active):
        self._check(self._lib.tnsb_set_all_searches(self._h, int(bool(active))))

    def set_active_search(self, set_i, set_j_or_search=True, active_or_find=True):
        """Both reference overloads: (set_i, set_j, active=True) and (set_i, search_in_all=True, be_found_by_all=True);
        as in C++, a bool second argument selects the second overload."""
        if isinstance(set_j_or_search, (bool, np.bool_)):
            self._check(self._lib.tnsb_set_active_search_of_set(self._h, int(set_i), int(set_j_or_search), int(bool(active_or_find))))
        else:
            self._check(self._lib.tnsb_set_active_search(self._h, int(set_i), int(set_j_or_search), int(bool(active_or_find))))

    def set_n_threads(self, n_threads):
        self._n_threads = int(n_threads)        # host-side hint only; the search runs on the GPU

    def set_recursion_cap(self, cap):
        if cap <= 0:                             # TreeNSearch.cpp:372-375 (checked at run() there)
            raise TreeNSearchError(L.TNSB_ERR_INVALID_STATE, "TreeNSearch error: n_points_to_stop_recursion <= 0.")

    def set_n_points_for_parallel_octree(self, n_points=200000):
        pass                                     # no octree in this engine

    def get_n_sets(self):
        return int(self._lib.tnsb_get_n_sets(self._h))

    def get_n_threads(self):
        return self._n_threads

    def get_n_points_in_set(self, set_i):
        return int(self._lib.tnsb_get_n_points_in_set(self._h, int(set_i)))

    def get_total_n_points(self):
        return int(self._lib.tnsb_get_total_n_points(self._h))

    def is_search_active(self, set_i, set_j):
        return bool(self._lib.tnsb_is_search_active(self._h, int(set_i), int(set_j)))

    def does_set_exist(self, set_i):
        return bool(self._lib.tnsb_does_set_exist(self._h, int(set_i)))

    # ------------------------------------------------------------------ engine extras (not in the reference)
    def set_option(self, option, value):
        self._check(self._lib.tnsb_set_option(self._h, int(option), int(value)))
        if int(option) == L.TNSB_OPT_QUERY_LIMIT:
            self._query_limit = int(value)

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (an int cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream; 0 is the legacy
        default stream).  None restores the context's own stream."""
        handle = C.c_void_p(-1) if cuda_stream is None else C.c_void_p(int(cuda_stream))
        self._check(self._lib.tnsb_set_stream(self._h, handle))

    def stats(self) -> dict:
        st = L.Stats()
        self._check(self._lib.tnsb_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def neighbor_lists(self, set_i, set_j):
        """Bulk host view: (ragged int32[n_ints], list_pos int64[n_i]); list i = ragged[list_pos[i]] ids starting at list_pos[i]+1."""
        return self._pair(set_i, set_j)

    def neighbor_csr(self, set_i, set_j, sort_lists=True):
        """Canonical CSR (offsets int64[n+1], indices int32[K]) in point order, each list ascending if sort_lists."""
        ragged, list_pos = self._pair(set_i, set_j)
        n = list_pos.shape[0]
        cnt = ragged[list_pos].astype(np.int64) if n else np.zeros(0, np.int64)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(cnt, out=off[1:])
        total = int(off[-1])
        # gather: for every output slot, its source position in the ragged buffer
        src = np.repeat(list_pos + 1 - off[:-1], cnt) + np.arange(total, dtype=np.int64)
        idx = ragged[src]
        if sort_lists and total:
            owner = np.repeat(np.arange(n, dtype=np.int64), cnt)
            order = np.lexsort((idx, owner))
            idx = idx[order]
        return off, np.ascontiguousarray(idx, dtype=np.int32)

    def neighbor_lists_device(self, set_i, set_j):
        """(d_ragged_ptr, d_list_pos_ptr, n_ints) raw device pointers of the last run."""
        rag = C.POINTER(C.c_int32)()
        pos = C.POINTER(C.c_int64)()
        n_ints = C.c_int64()
        self._check(self._lib.tnsb_get_neighborlists_device(self._h, set_i, set_j, C.byref(rag), C.byref(pos), C.byref(n_ints)))
        return C.cast(rag, C.c_void_p).value, C.cast(pos, C.c_void_p).value, n_ints.value
