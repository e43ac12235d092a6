"""ctypes binding of the C ABI in include/tnsb.h (libtnsb.so, built in-tree by build.py / __graft_entry__.build()).

There is no fallback: if the shared library is missing the import of the symbols fails loudly, and if no B200 is visible
tnsb_create() returns TNSB_ERR_NO_DEVICE which the Python mirror turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TNSB_LIB", os.path.join(_HERE, "libtnsb.so"))     # TNSB_LIB: experiment builds only

TNSB_OK = 0
TNSB_ERR_INVALID_ARGUMENT = -1
TNSB_ERR_INVALID_STATE = -2
TNSB_ERR_CUDA = -3
TNSB_ERR_LIMIT = -4
TNSB_ERR_NO_DEVICE = -5

TNSB_OPT_HOST_RESULTS = 1
TNSB_OPT_PIN_USER_MEMORY = 2
TNSB_OPT_LIST_CAPACITY = 3
TNSB_OPT_QUERY_LIMIT = 4
TNSB_OPT_SORT_LISTS = 5
TNSB_OPT_POINT_STRIDE = 6
TNSB_OPT_ZERO_COPY_RESULTS = 7
TNSB_OPT_QUERY_KERNEL = 8
TNSB_OPT_BUILD = 9


class Stats(C.Structure):
    """Mirror of `tnsb_stats` (include/tnsb.h)."""
    _fields_ = [
        ("ms_total_device", C.c_double), ("ms_upload", C.c_double), ("ms_aabb", C.c_double), ("ms_keys", C.c_double),
        ("ms_sort", C.c_double), ("ms_reorder", C.c_double), ("ms_cells", C.c_double), ("ms_query", C.c_double),
        ("ms_download", C.c_double), ("ms_wall", C.c_double),
        ("n_points_total", C.c_int64), ("n_queries", C.c_int64), ("n_neighbors", C.c_int64), ("n_list_ints", C.c_int64),
        ("n_cells", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("n_kernel_launches", C.c_int32), ("n_query_launches", C.c_int32), ("key_bits", C.c_int32), ("sort_passes", C.c_int32),
        ("n_reruns", C.c_int32), ("cell_size", C.c_float), ("domain_bottom", C.c_float * 3), ("domain_top", C.c_float * 3),
        ("brick_query", C.c_int32), ("n_slow_queries", C.c_int64), ("max_list", C.c_int32), ("speculative_grid", C.c_int32), ("graph_replay", C.c_int32),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


_vp = C.c_void_p
_i32pp = C.POINTER(C.POINTER(C.c_int32))
_i64pp = C.POINTER(C.POINTER(C.c_int64))

# name -> (restype, argtypes): every symbol include/tnsb.h declares
SIGNATURES = {
    "tnsb_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "tnsb_destroy": (None, [_vp]),
    "tnsb_last_error": (C.c_char_p, [_vp]),
    "tnsb_version": (C.c_char_p, []),
    "tnsb_add_point_set_f32": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "tnsb_add_point_set_f64": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "tnsb_resize_point_set_f32": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int]),
    "tnsb_resize_point_set_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int]),
    "tnsb_set_search_radius": (C.c_int, [_vp, C.c_float]),
    "tnsb_set_cell_size": (C.c_int, [_vp, C.c_float]),
    "tnsb_set_symmetric_search": (C.c_int, [_vp, C.c_int]),
    "tnsb_set_active_search": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "tnsb_set_active_search_of_set": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "tnsb_set_all_searches": (C.c_int, [_vp, C.c_int]),
    "tnsb_set_option": (C.c_int, [_vp, C.c_int, C.c_int64]),
    "tnsb_set_stream": (C.c_int, [_vp, _vp]),
    "tnsb_get_n_sets": (C.c_int, [_vp]),
    "tnsb_get_n_points_in_set": (C.c_int, [_vp, C.c_int]),
    "tnsb_get_total_n_points": (C.c_int, [_vp]),
    "tnsb_is_search_active": (C.c_int, [_vp, C.c_int, C.c_int]),
    "tnsb_does_set_exist": (C.c_int, [_vp, C.c_int]),
    "tnsb_run": (C.c_int, [_vp]),
    "tnsb_get_neighborlists": (C.c_int, [_vp, C.c_int, C.c_int, _i32pp, _i64pp, C.POINTER(C.c_int64)]),
    "tnsb_get_neighborlists_u32": (C.c_int, [_vp, C.c_int, C.c_int, _i32pp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int64)]),
    "tnsb_get_neighborlists_device": (C.c_int, [_vp, C.c_int, C.c_int, _i32pp, _i64pp, C.POINTER(C.c_int64)]),
    "tnsb_prepare_zsort": (C.c_int, [_vp]),
    "tnsb_get_zsort_order": (C.c_int, [_vp, C.c_int, _i32pp, C.POINTER(C.c_int)]),
    "tnsb_apply_zsort_device": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int)]),
    "tnsb_apply_zsort_device_f32": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "tnsb_shard_aabb": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "tnsb_shard_histogram": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _vp]),
    "tnsb_shard_partition": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, _vp, C.c_int64,
                                       C.POINTER(C.c_int64)]),
    "tnsb_shard_window_create": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp]),
    "tnsb_shard_window_open": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "tnsb_shard_push": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, _vp]),
    "tnsb_shard_collect": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tnsb_shard_collect_flag": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp, C.POINTER(C.c_int)]),
    "tnsb_get_neighborlist_n_bytes": (C.c_uint64, [_vp]),
    "tnsb_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "tnsb_get_pair_neighbor_stats": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libtnsb.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built (python -c 'import __graft_entry__ as g; g.build()'). "
            "treensearch_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
