"""treensearch_b200 -- B200-native fixed-radius neighbour search behind the tns::TreeNSearch API.

The hot path (cell hash -> Morton keys -> radix sort -> cell start/end -> 27-cell query -> ragged neighbour lists) is
hand-written sm_100a CUDA in csrc/, exported through the C ABI of include/tnsb.h (libtnsb.so).  This package is the
Python host-side mirror of the reference's public class; the C++ mirror is include/tns/TreeNSearch.h.
"""
from .api import NeighborList, TreeNSearch, TreeNSearchError  # noqa: F401
from . import clouds  # noqa: F401
from ._lib import (TNSB_OPT_HOST_RESULTS, TNSB_OPT_LIST_CAPACITY, TNSB_OPT_PIN_USER_MEMORY,  # noqa: F401
                   TNSB_OPT_QUERY_LIMIT, TNSB_OPT_SORT_LISTS, TNSB_OPT_POINT_STRIDE, TNSB_OPT_ZERO_COPY_RESULTS, TNSB_OPT_QUERY_KERNEL, TNSB_OPT_BUILD, LIB_PATH)

__all__ = ["TreeNSearch", "NeighborList", "TreeNSearchError", "clouds"]
