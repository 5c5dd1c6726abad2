"""ctypes binding of libasr_b200.so (C ABI declared in include/asr_b200.h).

The shared library is the product: it is built in-tree by `make -C csrc` (or
`__graft_entry__.build()`), holds every CUDA kernel of the hot path, and has no
CPU fallback.  If it is missing or cannot be loaded this module raises — nothing
in the package silently degrades to PyTorch or CPU code.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasr_b200.so")

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/asr_b200.h one to one
SIGNATURES = {
    "asr_version": (_i32, []),
    "asr_last_error": (C.c_char_p, []),
    "asr_kernel_launches": (_i64, []),
    "asr_set_option": (_i32, [C.c_char_p, _i32]),
    "asr_pool_stats": (_i32, [_pi64, _pi64, _pi64, _pi64]),
    "asr_profile_enable": (None, [_i32]),
    "asr_profile_reset": (None, []),
    "asr_profile_count": (_i32, []),
    "asr_profile_get": (_i32, [_i32, C.c_char_p, _i32, C.POINTER(C.c_double), _pi64, C.POINTER(C.c_double)]),
    "asr_octree_create": (_i32, [_vp, _vp, _i64, _vp, _vp, _f32, _i32, _i32, _vp, _pp]),
    "asr_octree_destroy": (None, [_vp]),
    "asr_octree_num_leaves": (_i64, [_vp]),
    "asr_octree_num_nodes": (_i64, [_vp]),
    "asr_octree_balance_rounds": (_i32, [_vp]),
    "asr_octree_get_leaves": (_i32, [_vp, _vp, _vp]),
    "asr_octree_get_frame": (_i32, [_vp, _vp, _vp, _vp]),
    "asr_grids_build": (_i32, [_vp, _i32, _i32, _vp]),
    "asr_grids_level_size": (_i32, [_vp, _i32, _pi64, _pi64]),
    "asr_grids_get": (_i32, [_vp, _i32] + [_vp] * 10),
    "asr_duals_count": (_i32, [_vp, _pi64, _vp]),
    "asr_duals_fill": (_i32, [_vp, _vp, _vp]),
    "asr_duals_begin": (_i32, [_vp, _vp]),
    "asr_duals_check": (_i32, [_vp]),
    "asr_radius_search_create": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _pp, _pi64]),
    "asr_radius_search_fill": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "asr_radius_search_destroy": (None, [_vp]),
    "asr_scale_compatibility": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "asr_aggregation_importance": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "asr_continuous_conv": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                   _i32, _vp, _i32, _vp, _vp]),
    "asr_conv_plan_create": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _pp]),
    "asr_conv_plan_destroy": (None, [_vp]),
    "asr_packed_conv_filters_size": (_i64, [_i32, _i32, _i32]),
    "asr_pack_conv_filters": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "asr_sparse_conv": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "asr_gx_plan_begin": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _pp]),
    "asr_gx_plan_finish": (_i32, [_vp, _vp, _pi64]),
    "asr_gx_plan_destroy": (None, [_vp]),
    "asr_gx_packed_filters_bytes": (_i64, [_i32, _i32, _i32]),
    "asr_gx_pack_filters": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "asr_gx_from_f32": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp]),
    "asr_gx_to_f32": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "asr_gx_scale_rows": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "asr_gx_conv": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32,
                           _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "asr_gx_overflow": (_i32, [_vp, C.POINTER(C.c_int)]),
    "asr_gx_trace": (_i32, [_vp, _i32, C.POINTER(C.c_uint)]),
    "asr_shard_positions": (_i32, [_vp, _i64, _vp, _vp]),
    "asr_shard_owner": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "asr_shard_need_mask": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp]),
    "asr_shard_push": (_i32, [_vp, _i32, _i32, _i64, _i64, _vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "asr_reduce_subarrays_sum": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "asr_invert_neighbors_list": (_i32, [_i64, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "asr_decode": (_i32, [_vp, _vp, _i64] + [_vp] * 9),
    "asr_packed_weights_size": (_i64, [_i32, _i32]),
    "asr_pack_weights": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "asr_dense_tf32x3": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "asr_contour_count": (_i32, [_vp, _vp, _i64, _f32, _vp, _vp, _pi64, _vp]),
    "asr_contour_fill": (_i32, [_vp, _vp, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asr_contour_triangles_create": (_i32, [_vp, _vp, _i64, _f32, _vp, _i64, _i64, _vp, _pp, _pi64, _pi64]),
    "asr_contour_triangles_fill": (_i32, [_vp, _vp, _vp, _vp]),
    "asr_contour_triangles_destroy": (None, [_vp]),
    "asr_mesh_components": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "asr_kdtree_create": (_i32, [_vp, _i64, _vp, _pp]),
    "asr_kdtree_k_radius": (_i32, [_vp, _i32, _vp, _vp]),
    "asr_kdtree_inlier": (_i32, [_vp, _vp, _f32, _i32, _i32, _vp, _vp]),
    "asr_radius_neighbor_counts": (_i32, [_vp, _i64, _vp, _vp, _vp]),
}

_lib = None


def pool_stats():
    """(reserved, used, release threshold, high-water mark in use) bytes of the library's stream-ordered memory pool"""
    a, b, c, d = _i64(0), _i64(0), _i64(0), _i64(0)
    check(lib().asr_pool_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return a.value, b.value, c.value, d.value


def lib():
    """Loads libasr_b200.so (once).  Raises if the CUDA extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "asr_b200: %s not found — the CUDA extension is not built (run `make -C "
                "adaptive-surface-reconstruction_b200/csrc` or __graft_entry__.build()); there is no "
                "CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Maps the C status to the exception the reference's pybind layer raises
    (std::invalid_argument -> ValueError, std::runtime_error -> RuntimeError)."""
    if rc == 0:
        return
    msg = lib().asr_last_error().decode(errors="replace")
    if rc == 1:
        raise ValueError(msg)
    raise RuntimeError(msg)


def kernel_launches():
    return int(lib().asr_kernel_launches())


def set_option(name, value):
    check(lib().asr_set_option(name.encode(), int(value)))


def profile_enable(on=True):
    lib().asr_profile_enable(int(bool(on)))


def profile_reset():
    lib().asr_profile_reset()


def profile_read():
    """{kernel name: {"ms": total device ms, "launches": n, "flops": algorithmic flops}}"""
    L = lib()
    out = {}
    for i in range(L.asr_profile_count()):
        name = C.create_string_buffer(128)
        ms, fl, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        if L.asr_profile_get(i, name, 128, C.byref(ms), C.byref(n), C.byref(fl)) == 0:
            out[name.value.decode()] = {"ms": ms.value, "launches": n.value, "flops": fl.value}
    return out
