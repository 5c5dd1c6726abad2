"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/liboracle_geom.so, this
repo's own CPU restatement of the octree / grid / dual-cell / contouring-vertex
algorithms (oracle/geom_oracle.cpp).  Same Python surface as oracle/reflib.py so
tests can run either checker.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .reflib import GRID_FIELDS

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle_geom.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        L = C.CDLL(_PATH)
        vp, u64 = C.c_void_p, C.c_uint64
        L.og_last_error.restype = C.c_char_p
        L.og_octree_create.restype = vp
        L.og_octree_create.argtypes = [vp, u64, vp, vp, vp, C.c_float, C.c_int, C.c_int]
        L.og_octree_free.argtypes = [vp]
        for n in ("og_octree_num_leaves", "og_octree_num_nodes", "og_duals_size", "og_contour_num_vertices"):
            getattr(L, n).restype = u64
            getattr(L, n).argtypes = [vp]
        L.og_octree_leaves.argtypes = [vp, vp]
        L.og_octree_nodes.argtypes = [vp, vp]
        L.og_octree_params.argtypes = [vp, vp, vp, vp]
        L.og_grids_create.restype = vp
        L.og_grids_create.argtypes = [vp, C.c_int, C.c_int]
        L.og_grids_free.argtypes = [vp]
        L.og_grids_field.restype = u64
        L.og_grids_field.argtypes = [vp, C.c_int, C.c_int, vp]
        L.og_duals_create.restype = vp
        L.og_duals_create.argtypes = [vp]
        L.og_duals_copy.argtypes = [vp, vp]
        L.og_duals_free.argtypes = [vp]
        L.og_contour_create.restype = vp
        L.og_contour_create.argtypes = [vp, u64, vp, u64, vp, C.c_float]
        L.og_contour_copy.argtypes = [vp, vp, vp]
        L.og_contour_free.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PortOctree:
    """Restatement of CreateOctreeFromPoints (octree.cpp:230-280)."""

    def __init__(self, points, radii, bb_min, bb_max, radius_scale=1.0, grow_steps=0, max_depth=21):
        L = lib()
        points = np.ascontiguousarray(points, np.float32)
        radii = np.ascontiguousarray(radii, np.float32)
        bb_min = np.ascontiguousarray(bb_min, np.float32)
        bb_max = np.ascontiguousarray(bb_max, np.float32)
        self.h = L.og_octree_create(_p(points), points.shape[0], _p(radii), _p(bb_min), _p(bb_max),
                                    radius_scale, grow_steps, max_depth)
        if not self.h:
            raise RuntimeError(L.og_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            try:
                lib().og_octree_free(self.h)
            except Exception:  # interpreter shutdown
                pass
            self.h = None

    def leaves(self):
        out = np.empty(lib().og_octree_num_leaves(self.h), np.uint64)
        lib().og_octree_leaves(self.h, _p(out))
        return out

    def nodes(self):
        out = np.empty(lib().og_octree_num_nodes(self.h), np.uint64)
        lib().og_octree_nodes(self.h, _p(out))
        return out

    def params(self):
        vs = np.empty(22, np.float32)
        ivs = np.empty(22, np.float32)
        off = np.empty(3, np.int32)
        lib().og_octree_params(self.h, _p(vs), _p(ivs), _p(off))
        return vs, ivs, off

    def grids(self, num_levels, voxel_info_all_levels=False):
        L = lib()
        g = L.og_grids_create(self.h, num_levels, int(voxel_info_all_levels))
        res = []
        for lev in range(num_levels):
            d = {}
            for fid, (name, dt, cols) in enumerate(GRID_FIELDS):
                if lev == num_levels - 1 and name.startswith("up_"):
                    continue
                n = L.og_grids_field(g, lev, fid, None)
                if n == 0:
                    continue
                a = np.empty(n, dt)
                L.og_grids_field(g, lev, fid, _p(a))
                d[name] = a.reshape(-1, cols) if cols else a
            res.append(d)
        L.og_grids_free(g)
        return res

    def dual_vertex_indices(self):
        L = lib()
        d = L.og_duals_create(self.h)
        if not d:
            raise RuntimeError(L.og_last_error().decode())
        out = np.empty(L.og_duals_size(d), np.uint64)
        L.og_duals_copy(d, _p(out))
        L.og_duals_free(d)
        return out.reshape(-1, 8)


def contour_vertices(values, dual_indices, node_positions, unsigned_threshold=1.0):
    """Vertex part of CreateTriangleMesh (contouring.cpp:66-199).

    Returns (vertices f32[M,3], dual_of_vertex u64[M]); vertex i belongs to the
    i-th intersecting dual in dual order."""
    L = lib()
    values = np.ascontiguousarray(values, np.float32)
    dual_indices = np.ascontiguousarray(dual_indices, np.uint64)
    node_positions = np.ascontiguousarray(node_positions, np.float32)
    m = L.og_contour_create(_p(values), values.shape[0], _p(dual_indices), dual_indices.shape[0],
                            _p(node_positions), unsigned_threshold)
    n = L.og_contour_num_vertices(m)
    v = np.empty((n, 3), np.float32)
    d = np.empty(n, np.uint64)
    L.og_contour_copy(m, _p(v), _p(d))
    L.og_contour_free(m)
    return v, d
