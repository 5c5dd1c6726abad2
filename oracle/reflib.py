"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/_ref/libasr_ref.so.

libasr_ref.so is the reference's own geometry code (cpp/lib/{octree,grid,
contouring,postprocess}.cpp) compiled unmodified by oracle/Makefile.  It is the
integer-half oracle.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASR_REF_VARIANT=unordered loads the build whose hash-map stand-in iterates in std::unordered_map order
# (only for the order-dependence experiment, oracle/shim/libcuckoo/cuckoohash_map.hh)
_PATH = os.path.join(_HERE, "_ref", "libasr_ref_unordered.so" if os.environ.get("ASR_REF_VARIANT") == "unordered"
                     else "libasr_ref.so")

_lib = None

GRID_FIELDS = [
    ("voxel_keys", np.uint64, None),
    ("voxel_centers", np.float32, 3),
    ("voxel_sizes", np.float32, None),
    ("neighbors_index", np.int32, None),
    ("neighbors_kernel_index", np.uint8, None),
    ("neighbors_row_splits", np.int64, None),
    ("up_neighbors_index", np.int32, None),
    ("up_neighbors_kernel_index", np.uint8, None),
    ("up_neighbors_row_splits", np.int64, None),
]


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                "oracle/_ref/libasr_ref.so missing: run `make -C oracle ref` "
                "where /root/reference exists")
        L = C.CDLL(_PATH)
        vp, u64, f32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_float)
        L.ref_last_error.restype = C.c_char_p
        L.ref_octree_create.restype = vp
        L.ref_octree_create.argtypes = [vp, u64, vp, vp, vp, C.c_float, C.c_int, C.c_int]
        L.ref_octree_free.argtypes = [vp]
        for n in ("ref_octree_num_leaves", "ref_octree_num_nodes"):
            getattr(L, n).restype = u64
            getattr(L, n).argtypes = [vp]
        L.ref_octree_leaves.argtypes = [vp, vp]
        L.ref_octree_nodes.argtypes = [vp, vp]
        L.ref_octree_params.argtypes = [vp, vp, vp, vp]
        L.ref_grids_create.restype = vp
        L.ref_grids_create.argtypes = [vp, C.c_int, C.c_int]
        L.ref_grids_free.argtypes = [vp]
        L.ref_grids_field.restype = u64
        L.ref_grids_field.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ref_duals_create.restype = vp
        L.ref_duals_create.argtypes = [vp]
        L.ref_duals_size.restype = u64
        L.ref_duals_size.argtypes = [vp]
        L.ref_duals_copy.argtypes = [vp, vp]
        L.ref_duals_free.argtypes = [vp]
        L.ref_mesh_create.restype = vp
        L.ref_mesh_create.argtypes = [vp, u64, vp, u64, vp, C.c_float]
        L.ref_mesh_remove_components.restype = vp
        L.ref_mesh_remove_components.argtypes = [vp, u64, vp, u64, C.c_int64, C.c_int64]
        L.ref_mesh_num_vertices.restype = u64
        L.ref_mesh_num_vertices.argtypes = [vp]
        L.ref_mesh_num_triangles.restype = u64
        L.ref_mesh_num_triangles.argtypes = [vp]
        L.ref_mesh_copy.argtypes = [vp, vp, vp]
        L.ref_mesh_free.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefOctree:
    """asr::Octree built by the reference's CreateOctreeFromPoints (octree.cpp:230)."""

    def __init__(self, points, radii, bb_min, bb_max, radius_scale=1.0, grow_steps=0, max_depth=21):
        L = lib()
        points = np.ascontiguousarray(points, np.float32)
        radii = np.ascontiguousarray(radii, np.float32)
        bb_min = np.ascontiguousarray(bb_min, np.float32)
        bb_max = np.ascontiguousarray(bb_max, np.float32)
        self.h = L.ref_octree_create(_p(points), points.shape[0], _p(radii), _p(bb_min), _p(bb_max),
                                     radius_scale, grow_steps, max_depth)
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            try:
                lib().ref_octree_free(self.h)
            except Exception:  # interpreter shutdown
                pass
            self.h = None

    def leaves(self):
        L = lib()
        out = np.empty(L.ref_octree_num_leaves(self.h), np.uint64)
        L.ref_octree_leaves(self.h, _p(out))
        return out

    def nodes(self):
        L = lib()
        out = np.empty(L.ref_octree_num_nodes(self.h), np.uint64)
        L.ref_octree_nodes(self.h, _p(out))
        return out

    def params(self):
        vs = np.empty(22, np.float32)
        ivs = np.empty(22, np.float32)
        off = np.empty(3, np.int32)
        lib().ref_octree_params(self.h, _p(vs), _p(ivs), _p(off))
        return vs, ivs, off

    def grids(self, num_levels, voxel_info_all_levels=False):
        """Mirror of pyCreateGridsFromOctree (module.cpp:163-228): empty vectors -> key omitted."""
        L = lib()
        g = L.ref_grids_create(self.h, num_levels, int(voxel_info_all_levels))
        if not g:
            raise RuntimeError(L.ref_last_error().decode())
        res = []
        for lev in range(num_levels):
            d = {}
            for fid, (name, dt, cols) in enumerate(GRID_FIELDS):
                n = L.ref_grids_field(g, lev, fid, None)
                if n == 0:
                    continue
                a = np.empty(n, dt)
                L.ref_grids_field(g, lev, fid, _p(a))
                d[name] = a.reshape(-1, cols) if cols else a
            res.append(d)
        L.ref_grids_free(g)
        return res

    def dual_vertex_indices(self):
        L = lib()
        d = L.ref_duals_create(self.h)
        if not d:
            raise RuntimeError(L.ref_last_error().decode())
        out = np.empty(L.ref_duals_size(d), np.uint64)
        L.ref_duals_copy(d, _p(out))
        L.ref_duals_free(d)
        return out.reshape(-1, 8)


def _mesh_out(m):
    L = lib()
    v = np.empty((L.ref_mesh_num_vertices(m), 3), np.float32)
    t = np.empty((L.ref_mesh_num_triangles(m), 3), np.int32)
    L.ref_mesh_copy(m, _p(v), _p(t))
    L.ref_mesh_free(m)
    return {"vertices": v, "triangles": t}


def create_triangle_mesh(values, dual_indices, node_positions, unsigned_threshold=1.0):
    """Reference CreateTriangleMesh (contouring.cpp:29)."""
    L = lib()
    values = np.ascontiguousarray(values, np.float32)
    dual_indices = np.ascontiguousarray(dual_indices, np.uint64)
    node_positions = np.ascontiguousarray(node_positions, np.float32)
    m = L.ref_mesh_create(_p(values), values.shape[0], _p(dual_indices), dual_indices.shape[0],
                          _p(node_positions), unsigned_threshold)
    if not m:
        raise RuntimeError(L.ref_last_error().decode())
    return _mesh_out(m)


def remove_connected_components(vertices, triangles, keep_n_largest_components, minimum_component_size=3):
    """Reference RemoveConnectedComponents (postprocess.cpp:141)."""
    L = lib()
    vertices = np.ascontiguousarray(vertices, np.float32)
    triangles = np.ascontiguousarray(triangles, np.int32)
    m = L.ref_mesh_remove_components(_p(vertices), vertices.shape[0], _p(triangles), triangles.shape[0],
                                     keep_n_largest_components, minimum_component_size)
    if not m:
        raise RuntimeError(L.ref_last_error().decode())
    return _mesh_out(m)
