"""PLY point-cloud reader and triangle-mesh writer for the command line tool (SURVEY.md §8 row
f-4).  Mirrors what the reference's `asrtool` reads and writes: vertex properties x, y, z, nx, ny,
nz and an optional per-point radius called `value` or `radius` (cpp/bin/main.cpp:25-112, PlyReader
in cpp/bin/plyreader.h); the mesh goes out as a binary little-endian PLY with float32 vertices and
`list uchar int vertex_indices` faces (main.cpp:161-171 writes it through Open3D).
ascii, binary_little_endian and binary_big_endian inputs are supported."""
import numpy as np

_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
          "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
          "double": "f8", "float64": "f8"}


def _read_header(f):
    if f.readline().strip() != b"ply":
        raise RuntimeError("not a PLY file")
    fmt, elements = None, []
    while True:
        line = f.readline()
        if not line:
            raise RuntimeError("unexpected end of the PLY header")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1]["props"].append((tok[4], ("list", _TYPES[tok[2]], _TYPES[tok[3]])))
            else:
                elements[-1]["props"].append((tok[2], _TYPES[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise RuntimeError("unsupported PLY format %r" % fmt)
    return fmt, elements


def read_points(path):
    """(points f32[N,3], normals f32[N,3], radii f32[N] or empty) like ReadPoints in main.cpp:
    empty arrays when the vertex element or one of x, y, z, nx, ny, nz is missing."""
    empty = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    with open(path, "rb") as f:
        fmt, elements = _read_header(f)
        for el in elements:
            scalar = all(not isinstance(t, tuple) for _, t in el["props"])
            if el["name"] != "vertex":
                if not scalar or fmt == "ascii":
                    if fmt == "ascii":
                        for _ in range(el["count"]):
                            f.readline()
                        continue
                    raise RuntimeError("cannot skip the list element %r before the vertices" % el["name"])
                end = "<" if fmt == "binary_little_endian" else ">"
                f.seek(el["count"] * np.dtype([(n, end + t) for n, t in el["props"]]).itemsize, 1)
                continue
            if not scalar:
                raise RuntimeError("list properties in the vertex element are not supported")
            names = [n for n, _ in el["props"]]
            if el["count"] == 0 or any(k not in names for k in ("x", "y", "z", "nx", "ny", "nz")):
                return empty
            if fmt == "ascii":
                data = np.loadtxt(f, dtype=np.float64, max_rows=el["count"], ndmin=2)
                col = {n: data[:, i] for i, n in enumerate(names)}
            else:
                end = "<" if fmt == "binary_little_endian" else ">"
                rec = np.fromfile(f, dtype=np.dtype([(n, end + t) for n, t in el["props"]]), count=el["count"])
                col = {n: rec[n] for n in names}
            pts = np.stack([col["x"], col["y"], col["z"]], 1).astype(np.float32)
            nrm = np.stack([col["nx"], col["ny"], col["nz"]], 1).astype(np.float32)
            rad = np.zeros(0, np.float32)
            for key in ("value", "radius"):  # main.cpp:100-102
                if key in col:
                    rad = np.asarray(col[key], np.float32)
                    break
            return pts, nrm, rad
    return empty


def write_mesh(path, vertices, triangles):
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    triangles = np.ascontiguousarray(triangles, np.int32).reshape(-1, 3)
    header = ("ply\nformat binary_little_endian 1.0\ncomment asr_b200\nelement vertex %d\nproperty float x\n"
              "property float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\n"
              "end_header\n" % (len(vertices), len(triangles)))
    faces = np.empty(len(triangles), dtype=[("n", "u1"), ("v", "<i4", 3)])
    faces["n"] = 3
    faces["v"] = triangles
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(vertices.astype("<f4").tobytes())
        f.write(faces.tobytes())


def read_mesh(path):
    """(vertices f32[V,3], triangles i32[T,3]) of a mesh written by write_mesh (tests)."""
    with open(path, "rb") as f:
        fmt, elements = _read_header(f)
        assert fmt == "binary_little_endian"
        nv = next(e["count"] for e in elements if e["name"] == "vertex")
        nf = next(e["count"] for e in elements if e["name"] == "face")
        v = np.fromfile(f, "<f4", nv * 3).reshape(-1, 3)
        faces = np.fromfile(f, np.dtype([("n", "u1"), ("v", "<i4", 3)]), nf)
        return v, faces["v"].astype(np.int32)
