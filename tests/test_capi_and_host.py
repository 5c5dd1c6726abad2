"""CPU: the C-ABI library loads and exports every symbol include/asr_b200.h
declares (no compute without a GPU), the ctypes table mirrors the header, the
product fails loudly without CUDA, and the host-side logic behaves."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "asr_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from asr_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libasr_b200.so does not export " + n
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table mirrors the header one to one
    assert _lib.lib().asr_version() == 100


def test_every_declaration_cites_the_reference():
    src = open(HEADER).read()
    for anchor in ("octree.cpp:230", "grid.cpp:245", "grid.cpp:450", "nsearch.cpp:107", "common_torch.py:95",
                   "net_definitions_torch.py:655", "contouring.cpp:66", "module.cpp:372"):
        assert anchor in src, anchor


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from asr_b200 import _lib, ops
    L = _lib.lib()
    h = ctypes.c_void_p(0)
    bb = (ctypes.c_float * 3)(0, 0, 0)
    rc = L.asr_octree_create(None, None, 0, bb, bb, 1.0, 0, 21, None, ctypes.byref(h))
    assert rc == 3 and b"no CUDA device" in L.asr_last_error()
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.multi_radius_search(torch.zeros(3, 3), torch.zeros(1, 3), torch.ones(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        import open3d.ml.torch as ml3d
        ml3d.ops.reduce_subarrays_sum(torch.zeros(3), torch.tensor([0, 3]))
    # the product package never imports the oracle
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path[:0]=[%r, %r]; import asr_b200, asr_b200.model, asr_b200.pipeline, "
                        "adaptivesurfacereconstruction, open3d.ml.torch; "
                        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules); print('clean')"
                        % (ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200"))],
                       capture_output=True, text=True)
    assert "clean" in r.stdout, r.stderr[-1500:]


def test_python_module_signatures_and_errors():
    """Same names / keyword defaults / exception types as cpp/pybind/module.cpp."""
    import inspect

    import adaptivesurfacereconstruction as asr
    sig = inspect.signature(asr.reconstruct_surface)
    want = {"point_radius_scale": 1.0, "density_percentile_threshold": 10.0, "point_radius_estimation_knn": 24,
            "octree_max_depth": 21, "contouring_value_threshold": 1.0,
            "keep_n_connected_components": np.iinfo(np.int64).max, "minimum_component_size": 3}
    for k, v in want.items():
        assert sig.parameters[k].default == v
    assert list(sig.parameters)[:3] == ["points", "normals", "radii"]
    sig = inspect.signature(asr.create_octree)
    assert [(k, p.default) for k, p in sig.parameters.items()][4:] == [("radius_scale", 1), ("grow_steps", 0),
                                                                      ("max_depth", 21)]
    assert inspect.signature(asr.create_grids_from_octree).parameters["voxel_info_all_levels"].default is False
    assert inspect.signature(asr.remove_connected_components).parameters["minimum_component_size"].default == 3
    with pytest.raises(ValueError):
        asr.create_octree(np.zeros((4, 2), np.float32), np.zeros(4, np.float32), [0, 0, 0], [1, 1, 1])
    with pytest.raises(ValueError):
        asr.create_octree(np.zeros((4, 3), np.float32), np.zeros(5, np.float32), [0, 0, 0], [1, 1, 1])
    with pytest.raises(ValueError):
        asr.reconstruct_surface(np.zeros((4, 3)), np.zeros((3, 3)), np.zeros(4))
    with pytest.raises(ValueError):
        asr.remove_connected_components(np.zeros((4, 2)), np.zeros((1, 3), np.int32), 1)
    assert isinstance(asr.get_version_str(), str) and isinstance(asr.get_third_party_notices(), str)


def test_model_mirror_matches_reference_state_dict_layout():
    from asr_b200 import model
    from oracle import model_cpu
    for levels in (3, 5, 6):
        net = model.UNet(levels)
        P = model_cpu.init_params(levels)
        sd = net.state_dict()
        assert set(sd) == set(P)
        assert all(tuple(sd[k].shape) == tuple(P[k].shape) for k in sd)
    net = model.seeded_weights(model.UNet(3), seed=1)
    k, b = net.sparseconv_encblock0.first_conv()
    assert k.shape == (55, 32, 64) and b.shape == (64,)
    assert torch.equal(k[:, :, 56:], net.sparseconv_encblock0.conv1b.kernel)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_reference_model_script_imports_unmodified_on_the_product_shim():
    code = ("import sys, warnings; warnings.filterwarnings('ignore'); sys.path[:0]=[%r, '/root/reference']; "
            "import open3d.ml.torch as ml3d; from models.v0.net_definitions_torch import UNet5; "
            "from asr_b200 import model; n = UNet5(with_importance='all', normalized_channels=8, "
            "residual_skip_connection=True); m = model.UNet(5); m.load_state_dict(n.state_dict()); print('OK')"
            % os.path.join(ROOT, "adaptive-surface-reconstruction_b200"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "OK" in r.stdout, r.stderr[-1500:]


def test_clouds_are_deterministic_and_bench_byte_model():
    from asr_b200 import clouds
    a, b = clouds.thingi_like(20000, seed=2), clouds.thingi_like(20000, seed=2)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["points"].dtype == np.float32 and a["points"].shape == (20000, 3)
    assert np.allclose(np.linalg.norm(a["normals"], axis=1), 1, atol=1e-5) and a["radii"].min() > 0
    sys.path.insert(0, ROOT)
    import bench
    sizes = {"N": 10, "V": [4, 2], "E": [20, 6], "P": 30, "D": 5, "M": 3}
    by = bench.algorithmic_bytes(sizes, [{"V_in": 4, "V_out": 4, "E": 20, "K": 55, "Cin": 32, "Cout": 64,
                                          "importance": True}])
    assert by["sparse_conv_stack"] == 4 * 4 * 32 + 4 * 4 * 64 + 5 * 20 + 8 * 5 + 4 * 55 * 32 * 64 + 32
    assert by["search"] == 12 * 10 + 16 * 4 + 12 * 30 + 8 * 5


def test_ply_reader_and_writer(tmp_path):
    """Row f-4: the PLY flavours asrtool reads (cpp/bin/main.cpp:25-112) and the mesh it writes."""
    from asr_b200 import plyio
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((50, 3)).astype(np.float32)
    nrm = rng.standard_normal((50, 3)).astype(np.float32)
    rad = rng.uniform(0.1, 1, 50).astype(np.float32)
    head = "ply\nformat %s 1.0\ncomment test\nelement vertex 50\n" + "".join(
        "property float %s\n" % n for n in ("x", "y", "z", "nx", "ny", "nz")) + "%send_header\n"
    # ascii with radii called `value`
    a = tmp_path / "a.ply"
    with open(a, "w") as f:
        f.write(head % ("ascii", "property double value\n"))
        for i in range(50):
            f.write(" ".join(repr(float(v)) for v in (*pts[i], *nrm[i], rad[i])) + "\n")
    p, n, r = plyio.read_points(str(a))
    assert np.array_equal(p, pts) and np.array_equal(n, nrm) and np.array_equal(r, rad)
    # binary little / big endian, no radii
    for fmt, end in (("binary_little_endian", "<"), ("binary_big_endian", ">")):
        b = tmp_path / (fmt + ".ply")
        with open(b, "wb") as f:
            f.write((head % (fmt, "")).encode())
            f.write(np.concatenate([pts, nrm], 1).astype(end + "f4").tobytes())
        p, n, r = plyio.read_points(str(b))
        assert np.array_equal(p, pts) and np.array_equal(n, nrm) and r.size == 0
    # normals missing -> empty result like ReadPoints
    c = tmp_path / "c.ply"
    with open(c, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\n"
                "end_header\n0 0 0\n")
    assert plyio.read_points(str(c))[0].shape == (0, 3)
    tri = np.array([[0, 1, 2], [2, 3, 4]], np.int32)
    m = tmp_path / "m.ply"
    plyio.write_mesh(str(m), pts, tri)
    v, t = plyio.read_mesh(str(m))
    assert np.array_equal(v, pts) and np.array_equal(t, tri)


def test_bench_byte_model_matches_the_survey_formulas():
    """bench.py's algorithmic byte model on a hand-checked configuration."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    sizes = {"N": 1000, "V": [100, 20], "E": [700, 120], "P": 1500, "D": 60, "M": 30}
    convs = [{"V_in": 100, "V_out": 100, "E": 700, "K": 55, "Cin": 32, "Cout": 64, "importance": True},
             {"V_in": 100, "V_out": 20, "E": 100, "K": 9, "Cin": 64, "Cout": 128, "importance": False}]
    got = b.algorithmic_bytes(sizes, convs)
    assert got["octree_keys"] == 16 * 1000 + 8 * 100
    assert got["neighbor_tables"] == (8 * 100 + 5 * 700 + 8 * 101) + (8 * 20 + 5 * 120 + 8 * 21) + 21 * 100
    assert got["duals"] == 8 * 100 + 64 * 60
    assert got["search"] == 12 * 1000 + 16 * 100 + 12 * 1500 + 8 * 101
    assert got["continuous_conv"] == 12 * 1500 + 8 * 101 + 28 * 1000 + 16 * 100 + 128 * 100 + 4 * 1500
    c0 = 4 * 100 * 32 + 4 * 100 * 64 + 5 * 700 + 8 * 101 + 4 * 55 * 32 * 64 + 4 * 100 + 4 * 100
    c1 = 4 * 100 * 64 + 4 * 20 * 128 + 5 * 100 + 8 * 21 + 4 * 9 * 64 * 128
    assert got["sparse_conv_stack"] == c0 + c1
    assert got["decode"] == 136 * 100 and got["contour"] == 64 * 60 + 20 * 100 + 12 * 30
    assert b.conv_traffic_per_launch() is None or b.conv_traffic_per_launch() > 0


_FULL_SIGNATURE_MODULE = '''
import torch
class M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.kernel = torch.nn.Parameter(torch.zeros(55, 4, 8))
        self.bias = torch.nn.Parameter(torch.zeros(8))
    def forward(self, x, pos, ext, idx, kidx, rs):
        e = torch.empty((0,), dtype=torch.float32)
        # the calls a graph traced against real Open3D v0.14.1 holds: every argument, defaulted ones included
        y = torch.ops.open3d.sparse_conv(self.kernel, x, e, idx, kidx, e, rs, False, 64)
        z = torch.ops.open3d.continuous_conv(torch.zeros(4, 4, 4, 4, 8), pos, ext, torch.zeros(3), pos, x, e, idx, e, rs,
                                             True, "ball_to_cube_radial", True, "linear", 64)
        s = torch.ops.open3d.reduce_subarrays_sum(ext, rs)
        a, b, c = torch.ops.open3d.invert_neighbors_list(5, idx, rs, kidx)
        return y + self.bias, z, s, a
'''


def test_weights_loader_resolves_full_open3d_signatures(tmp_path):
    """an archive whose graph calls `open3d::*` with Open3D's full argument lists must
    load through the product's loader (the shim is imported first, errors are not swallowed)."""
    pkg = os.path.join(ROOT, "adaptive-surface-reconstruction_b200")
    path = str(tmp_path / "model.pt")
    make = tmp_path / "make_archive.py"  # torch.jit.script reads the class source from its file
    make.write_text("import sys; sys.path[:0]=[%r]; import open3d.ml.torch.ops\n" % pkg + _FULL_SIGNATURE_MODULE +
                    "torch.jit.script(M()).save(%r); print('SAVED')" % path)
    r = subprocess.run([sys.executable, str(make)], capture_output=True, text=True, timeout=300)
    assert "SAVED" in r.stdout, r.stderr[-2000:]
    # a fresh process that has NOT imported the shim itself: the loader has to register the ops
    load = ("import sys; sys.path[:0]=[%r]; from asr_b200 import model; sd = model.load_weights_file(%r); "
            "print(sorted(sd), tuple(sd['kernel'].shape))" % (pkg, path))
    r = subprocess.run([sys.executable, "-c", load], capture_output=True, text=True, timeout=300)
    assert "['bias', 'kernel'] (55, 4, 8)" in r.stdout, r.stderr[-2000:]
    # module-level lookup of the python module (asr.cpp:50,138-141): ASR_RESOURCE_DIR/model.pt
    code = ("import os, sys; sys.path[:0]=[%r]; os.environ['ASR_RESOURCE_DIR']=%r; "
            "import adaptivesurfacereconstruction as asr\n"
            "try:\n    asr._load_model()\nexcept RuntimeError as e:\n    print('ERR', str(e)[:200])" % (pkg, str(tmp_path)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    # the toy archive is not a UNet5: the state dict is found and rejected by load_state_dict (keys named)
    assert "ERR" in r.stdout and "sparseconv_encblock0" in r.stdout, r.stdout + r.stderr[-2000:]
    # a plain state-dict file goes through torch.load
    from asr_b200 import model
    torch.save({"a": torch.ones(2)}, str(tmp_path / "sd.pt"))
    assert torch.equal(model.load_weights_file(str(tmp_path / "sd.pt"))["a"], torch.ones(2))
    # a TorchScript archive that cannot be resolved raises its own error (not retried as a pickle)
    bad = tmp_path / "make_bad.py"
    bad.write_text("import sys, torch; sys.path[:0]=[%r]; import open3d.ml.torch.ops\n"
           "lib = torch.library.Library('open3d', 'FRAGMENT'); lib.define('not_an_open3d_op(Tensor x) -> Tensor')\n"
           "class B(torch.nn.Module):\n    def forward(self, x):\n        return torch.ops.open3d.not_an_open3d_op(x)\n"
           "torch.jit.script(B()).save(%r); print('SAVED')" % (pkg, str(tmp_path / "bad.pt")))
    r = subprocess.run([sys.executable, str(bad)], capture_output=True, text=True, timeout=300)
    assert "SAVED" in r.stdout, r.stderr[-2000:]
    with pytest.raises(Exception, match="not_an_open3d_op"):
        model.load_weights_file(str(tmp_path / "bad.pt"))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: compiled reference geometry + restated network, no GPU, no product code)
    on a small cloud: one JSON line with the same metric / unit as the GPU arm, its own cpu_baseline and a zero-copy e2e."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--points", "20000",
                          "--levels", "3", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"].startswith("reference") and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "libasr_b200" not in out.stderr  # the product library is not on this path
