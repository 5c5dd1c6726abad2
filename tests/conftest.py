import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "adaptive-surface-reconstruction_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def geom_checkers():
    """[(name, OctreeClass)] of the available geometry oracles: the compiled
    reference (oracle/_ref, prebuilt where /root/reference exists) and the port."""
    from oracle import geomlib, reflib
    out = [("port", geomlib.PortOctree)]
    if reflib.available():
        out.append(("reference", reflib.RefOctree))
    return out


@pytest.fixture(autouse=True)
def _reset_kernel_options(request):
    """GPU tests may change development knobs of the kernels; restore the defaults."""
    yield
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from asr_b200 import _lib
            for name in ("gx_acc_groups", "gx_tma_gather", "gx_l1_gather", "gx_max_stages", "gx_ablate", "gx_single_tmem", "gx_one_team", "gx_trace"):
                _lib.set_option(name, 0)
