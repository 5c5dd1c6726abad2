"""GPU box: gx convolutions with output widths whose 16-byte chunk count per row is not a power of two (the copy-out's
general path): same check as tests/test_gpu_gx.py::test_within_grid_conv_matches_oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200"), os.path.join(ROOT, "tests")]
import test_gpu_gx as t

for cin, cout in ((64, 48), (128, 96), (64, 24), (32, 40)):
    try:
        t.test_within_grid_conv_matches_oracle(cin, cout)
        print("ok", cin, cout, flush=True)
    except Exception as e:  # noqa
        print("FAILED", cin, cout, type(e).__name__, str(e)[:300], flush=True)
