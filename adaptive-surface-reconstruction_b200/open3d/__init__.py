"""Drop-in stand-in for the slice of Open3D the reference's model code imports
(`open3d.ml.torch.ops`, `open3d.ml.torch.layers`), backed by
the asr_b200 CUDA kernels.  Put `adaptive-surface-reconstruction_b200/` on
sys.path ahead of a real Open3D to run models/v0/net_definitions_torch.py and
models/common_torch.py unmodified on the B200 backend."""
__version__ = "0.14.1+asr_b200"
