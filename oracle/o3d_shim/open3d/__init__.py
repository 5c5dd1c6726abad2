"""TEST INFRASTRUCTURE ONLY — minimal `open3d` stand-in backed by the CPU oracle
(oracle/ops_cpu.py) so that the reference's models/v0/net_definitions_torch.py
can be imported and run unmodified in the build container (Open3D itself is not
installed).  Used to validate oracle/model_cpu.py and to generate tests/golden/.
The product's own drop-in shim lives in adaptive-surface-reconstruction_b200/open3d.
"""
