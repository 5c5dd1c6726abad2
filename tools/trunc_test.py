"""Dev: does tcgen05 kind::tf32 ignore the low 13 mantissa bits of its operands (truncate)?"""
import sys
sys.path.insert(0, '/root/repo/adaptive-surface-reconstruction_b200'); sys.path.insert(0, '/root/repo')
import torch
from asr_b200 import _lib, ops
torch.manual_seed(0)
V, Cin, Cout, K = 20000, 128, 128, 55
E = V * 8
idx = torch.randint(0, V, (E,), dtype=torch.int32, device='cuda')
rows = torch.arange(V, device='cuda').repeat_interleave(8)
slot = torch.randint(0, K, (E,), dtype=torch.uint8, device='cuda')
rs = torch.arange(0, E + 1, 8, dtype=torch.int64, device='cuda')
plan = ops.ConvPlan(idx, slot, rs, K)
x = torch.randn(V, Cin, device='cuda'); W = torch.randn(K, Cin, Cout, device='cuda') / 30
ref = torch.zeros(V, Cout, dtype=torch.float64, device='cuda')
xs = x.double(); Wd = W.double()
for k in range(K):
    m = slot == k
    ref.index_add_(0, rows[m], xs[idx[m].long()] @ Wd[k])
for mode in (0, 2):
    _lib.set_option("pm_debug", mode)
    out = ops.sparse_conv(plan, W, x)
    print("mode", mode, "max abs err", float((out.double() - ref).abs().max()), "ref scale", float(ref.abs().max()))
