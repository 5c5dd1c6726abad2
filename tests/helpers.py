import numpy as np
import torch


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def u64(t):
    """int64 CUDA tensor holding uint64 bit patterns -> numpy uint64"""
    return t.cpu().numpy().view(np.uint64)


def small_clouds():
    from asr_b200 import clouds
    c = clouds.sphere(20000, seed=0)
    yield "sphere20k", c, 5
    c = clouds.adaptive_blob(30000, seed=1)
    yield "blob30k", c, 5
    c = clouds.thingi_like(40000, seed=2)
    yield "thingi40k", c, 6
    c = clouds.gaussian_blob(30000, seed=3)
    yield "gauss30k", c, 3
