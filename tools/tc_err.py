import sys, torch
sys.path.insert(0,'/root/repo/adaptive-surface-reconstruction_b200')
from asr_b200 import ops
torch.backends.cuda.matmul.allow_tf32=False
for pos in (False, True):
    for K in (32,128,512,2048):
        g=torch.Generator().manual_seed(K)
        a=torch.randn((2048,K),generator=g); w=torch.randn((K,64),generator=g)/K**0.5
        if pos: a=a.abs(); w=w.abs()
        ref=a.double()@w.double()
        out=ops.dense_tf32x3(a.cuda(), ops.pack_weights(w.cuda())).cpu().double()
        f32=(a.cuda()@w.cuda()).cpu().double()
        sc=ref.abs().max()
        d=(out-ref)
        print('pos' if pos else 'rnd','K',K,'tc max rel %.2e mean signed %.2e | fp32 max rel %.2e'%((d.abs().max()/sc).item(), (d.mean()/sc).item(), ((f32-ref).abs().max()/sc).item()))
