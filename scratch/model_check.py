import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle/o3d_shim'); sys.path.insert(0,'/root/reference')
from oracle import reflib, ops_cpu, model_cpu
from models.v0.net_definitions_torch import UNet5
torch.manual_seed(0)
net=UNet5(with_importance='all', normalized_channels=8, residual_skip_connection=True)
sd=net.state_dict()
P=model_cpu.init_params(5)
print(len(sd), len(P)); assert set(sd)==set(P), set(sd)^set(P)
for k in sd: assert sd[k].shape==P[k].shape,(k,sd[k].shape,P[k].shape)
# build small input
rng=np.random.default_rng(0); N=20000
p=(rng.standard_normal((N,3))*0.25).astype(np.float32)
d=np.linalg.norm(p,axis=1); r=(0.01*np.exp(3*d)*2**rng.uniform(0,1,N)).astype(np.float32)
nrm=p/np.linalg.norm(p,axis=1,keepdims=True)
tr=reflib.RefOctree(p,r,p.min(0)-0.1,p.max(0)+0.1)
grids=tr.grids(5,True)
print([len(g['voxel_sizes']) for g in grids])
inp={'points':torch.from_numpy(p),'feats':torch.from_numpy(np.concatenate([nrm,np.ones((N,1),np.float32)],1))}
for i,g in enumerate(grids):
    for k,v in g.items():
        if k!='voxel_keys': inp[k+str(i)]=torch.from_numpy(v)
t=time.time()
idx,d2,rs=ops_cpu.multi_radius_search(p,grids[0]['voxel_centers'],grids[0]['voxel_sizes'])
print('pairs',len(idx),time.time()-t, 'V0',len(rs)-1)
sc=ops_cpu.scale_compatibility(grids[0]['voxel_sizes'],r,idx,rs)
inp['aggregation_neighbors_index']=torch.from_numpy(idx); inp['aggregation_neighbors_dist']=torch.from_numpy(d2)
inp['aggregation_row_splits']=torch.from_numpy(rs); inp['aggregation_scale_compat']=torch.from_numpy(sc)
Psd={k:v.detach().clone() for k,v in sd.items()}
with torch.no_grad():
    t=time.time(); a=net.aggregate(inp); code=net.unet(a,inp); val=net.decode(torch.zeros(code.shape[0],3),code); print('ref',time.time()-t)
    t=time.time(); a2=model_cpu.aggregate(Psd,inp); code2=model_cpu.unet(Psd,a2,inp,5); val2=model_cpu.decode(Psd,torch.zeros(code2.shape[0],3),code2); print('port',time.time()-t)
print((a[0]-a2[0]).abs().max(), (a[1]-a2[1]).abs().max(), (code-code2).abs().max(), (val-val2).abs().max())
print(a[0].abs().mean(), code.abs().mean(), val.abs().mean(), val[:3])
