import numpy as np, time, sys
sys.path.insert(0,'/root/repo')
from oracle import reflib, geomlib
def check(p,r,mn,mx,levels=5,md=21):
    a=reflib.RefOctree(p,r,mn,mx,max_depth=md); b=geomlib.PortOctree(p,r,mn,mx,max_depth=md)
    assert np.array_equal(a.nodes(),b.nodes()), (len(a.nodes()),len(b.nodes()))
    assert np.array_equal(a.leaves(),b.leaves())
    for x,y in zip(a.params(),b.params()): assert np.array_equal(x,y)
    for allv in (True,False):
        ga=a.grids(levels,allv); gb=b.grids(levels,allv)
        for l,(x,y) in enumerate(zip(ga,gb)):
            assert set(x)==set(y),(l,set(x)^set(y))
            for k in x: assert np.array_equal(x[k],y[k]),(l,k)
    da=a.dual_vertex_indices(); db=b.dual_vertex_indices()
    assert np.array_equal(da,db)
    g=a.grids(1,True)[0]; c=g['voxel_centers']; s=g['voxel_sizes']
    rng=np.random.default_rng(5)
    dist=np.linalg.norm(c,axis=1)-0.7+0.05*np.sin(9*c[:,0])
    vals=np.stack([dist,np.abs(dist)/s*rng.uniform(0.5,1.5,len(s))],1).astype(np.float32)
    m=reflib.create_triangle_mesh(vals,da,c,1.0)
    v,d=geomlib.contour_vertices(vals,da,c,1.0)
    assert np.array_equal(m['vertices'][:len(v)],v), np.abs(m['vertices'][:len(v)]-v).max()
    return len(a.leaves()),len(da),len(v),len(m['vertices'])
rng=np.random.default_rng(0)
N=100000
p=rng.standard_normal((N,3)).astype(np.float32); p/=np.linalg.norm(p,axis=1,keepdims=True)
r=np.full(N,np.sqrt(96.0/N),np.float32)
print(check(p,r,p.min(0)-0.1,p.max(0)+0.1))
print('nomargin',check(p,r,p.min(0),p.max(0)))
for seed in range(3):
    rng=np.random.default_rng(seed)
    N=200000
    p=(rng.standard_normal((N,3))*0.25).astype(np.float32)
    d=np.linalg.norm(p,axis=1)
    r=(0.002*np.exp(6*d)*2**rng.uniform(0,2,N)).astype(np.float32)
    print(check(p,r,p.min(0)-0.1,p.max(0)+0.1))
    print(check(p,r,p.min(0)-0.1,p.max(0)+0.1,levels=3,md=7))
