import torch, torch.distributed as dist, os, torch.multiprocessing as mp
def w(rank):
    os.environ['MASTER_ADDR']='127.0.0.1'; os.environ['MASTER_PORT']='29533'
    dist.init_process_group('gloo', rank=rank, world_size=2)
    x=torch.arange(6, dtype=torch.float32)+10*rank
    out=torch.empty(12)
    dist.all_gather_into_tensor(out, x)
    try:
        o=torch.empty(4); dist.all_to_all_single(o, torch.arange(4.)+rank, [2,2],[2,2]); a2a='ok'
    except Exception as e: a2a='no: '+str(e)[:60]
    full=torch.zeros(12); full[rank*6:(rank+1)*6]=x
    try:
        dist.all_gather_into_tensor(full, full[rank*6:(rank+1)*6]); inplace='ok'
    except Exception as e: inplace='no: '+str(e)[:80]
    if rank==0: print(out.tolist(), a2a, inplace, full.tolist())
    dist.destroy_process_group()
if __name__=='__main__':
    mp.spawn(w, nprocs=2)
