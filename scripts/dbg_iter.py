import sys, os, torch
sys.path.insert(0, "/root/repo")
from ocrl_b200 import abi, functional as F
from oracle import slot_oracle as so
def run(B,N,K,T,lanes):
    p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, 192, 192, seed=11).items()}
    g = torch.Generator().manual_seed(B*7+N)
    x = torch.randn(B,N,64,generator=g).cuda(); s0 = torch.randn(B,K,192,generator=g).cuda()
    k,v,_ = F.kv_project(x,p,kv="bf16")
    ref = F.iterate(k,v,s0,p,T,opts=abi.launch_opts(variant="pipe",strict=True))
    out = F.iterate(k,v,s0,p,T,opts=abi.launch_opts(variant="tcgen05",lanes=lanes,strict=True))
    torch.cuda.synchronize()
    es = ((out[0]-ref[0]).flatten(1).norm(dim=1)/ref[0].flatten(1).norm(dim=1)).tolist()
    print(f"B={B} N={N} K={K} T={T} lanes={lanes}: per-image slot err", [round(e,4) for e in es])
for cfg in [(1,4096,6,1,2),(1,4096,6,2,2),(1,4096,6,3,2),(3,100,5,2,2),(3,100,5,2,3),(5,1000,7,4,2),(1,512,6,2,3),(2,4096,6,3,2)]:
    run(*cfg)
