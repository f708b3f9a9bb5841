import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lds_numpy as O
from svae_b200.lds import lds_inference as K
n,T,B,S=10,200,4096,1
dev=torch.device("cuda")
rng=np.random.default_rng(0)
init,pair=O.rand_lds_natparam(n,rng)
g=lambda x: torch.as_tensor(np.asarray(x),dtype=torch.float64,device=dev)
gi,gp=tuple(g(x) for x in init),tuple(g(x) for x in pair)
h_nJ=(-0.5*torch.nn.functional.softplus(torch.randn(B,T,n,dtype=torch.float64))).pin_memory(); h_nh=torch.randn(B,T,n,dtype=torch.float64).pin_memory()
out_host=dict(samples=torch.empty((B,T,S,n),dtype=torch.float64).pin_memory(),E_node_xx=torch.empty((B,T,n),dtype=torch.float64).pin_memory(),
              E_node_x=torch.empty((B,T,n),dtype=torch.float64).pin_memory(),lognorm=torch.empty((B,),dtype=torch.float64).pin_memory())
scheds=[1,3,4,6,8]+[K.ramp_schedule(B,f,gr,mx) for f in (128,256,512) for gr in (1.5,2.0) for mx in (1024,2048)]
for chunks in scheds:
    for _ in range(3): K.lds_estep_host((gi,gp),(h_nJ,h_nh),S,out_host,chunks=chunks)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): K.lds_estep_host((gi,gp),(h_nJ,h_nh),S,out_host,chunks=chunks)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/8
    print("chunks",chunks,"%.2f ms  %.0f seq/s"%(ms,B/ms*1e3),flush=True)
