import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from svae_b200.hmm import hmm_inference as K
from oracle import hmm_numpy as H
Kk,T,B=[int(x) for x in (sys.argv[1:4] if len(sys.argv)>3 else (32,500,2048))]
rng=np.random.default_rng(0)
lp,lA=H.rand_hmm_natparam(Kk,rng)
g=lambda x: torch.as_tensor(x,dtype=torch.float64,device="cuda")
node=torch.randn(B,T,Kk,dtype=torch.float64,device="cuda")
nat=(g(lp),g(lA),node)
for _ in range(3): K.hmm_estep(nat)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): K.hmm_estep(nat)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("HMM K=%d T=%d B=%d: %.3f ms  %.3e seq/s  alg %.1f GB/s"%(Kk,T,B,ms,B/ms*1e3, 8*(2*T*Kk+Kk*Kk+Kk+1)*B/ms/1e6))
