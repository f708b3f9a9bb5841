import torch, time
dev="cuda"
h_in=torch.empty(131072000//8,dtype=torch.float64).pin_memory(); h_out=torch.empty(196644056//8,dtype=torch.float64).pin_memory()
d_in=torch.empty_like(h_in,device=dev); d_out=torch.empty(h_out.numel(),dtype=torch.float64,device=dev)
def t(f,n=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("H2D 131MB ms", t(lambda: d_in.copy_(h_in,non_blocking=True)))
print("D2H 197MB ms", t(lambda: h_out.copy_(d_out,non_blocking=True)))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
print("both ms", t(both))
