import torch, sys
sys.path.insert(0, ".")
from dynamic_multiview_3d_b200 import functional as F
B,H=64,224
g=torch.Generator(device="cuda").manual_seed(1)
d=torch.rand((B,H,H,3),device="cuda",generator=g); fl=((torch.rand((B,H,H,2),device="cuda",generator=g)-0.5)*6); t=torch.rand((B,H,H,3),device="cuda",generator=g)
big=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
def fused():
    f=fl.detach().requires_grad_(True); l,gen=F.flow_resample_loss(d,f,t,"l2",unit_upstream=True); l.backward()
def three():
    f=fl.detach().requires_grad_(True); gen=F.flow_resampler(d,f); l=F.reconstruction_loss(gen,t,"l2",unit_upstream=True); l.backward()
for name,fn in (("three kernels",three),("fused",fused)):
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    ts=[]
    for i in range(20):
        big.zero_()   # flush L2
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    ts.sort(); print(name, "median %.1f us (L2 flushed between replays)"%ts[len(ts)//2])
