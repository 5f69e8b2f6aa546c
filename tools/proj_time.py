import torch, time
dev='cuda:0'
torch.backends.cuda.matmul.allow_tf32=False
for N,din,dout in ((121800,128,128),(490800,128,128),(169343,128,64),(232965,128,128)):
    x=torch.randn(N,din,device=dev)
    lin=[torch.nn.Linear(din,dout).to(dev) for _ in range(3)]
    wcat=torch.cat([l.weight for l in lin],0).contiguous(); bcat=torch.cat([l.bias for l in lin],0)
    flush=torch.empty(64*1024*1024,device=dev)
    def three():
        q=lin[0](x).reshape(N,1,dout)*0.088; k=lin[1](x).reshape(N,1,dout); v=lin[2](x).reshape(N,1,dout); return q,k,v
    def fusedcat():
        return torch.addmm(bcat,x,wcat.t())
    def tf32cat():
        torch.backends.cuda.matmul.allow_tf32=True
        r=torch.addmm(bcat,x,wcat.t())
        torch.backends.cuda.matmul.allow_tf32=False
        return r
    for name,fn in (("3xLinear fp32",three),("1 addmm [N,3d] fp32",fusedcat),("1 addmm tf32",tf32cat)):
        for _ in range(3): fn()
        ts=[]
        for _ in range(10):
            flush.fill_(1.0)
            a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        print(N,din,dout,name, f"{sorted(ts)[5]*1e3:.1f} us")

# the fused tcgen05 projection on the same shapes
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfgnn_b200.operators.projection import PackedWeights, proj_forward
for N, din, dout in ((121800, 128, 128), (490800, 128, 128), (169343, 128, 64), (232965, 128, 128)):
    x = torch.randn(N, din, device=dev)
    lin = [torch.nn.Linear(din, dout).to(dev) for _ in range(3)]
    cache = PackedWeights()
    img, bias = cache.get([l.weight for l in lin], [l.bias for l in lin])
    scale = torch.ones(3 * dout, device=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    fn = lambda: proj_forward(x, img, bias, scale, 3 * dout, dout)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    t = sorted(ts)[5]
    gb = (N * din + 3 * N * dout) * 4 / 1e9
    print(N, din, dout, f"fused tcgen05 3xTF32 qkv {t*1e3:.1f} us  ({gb / (t*1e-3):.0f} GB/s of compulsory traffic)")
