import sys, os, torch
sys.path.insert(0, '/root/repo')
import extdm_b200
from extdm_b200 import ops
dev='cuda'
BF=torch.bfloat16
x=torch.randn(1,1,4,4,64,device=dev).to(BF); y=torch.zeros_like(x); g=torch.ones(64,device=dev)
rec=ops.Recorder(record=True)
N=2000
for i in range(N):
    ops.chan_layernorm(rec, x if i%2==0 else y, g, y if i%2==0 else x)
rec.run(); torch.cuda.synchronize()
gr=torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    rec.run()
for _ in range(3): gr.replay()
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): gr.replay()
b.record(); torch.cuda.synchronize()
print("graph: %.2f us per dependent tiny-kernel node" % (a.elapsed_time(b)/5/N*1e3))
a.record()
for _ in range(5): rec.run()
b.record(); torch.cuda.synchronize()
print("eager: %.2f us per launch" % (a.elapsed_time(b)/5/N*1e3))
