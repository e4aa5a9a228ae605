#!/bin/bash
# scratch: phase profile of the tcgen05 STW kernel at the KTH level-0 shape
EXTDM_STW_PROF=1 timeout 300 python - <<'PY' 2>&1 | tail -8
import torch, sys
sys.path.insert(0, ".")
import extdm_b200
from extdm_b200 import ops
from extdm_b200.unet import _rope_tables
torch.manual_seed(0)
B,T,H,W,C=32,30,32,32,64
dev="cuda"
x=torch.randn(B,T,H,W,C,device=dev).bfloat16(); y=torch.empty_like(x)
g=torch.ones(C,device=dev); wqkv=(torch.randn(384,C,device=dev)*0.1).bfloat16(); wp=(torch.randn(C,128,device=dev)*0.1).bfloat16()
pb=torch.zeros(C,device=dev); tbl=torch.randn(343,8,device=dev)*0.1
rc,rs=_rope_tables(64,16,dev)
for shift in ((0,0,0),(2,2,2)):
    for _ in range(2):
        ops.stw_fused(ops.IMMEDIATE,x,y,g,wqkv,wp,pb,tbl,rc,rs,8,16,(4,4,4),shift)
torch.cuda.synchronize()
PY
