import os, sys, statistics
sys.path.insert(0, os.getcwd())
import torch
from ppt_b200 import ops
dev=torch.device("cuda")
g=torch.Generator().manual_seed(1)
x=torch.randn(8,32768,3,generator=g); x=(x/x.norm(dim=-1,keepdim=True)).to(dev)
z=torch.zeros(8,dtype=torch.int64,device=dev)
ev=[]
for i in range(12):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); idx=ops.fps(x,512,z); b.record()
    if i>=3: ev.append((a,b))
torch.cuda.synchronize()
print(os.path.basename(os.environ.get("PPT_B200_LIB","default")), "fps 8x32768 ms %.4f"%statistics.mean(a.elapsed_time(b) for a,b in ev), int(idx.sum()))
