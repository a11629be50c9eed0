"""Launch-level timing of the fused feature-propagation module (row f4) at the part-seg shape.
    ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/fp_launches.csv python tools/fp_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ppt_b200 import pointnet2  # noqa: E402

dev = torch.device("cuda", 0)
fp = pointnet2.PointNetFeaturePropagation(384 + 19, [1536, 384]).to(dev).eval()
for p in fp.parameters():
    p.requires_grad_(False)
B = 32
x1, x2 = torch.randn(B, 3, 2048, device=dev), torch.randn(B, 3, 512, device=dev)
p1, p2 = torch.randn(B, 19, 2048, device=dev), torch.randn(B, 384, 512, device=dev)
with torch.no_grad():
    for _ in range(3):
        fp(x1, x2, p1, p2)
torch.cuda.synchronize()
