"""Launch-level timing of one fused set-abstraction level (row f1): SSG level 2 at cfg 3 size.
    ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/sa_launches.csv python tools/sa_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import torch_port  # noqa: E402
from ppt_b200 import pointnet2  # noqa: E402

dev = torch.device("cuda", 0)
sa2 = pointnet2.PointNetSetAbstraction(128, 0.4, 64, 131, [128, 128, 256], False).to(dev).eval()
sa2.load_state_dict({k: v.to(dev) for k, v in torch_port.make_sa_state(131, [128, 128, 256], 11).items()}, strict=False)
sa2.start_idx = 0
pn = torch.randn(32, 512, 3, device=dev)
sx = (pn / pn.norm(dim=-1, keepdim=True)).permute(0, 2, 1).contiguous()
sf = torch.randn(32, 128, 512, device=dev)
with torch.no_grad():
    for _ in range(3):
        sa2(sx, sf)
torch.cuda.synchronize()
