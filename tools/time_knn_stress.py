"""Times ppt_knn_group at the 8 x 32768-point stress size (BASELINE configs[4]): index build + pruned search against
the full scan, and checks that both give the same bits."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ppt_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(99)
for B, N in ((8, 32768), (32, 16384)):
    p = torch.randn(B, N, 3, generator=g)
    xyz = (p / p.norm(dim=-1, keepdim=True)).to(dev)
    zeros = torch.zeros(B, dtype=torch.int64, device=dev)
    _, center = ops.fps(xyz, 512, zeros, return_centers=True)
    res = {}
    for name, ix in (("pruned+build", ops.AUTO), ("full", None)):
        ev = []
        for i in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nb, idx = ops.knn_group(xyz, center, 32, return_idx=True, index=ix)
            b.record()
            if i >= 5:
                ev.append((a, b))
        torch.cuda.synchronize()
        res[name] = (statistics.mean(a.elapsed_time(b) for a, b in ev), nb, idx)
    same = torch.equal(res["full"][2], res["pruned+build"][2]) and torch.equal(res["full"][1], res["pruned+build"][1])
    print("knn_group %dx%d: pruned+build %.4f ms, full scan %.4f ms, identical %s" %
          (B, N, res["pruned+build"][0], res["full"][0], same))
