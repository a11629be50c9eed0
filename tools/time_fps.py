"""Times ppt_fps (B=128, N=8192, G=512) with and without the spatial index and checks they agree (GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ppt_b200 import ops  # noqa: E402

B, N, G = 128, 8192, 512
kind = sys.argv[1] if len(sys.argv) > 1 else "U"
g = torch.Generator().manual_seed(1)
if kind == "U":
    xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).cuda()
else:
    p = torch.randn(B, N, 3, generator=g)
    xyz = (p / p.norm(dim=-1, keepdim=True)).cuda()
z = torch.zeros(B, dtype=torch.int64, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


index = ops.spatial_index(xyz)
plain = ops.fps(xyz, G, z, index=None)
grid = ops.fps(xyz, G, z, index=index)
print("warps=%s kind=%s equal=%s plain_us=%.1f grid_us=%.1f index_us=%.1f" % (
    "16 warps", kind, bool(torch.equal(plain, grid)),
    timed(lambda: ops.fps(xyz, G, z, index=None)), timed(lambda: ops.fps(xyz, G, z, index=index)),
    timed(lambda: ops.spatial_index(xyz))))
