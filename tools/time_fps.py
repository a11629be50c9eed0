"""Times ppt_fps (B=128, N=8192, G=512) with and without the spatial index and checks they agree (GPU box).
Clocks are warmed first (a cold GPU runs the first milliseconds well below its boost clock); every launch has its own
event pair, as in bench.py; several distinct batches rotate."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ppt_b200 import ops  # noqa: E402

B, N, G = 128, 8192, 512
kind = sys.argv[1] if len(sys.argv) > 1 else "U"
g = torch.Generator().manual_seed(1)
batches = []
for _ in range(4):
    if kind == "U":
        batches.append((torch.rand(B, N, 3, generator=g) * 2 - 1).cuda())
    else:
        p = torch.randn(B, N, 3, generator=g)
        batches.append((p / p.norm(dim=-1, keepdim=True)).cuda())
z = torch.zeros(B, dtype=torch.int64, device="cuda")
a = torch.randn(4096, 4096, device="cuda")
for _ in range(60):  # ~0.1 s of work: boost clocks
    a @ a
torch.cuda.synchronize()


def timed(fn, reps=24):
    ev = []
    for i in range(reps + 4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x = batches[i % 4]
        index = ops.spatial_index(x)
        e0.record()
        fn(x, index)
        e1.record()
        if i >= 4:
            ev.append((e0, e1))
    torch.cuda.synchronize()
    t = [e0.elapsed_time(e1) * 1e3 for e0, e1 in ev]
    return statistics.mean(t), min(t)


x0 = batches[0]
same = torch.equal(ops.fps(x0, G, z, index=None), ops.fps(x0, G, z, index=ops.spatial_index(x0)))
pm, pn = timed(lambda x, ix: ops.fps(x, G, z, return_centers=True, index=None))
gm, gn = timed(lambda x, ix: ops.fps(x, G, z, return_centers=True, index=ix))
print("kind=%s equal=%s plain_us=%.1f (min %.1f) grid_us=%.1f (min %.1f)" % (kind, same, pm, pn, gm, gn))
