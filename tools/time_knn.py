"""Times ppt_knn_group at BASELINE configs[1] (128 clouds x 8192 points, 512 queries, k = 32) for A/B builds
(PPT_B200_LIB selects the library)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ppt_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1234)
xyz = (torch.rand(128, 8192, 3, generator=g) * 2 - 1).to(dev)
zeros = torch.zeros(128, dtype=torch.int64, device=dev)
index = ops.spatial_index(xyz).clone()
_, center = ops.fps(xyz, 512, zeros, return_centers=True, index=index)
ev = []
for i in range(25):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    nb = ops.knn_group(xyz, center, 32, index=index)
    b.record()
    if i >= 5:
        ev.append((a, b))
torch.cuda.synchronize()
print(os.path.basename(os.environ.get("PPT_B200_LIB", "libppt_b200.so")), "knn_group ms %.4f" % statistics.mean(a.elapsed_time(b) for a, b in ev),
      "checksum %.6f" % float(nb.double().abs().sum()))
