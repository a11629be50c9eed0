"""Timing probe for the HBM-bound gathers (group_concat cfg 3 sa2, three_interpolate cfg 4) next to torch's fill / copy
of the same number of bytes (what the memory system gives a write-only and a read+write stream)."""
import os
import sys
import statistics

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppt_b200 import ops  # noqa: E402


def t(fn, n=20):
    for _ in range(3):
        fn()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in ev)


dev = torch.device("cuda")
B, N, S, K, D = 64, 512, 128, 64, 128
xyz = torch.randn(B, N, 3, device=dev)
ctr = xyz[:, :S].contiguous()
f = torch.randn(B, N, D, device=dev)
idx = torch.randint(0, N, (B, S, K), device=dev)
out = ops.group_concat(xyz, ctr, f, idx)
nbytes = out.numel() * 4
src = torch.empty_like(out)
print("group_concat  %.1f MB  ms %.4f" % (nbytes / 1e6, t(lambda: ops.group_concat(xyz, ctr, f, idx))))
print("zero_ (write) ms %.4f -> %.0f GB/s" % ((z := t(lambda: out.zero_())), nbytes / z / 1e6))
print("copy_ (r+w)   ms %.4f -> %.0f GB/s" % ((c := t(lambda: out.copy_(src))), 2 * nbytes / c / 1e6))
g128 = ops.gather(f, idx)
print("gather C=128  %.1f MB ms %.4f" % (g128.numel() * 4 / 1e6, t(lambda: ops.gather(f, idx))))
B, N, S, D = 64, 2048, 512, 384
u = torch.randn(B, N, 3, device=dev)
k = u[:, :S].contiguous()
feats = torch.randn(B, S, D, device=dev)
dist, i3 = ops.three_nn(u, k)
o = ops.three_interpolate(feats, i3, dist)
print("three_interp  %.1f MB ms %.4f" % (o.numel() * 4 / 1e6, t(lambda: ops.three_interpolate(feats, i3, dist))))
print("zero_ same    ms %.4f" % t(lambda: o.zero_()))
