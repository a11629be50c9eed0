"""Per-phase cycle breakdown of one fps_grid_kernel launch (needs the measurement build:
python tools/build_variant.py trace fps_grid.cu -DFPS_TRACE; PPT_B200_LIB=ppt_b200/libppt_b200_trace.so)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ppt_b200 import _lib, ops  # noqa: E402

B, N, G = 128, 8192, 512
g = torch.Generator().manual_seed(1)
xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).cuda()
z = torch.zeros(B, dtype=torch.int64, device="cuda")
index = ops.spatial_index(xyz)
lib = _lib.load()
fn = lib.ppt_debug_fps_trace
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]
buf = torch.zeros(64 + 4 * 4096, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.fps(xyz, G, z, index=index)
torch.cuda.synchronize()
fn(buf.data_ptr(), None)  # reset
torch.cuda.synchronize()
ops.fps(xyz, G, z, index=index)
torch.cuda.synchronize()
fn(buf.data_ptr(), None)
torch.cuda.synchronize()
raw = buf.cpu().numpy()
phase = raw[:64].view(np.uint64)
rows = raw[64:64 + 4096].view(np.uint32)[:G - 1]
rows0 = raw[64 + 4096:64 + 8192].view(np.uint32)[:G - 1]
it = raw[64 + 8192:64 + 12288].view(np.uint32)[:G - 1]
rmax = raw[64 + 12288:].view(np.uint32)[:G - 1]
names = ["A skip test", "B row passes", "C warp argmax", "D barrier wait", "E block argmax + centre"]
tot = float(phase[:5].sum())
print("cycles per iteration (warp 0 of cloud 0): %.0f" % (tot / (G - 1)))
for n, v in zip(names, phase[:5]):
    print("  %-26s %7.1f cycles/iter  %4.1f %%" % (n, v / (G - 1), 100 * v / tot))
for lo, hi in ((0, 8), (8, 32), (32, 128), (128, 256), (256, 511)):
    print("  iterations %3d-%3d: %6.0f cycles/iter, rows touched %6.1f (warp 0: %4.1f, busiest warp: %4.1f)" %
          (lo, hi, it[lo:hi].mean(), rows[lo:hi].mean(), rows0[lo:hi].mean(), rmax[lo:hi].mean()))
print("  busiest-warp rows histogram (iterations 32-511):", np.bincount(rmax[32:], minlength=9)[:12].tolist())
