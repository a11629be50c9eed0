"""Where the MMA-issuing warp of encoder_stage2_ta_kernel spends its cycles (needs the -DTA_TRACE build:
python tools/build_variant.py tatrace encoder.cu -DTA_TRACE; PPT_B200_LIB=ppt_b200/libppt_b200_tatrace.so)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from ppt_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
tok = bench.make_tokenizer("fp16").to(dev)
xyz = bench.make_host_batches(0, 1, device=dev)[0].to(dev)
zeros = torch.zeros(xyz.shape[0], dtype=torch.int64, device=dev)
index = ops.spatial_index(xyz)
_, center = ops.fps(xyz, bench.N_GROUP, zeros, return_centers=True, index=index)
nb = ops.knn_group(xyz, center, bench.GROUP_SIZE, index=index)
blob, mode = tok.encoder._blob(dev)
for _ in range(3):
    ops.encoder_forward(nb, blob, mode=mode)
torch.cuda.synchronize()
fn = _lib.load().ppt_debug_ta_trace
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]
buf = torch.zeros(8, dtype=torch.int64, device=dev)
fn(buf.data_ptr(), None)
torch.cuda.synchronize()
acc = buf.cpu().numpy().astype(np.float64)
tiles = (128 * 512 * 32 // 128 + 147) // 148
names = ["wait h1_ready", "wait empty_m (C_0: the max warps' release)", "wait empty_r (U_j: accumulator free)",
         "issue 8 MMAs + commits (U_j)", "-", "wait slot_ready (C_j: h3 chunk written)", "wait ring full",
         "issue 4 MMAs + commit (C_j, one output unit)"]
print("cycles per tile in the MMA warp of CTA 0 (%d tiles): total %.0f" % (tiles, acc.sum() / tiles))
for n, v in zip(names, acc):
    print("  %-52s %8.0f" % (n, v / tiles))
