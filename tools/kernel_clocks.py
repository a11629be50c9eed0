"""SM clock INSIDE each kernel of the tokenizer step (ops.ClockProbe): every kernel is launched back to back for a
few milliseconds while a one-thread probe kernel samples (globaltimer, clock64) on a side stream.
    python tools/kernel_clocks.py            (on a B200; prints one JSON line)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import torch_port  # noqa: E402
from ppt_b200 import _lib, encoder_pack, ops  # noqa: E402

B, N, G = 128, 8192, 512
dev = torch.device("cuda", 0)
xyz = torch.rand(B, N, 3, device=dev) * 2 - 1
zeros = torch.zeros(B, dtype=torch.int64, device=dev)
index = ops.spatial_index(xyz)
_, center = ops.fps(xyz, G, zeros, return_centers=True, index=index)
nb = ops.knn_group(xyz, center, 32, index=index)
blob = encoder_pack.pack_encoder(torch_port.make_encoder_state(), 0).to(dev)
ops.encoder_forward(nb, blob, mode=0)
lib = _lib.load()
ws = ops._workspace((dev, "encoder"), lib.ppt_encoder_workspace_bytes(B * G, 0))
tok = torch.empty(B, G, 384, device=dev)
st = torch.cuda.current_stream().cuda_stream


def phase(bit):
    return lambda: _lib.check(lib.ppt_encoder_forward_phases(nb.data_ptr(), blob.data_ptr(), ws.data_ptr(), None,
                                                             tok.data_ptr(), B * G, 0, 1 << bit, st), "phase")


kernels = {
    "fps": lambda: ops.fps(xyz, G, zeros, return_centers=True, index=index),
    "knn_group": lambda: ops.knn_group(xyz, center, 32, index=index),
    "stage1": phase(0), "stage2": phase(2), "group_linear_tokens": phase(3),
}
out = {}
for name, fn in kernels.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe = ops.ClockProbe(dev, duration_ms=12.0, period_us=20.0).start()
    e0.record()
    reps = 0
    while reps < 400:
        fn()
        reps += 1
        if reps % 8 == 0:
            e1.record()
            e1.synchronize()
            if e0.elapsed_time(e1) > 14.0:
                break
    torch.cuda.synchronize()
    m = probe.mhz()
    out[name] = {"sm_mhz_mean": round(m[0], 1), "sm_mhz_min_200us": round(m[1], 1), "sm_mhz_max_200us": round(m[2], 1)}
print(json.dumps({"clocks_inside_kernels": out}))
