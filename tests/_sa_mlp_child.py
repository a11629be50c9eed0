"""Child of test_gpu_sa_mlp.py: one set-abstraction module (fused shared MLP) vs the reference fixture."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from oracle import torch_port  # noqa: E402
from ppt_b200 import ops, pointnet2  # noqa: E402


def main():
    case, precision = sys.argv[1], sys.argv[2]
    f = np.load(os.path.join(ROOT, "tests", "golden", "sa_mlp.npz"))
    calls = []
    real = ops.sa_mlp_forward
    ops.sa_mlp_forward = lambda *a, **k: calls.append(1) or real(*a, **k)
    if case == "ssg2":
        mod = pointnet2.PointNetSetAbstraction(128, 0.4, 64, 131, [128, 128, 256], False)
        mod.load_state_dict(torch_port.make_sa_state(131, [128, 128, 256], 11), strict=False)
        args = (torch.from_numpy(f["ssg2.xyz"]).permute(0, 2, 1).cuda(), torch.from_numpy(f["ssg2.feats"]).permute(0, 2, 1).cuda())
        ref, ref_xyz = f["ssg2.out"], f["ssg2.new_xyz"]
    elif case == "ssg3":
        mod = pointnet2.PointNetSetAbstraction(None, None, None, 259, [256, 512, 1024], True)
        mod.load_state_dict(torch_port.make_sa_state(259, [256, 512, 1024], 12), strict=False)
        args = (torch.from_numpy(f["ssg3.xyz"]).permute(0, 2, 1).cuda(), torch.from_numpy(f["ssg3.feats"]).permute(0, 2, 1).cuda())
        ref, ref_xyz = f["ssg3.out"], None
    else:
        widths = [[32, 32, 64], [64, 64, 128], [64, 96, 128]]
        mod = pointnet2.PointNetSetAbstractionMsg(128, [0.1, 0.2, 0.4], [16, 32, 128], 0, widths)
        sd = {}
        for j, w in enumerate(widths):
            sd.update(torch_port.make_sa_state(3, w, 20 + j, "conv_blocks.%d." % j, "bn_blocks.%d." % j))
        mod.load_state_dict(sd, strict=False)
        args = (torch.from_numpy(f["msg1.xyz"]).permute(0, 2, 1).cuda(), None)
        ref, ref_xyz = f["msg1.out"], None
    mod = mod.cuda().eval()
    mod.start_idx = 0
    mod.ppt_precision = precision
    mod.ppt_sa_per_layer = len(sys.argv) > 3 and sys.argv[3] == "0"
    with torch.no_grad():
        new_xyz, out = mod(*args)
        fused_calls = len(calls)
        # the same module with the fused path disabled: the module's own torch layers on the kernels' geometry
        ops.sa_mlp_supported_real, ops.sa_mlp_supported = ops.sa_mlp_supported, lambda *a: False
        _, plain = mod(*args)
        ops.sa_mlp_supported = ops.sa_mlp_supported_real
        mod.train()
        n0 = len(calls)
        mod(*args)
        train_unfused = len(calls) == n0
    torch.cuda.synchronize()
    ref_t = torch.from_numpy(ref).double()
    d = out.double().cpu() - ref_t
    res = {"fused": fused_calls > 0, "finite": bool(torch.isfinite(out).all()),
           "max": float(d.abs().max() / ref_t.abs().max()), "rms": float(d.norm() / ref_t.norm()),
           "unfused_max": float((plain.double().cpu() - ref_t).abs().max() / ref_t.abs().max()),
           "geometry_equal": True if ref_xyz is None else bool(np.array_equal(new_xyz.cpu().numpy(), ref_xyz)),
           "train_mode_unfused": train_unfused}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
