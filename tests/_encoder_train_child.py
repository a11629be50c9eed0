"""Child process of test_gpu_encoder_train.py: the Encoder under .train() (batch-statistics BatchNorm, running-stat
update) through ppt_b200.pointbert.Encoder against the reference fixture / the torch restatement."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import torch_port  # noqa: E402
from ppt_b200 import pointbert  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return {"max": float((a - b).abs().max() / b.abs().max()), "rms": float((a - b).norm() / b.norm())}


def main():
    case, mode = sys.argv[1], int(sys.argv[2])
    sd = torch_port.make_encoder_state()
    steps = 1
    if case == "golden":
        f = np.load(os.path.join(ROOT, "tests", "golden", "encoder_train_small.npz"))
        nbs = [torch.from_numpy(f["neighborhood"])]
    else:
        groups, steps = (int(v) for v in case.split("x"))
        g = torch.Generator().manual_seed(groups)
        # off-centre, anisotropic patches: non-trivial batch mean / covariance
        nbs = [(torch.rand(1, groups, 32, 3, generator=g) - 0.3) * torch.tensor([0.4, 0.2, 0.6]) for _ in range(steps)]
    enc = pointbert.Encoder(256, precision={0: "fp16", 1: "bf16", 2: "fp32"}[mode])
    enc.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    reduce_dim = torch.nn.Linear(256, 384)
    reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
    enc.attach_reduce_dim(reduce_dim.cuda())
    enc = enc.cuda().train()
    for p in enc.parameters():
        p.requires_grad_(False)          # PPT freezes the Encoder (models/ULIP_models.py:505)
    for p in reduce_dim.parameters():
        p.requires_grad_(False)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    out = {}
    for step, nb in enumerate(nbs):
        with torch.no_grad():
            ref_feat, stats = torch_port.encoder_forward_train(ref_sd, nb)
        ref_sd.update(stats)
        feat = enc(nb.cuda())
        tok = enc.forward_tokens(nb.cuda()) if step == len(nbs) - 1 else None  # a second train forward (stats move on)
    torch.cuda.synchronize()
    if case == "golden":
        ref_feat = torch.from_numpy(f["features"])
        stats = {k[len("after."):]: torch.from_numpy(f[k]) for k in f.files if k.startswith("after.") and "running" in k}
    out["features"] = rel(feat, ref_feat)
    got = enc.state_dict()
    # running stats after the LAST enc(...) call are compared one update back: forward_tokens made one more
    with torch.no_grad():
        _, stats2 = torch_port.encoder_forward_train(ref_sd if case != "golden" else
                                                     {**sd, **stats}, nbs[-1])
        ref_tok = torch.nn.functional.linear(
            torch_port.encoder_forward_train({**sd, **stats} if case == "golden" else ref_sd, nbs[-1])[0],
            sd["reduce_dim.weight"], sd["reduce_dim.bias"])
    out["tokens"] = rel(tok, ref_tok)
    out["stats"] = max(rel(got[k], v)["max"] for k, v in stats2.items())
    out["nbt"] = int(got["first_conv.1.num_batches_tracked"]), int(got["second_conv.1.num_batches_tracked"])
    out["expected_nbt"] = steps + 1
    # back to eval(): the folded blob must pick up the running statistics the kernels wrote
    enc.eval()
    ev = enc(nbs[-1].cuda())
    with torch.no_grad():
        ref_ev = torch_port.encoder_forward({**sd, **stats2}, nbs[-1])
    out["eval_after_train"] = rel(ev, ref_ev)
    # anything that needs a gradient keeps the torch layers
    enc.train()
    for p in enc.parameters():
        p.requires_grad_(True)
    y = enc(nbs[-1].cuda())
    out["grad_path"] = bool(y.requires_grad)
    out["finite"] = bool(torch.isfinite(feat).all() and torch.isfinite(tok).all())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
