"""Child process of test_gpu_front_end.py: pos_embed + token assembly (ppt_tokenizer_forward) against the
reference-generated fixture or the torch restatement; prints norm-relative errors as one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import torch_port  # noqa: E402
from ppt_b200 import encoder_pack, ops  # noqa: E402
from ppt_b200.tokenizer import PointTokenizer  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return {"max": float((a - b).abs().max() / b.abs().max()), "rms": float((a - b).norm() / b.norm())}


def main():
    case, mode = sys.argv[1], int(sys.argv[2])
    sd = torch_port.make_encoder_state()
    front = torch_port.make_front_end_state()
    out = {}
    if case in ("golden", "module"):
        f = np.load(os.path.join(ROOT, "tests", "golden", "front_end_small.npz"))
        nb, center = torch.from_numpy(f["neighborhood"]), torch.from_numpy(f["center"])
        ref_x, ref_pos = torch.from_numpy(f["x"]), torch.from_numpy(f["pos"])
    else:
        B, G = (int(v) for v in case.split("x"))
        g = torch.Generator().manual_seed(B * 1000 + G)
        nb = (torch.rand(B, G, 32, 3, generator=g) - 0.5) * 0.4
        center = torch.rand(B, G, 3, generator=g) * 2 - 1
        with torch.no_grad():
            ref_x, ref_pos = torch_port.assemble_forward(front, torch_port.tokens_forward(sd, nb.reshape(1, B * G, 32, 3))
                                                         .reshape(B, G, 384), center)
    if case == "module":
        # the whole front end from raw points through the nn.Module (FPS start pinned to 0 like the fixture)
        prec = {0: "fp16", 1: "bf16", 2: "fp32"}[mode]
        tok = PointTokenizer(num_group=64, group_size=32, precision=prec).load_reference_state(sd)
        tok.load_front_end_state(front)
        tok = tok.cuda().eval()
        tok.start_idx = 0
        x, pos, ct = tok.forward_assembled(torch.from_numpy(f["xyz"]).cuda())
        out["center_equal"] = bool(torch.equal(ct.cpu(), center))
    else:
        blob = encoder_pack.pack_encoder(sd, mode).cuda()
        pblob = encoder_pack.pack_pos_embed({k[len("pos_embed."):]: v for k, v in front.items() if k.startswith("pos_embed.")},
                                            front["cls_token"], front["cls_pos"], mode).cuda()
        x, pos = ops.tokenizer_forward(nb.cuda(), center.cuda(), blob, pblob, mode=mode)
        _, pos_only = ops.tokenizer_forward(None, center.cuda(), None, pblob, mode=mode, want_x=False)
        out["pos_only_equal"] = bool(torch.equal(pos_only, pos))
    torch.cuda.synchronize()
    out.update({"x": rel(x, ref_x), "pos": rel(pos, ref_pos),
                "cls_rows_exact": bool(torch.equal(x[:, 0].cpu(), ref_x[:, 0]) and torch.equal(pos[:, 0].cpu(), ref_pos[:, 0])),
                "finite": bool(torch.isfinite(x).all() and torch.isfinite(pos).all())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
