"""Child process of test_gpu_encoder.py: runs the tcgen05 Encoder once and prints its error
against the torch fp32 reference (oracle/torch_port.py).  A separate process so that a pipeline
bug can at worst kill the child."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import torch_port  # noqa: E402
from ppt_b200 import encoder_pack, ops  # noqa: E402


def main():
    case, mode = sys.argv[1], int(sys.argv[2])
    sd = torch_port.make_encoder_state()
    if case == "golden":
        f = np.load(os.path.join(ROOT, "tests", "golden", "encoder_small.npz"))
        nb = torch.from_numpy(f["neighborhood"])
        ref_tok, ref_feat = torch.from_numpy(f["tokens"]), torch.from_numpy(f["features"])
    else:
        groups = int(case)
        g = torch.Generator().manual_seed(groups)
        nb = (torch.rand(1, groups, 32, 3, generator=g) - 0.5) * 0.4
        with torch.no_grad():  # chunked: the port materialises every [groups, C, 32] intermediate
            ref_feat = torch.cat([torch_port.encoder_forward(sd, c) for c in nb.split(8192, dim=1)], dim=1)
            ref_tok = torch.nn.functional.linear(ref_feat, sd["reduce_dim.weight"], sd["reduce_dim.bias"])
    blob = encoder_pack.pack_encoder(sd, mode).cuda()
    tok, feat = ops.encoder_forward(nb.cuda(), blob, mode=mode, return_features=True)
    torch.cuda.synchronize()
    tok2 = ops.encoder_forward(nb.cuda(), blob, mode=mode)  # second call: same workspace, no features
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = a.double().cpu(), b.double()
        return {"max": float((a - b).abs().max() / b.abs().max()), "rms": float((a - b).norm() / b.norm())}

    print(json.dumps({"tokens": rel(tok, ref_tok), "features": rel(feat, ref_feat),
                      "repeatable": bool(torch.equal(tok, tok2)), "finite": bool(torch.isfinite(tok).all())}))


if __name__ == "__main__":
    main()
