"""Child process of test_gpu_umma.py: one CTA-pair (cta_group::2) self-test GEMM variant per invocation."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ppt_b200 import encoder_pack, ops  # noqa: E402


def main():
    N, K, mode, b_mn = (int(v) for v in sys.argv[1:5])
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    a = torch.randn(256, K, generator=g)
    b = torch.randn(N, K, generator=g)
    dt = encoder_pack.operand_dtype(mode)
    d = ops.selftest_umma_pair(a.cuda(), b.cuda(), mode=mode, b_mn_major=bool(b_mn))
    torch.cuda.synchronize()
    ref = a.to(dt).double() @ b.to(dt).double().T
    err = (d.cpu().double() - ref).abs()
    print(json.dumps({"rel_err": float(err.max() / ref.abs().max()),
                      "rel_err_top": float(err[:128].max() / ref.abs().max()),
                      "rel_err_bottom": float(err[128:].max() / ref.abs().max()),
                      "rel_err_left": float(err[:, : N // 2].max() / ref.abs().max()),
                      "rel_err_right": float(err[:, N // 2:].max() / ref.abs().max())}))


if __name__ == "__main__":
    main()
