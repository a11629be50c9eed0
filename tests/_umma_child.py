"""Child process of test_gpu_umma.py: one tcgen05 self-test GEMM variant per invocation, so a
mis-programmed descriptor can at worst kill this process (the parent enforces a timeout)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ppt_b200 import encoder_pack, ops  # noqa: E402


def main():
    N, K, mode, b_mn, packed = (int(v) for v in sys.argv[1:6])
    a_tmem = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(N, K, generator=g)
    dt = encoder_pack.operand_dtype(mode)
    ap = None
    if packed:
        ap = encoder_pack.pack_kmajor(a, dt, encoder_pack.split_of(mode)).cuda()
    d = ops.selftest_umma(a.cuda(), b.cuda(), mode=mode, b_mn_major=bool(b_mn), a_packed=ap, a_in_tmem=bool(a_tmem))
    torch.cuda.synchronize()
    if mode == ops.ENC_FP16X3:
        ref = a.double() @ b.double().T
    else:
        ref = a.to(dt).double() @ b.to(dt).double().T
    err = float((d.cpu().double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({"rel_err": err}))


if __name__ == "__main__":
    main()
