"""tcgen05 building blocks (operand layouts, descriptors, bulk copy, MMA, TMEM load) against a
torch matmul.  Each variant runs in a child process under a timeout."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_umma_child.py")

# N, K, mode (0 fp16, 1 bf16, 2 bf16x3), B MN-major, A from packed image
VARIANTS = [
    (128, 64, 0, 0, 0), (128, 128, 0, 0, 0), (64, 512, 0, 0, 0), (64, 256, 0, 0, 0), (32, 128, 0, 0, 0),
    (256, 128, 0, 0, 0), (128, 128, 1, 0, 0), (128, 128, 0, 0, 1), (64, 512, 1, 0, 1), (64, 128, 2, 0, 0),
    (64, 256, 2, 0, 1), (128, 128, 0, 1, 0), (64, 512, 0, 1, 1), (64, 128, 1, 1, 0),
]


# the A operand read from TENSOR MEMORY (written there with tcgen05.st): N, K, mode, B MN-major
TMEM_A_VARIANTS = [(128, 64, 0, 0), (64, 128, 0, 0), (64, 512, 0, 0), (64, 256, 1, 0), (256, 128, 0, 0), (128, 128, 0, 1)]
VARIANTS = [v + (0,) for v in VARIANTS] + [(n, k, m, b, 0, 1) for n, k, m, b in TMEM_A_VARIANTS]


@pytest.mark.parametrize("N,K,mode,b_mn,packed,a_tmem", VARIANTS)
def test_umma_selftest(N, K, mode, b_mn, packed, a_tmem):
    try:
        out = subprocess.run([sys.executable, CHILD] + [str(v) for v in (N, K, mode, b_mn, packed, a_tmem)],
                             capture_output=True, text=True, timeout=120)
    except subprocess.TimeoutExpired:
        pytest.fail("tcgen05 self-test hung (killed after 120 s)")
    assert out.returncode == 0, out.stderr[-2000:]
    err = json.loads(out.stdout.strip().splitlines()[-1])["rel_err"]
    # operands are rounded identically on both sides; only fp32 accumulation order differs
    assert err < (2e-5 if mode == 2 else 1e-5), err
