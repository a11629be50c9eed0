"""Set-abstraction shared MLP + max-pool on the tensor cores: the modules of
ppt_b200.pointnet2 in eval mode against outputs recorded from the unmodified reference modules
(models/pointnet2/pointnet2_utils.py:161-266) with the same seeded weights.  Child process: a pipeline bug can at
worst kill the child."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_sa_mlp_child.py")
# three chained layers with 16-bit operands.  fp16: the token path's 1e-3 (measured 2.9e-4 .. 4.6e-4, which IS the
# arithmetic floor: tests/test_host_cpu.py emulates the same rounding on the CPU and lands on 4.56e-4 for ssg2)
TOL = {"fp16": 1e-3, "bf16": 1.5e-2}


@pytest.mark.parametrize("single_kernel", ["1", "0"])  # activations in shared memory / per-layer kernels (PPT_SA_PER_LAYER)
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("case", ["ssg2", "ssg3", "msg1"])
def test_sa_mlp_modules(case, precision, single_kernel):
    try:
        out = subprocess.run([sys.executable, CHILD, case, precision, single_kernel], capture_output=True, text=True,
                             timeout=300)
    except subprocess.TimeoutExpired:
        pytest.fail("sa_mlp child hung (killed after 300 s)")
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["fused"], "the tensor-core path must have run"
    assert r["finite"] and r["geometry_equal"], r
    assert r["max"] <= TOL[precision] and r["rms"] <= TOL[precision], r
    assert r["unfused_max"] <= 1e-4, r          # the torch layer stack on the same kernels' geometry
    assert r["train_mode_unfused"], r
