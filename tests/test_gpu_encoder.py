"""tcgen05 patch Encoder + reduce_dim against the torch fp32 reference and the reference-generated
fixture.  Tolerances are norm-relative: max|d|/max|ref| and rms(d)/rms(ref)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_encoder_child.py")
# mode -> tolerance: fp16 operands 1e-3 (BASELINE "fast" bar), bf16 operands reported at 8e-3
# (SURVEY.md F15 measured ~3e-3), scaled fp16 hi/lo split: fp32-parity bar 1e-5 (measured 2.4e-6 rms / 4e-6 max against the fp32 torch
# reference, whose own distance from an fp64 evaluation is ~1e-6)
TOL = {0: 1e-3, 1: 8e-3, 2: 1e-5}


def run_child(case, mode, env=None):
    try:
        out = subprocess.run([sys.executable, CHILD, str(case), str(mode)], capture_output=True, text=True,
                             timeout=300, env=dict(os.environ, **(env or {})))
    except subprocess.TimeoutExpired:
        pytest.fail("encoder child hung (killed after 300 s)")
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("case", ["golden", "4", "1", "131", "1500", "16384", "65536"])
def test_encoder_tokens(case, mode):
    """"65536" = BASELINE configs[1]: 128 clouds x 512 groups, the size bench.py times."""
    r = run_child(case, mode)
    assert r["finite"] and r["repeatable"], r
    for key in ("tokens", "features"):
        assert r[key]["max"] <= TOL[mode], (key, r)
        assert r[key]["rms"] <= TOL[mode], (key, r)
