"""pos_embed + token assembly: x = cat(cls_token, tokens), pos = cat(cls_pos,
pos_embed(center)) -- the arguments of self.blocks(x, pos), models/pointbert/point_encoder.py:241-249 -- against
the fixture recorded from the unmodified reference PointTransformer and against the torch restatement."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_front_end_child.py")
TOL = {0: 1e-3, 1: 8e-3, 2: 1e-5}  # as for the tokens (tests/test_gpu_encoder.py)


def run_child(case, mode):
    try:
        out = subprocess.run([sys.executable, CHILD, str(case), str(mode)], capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        pytest.fail("front-end child hung (killed after 300 s)")
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("case", ["golden", "module", "5x37", "1x32", "3x512", "130x64"])
def test_token_assembly(case, mode):
    r = run_child(case, mode)
    assert r["finite"] and r["cls_rows_exact"], r
    assert r.get("pos_only_equal", True) and r.get("center_equal", True), r
    for key in ("x", "pos"):
        assert r[key]["max"] <= TOL[mode], (key, r)
        assert r[key]["rms"] <= TOL[mode], (key, r)
