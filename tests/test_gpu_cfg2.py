"""Parity at the benchmarked size: BASELINE configs[1] (128 clouds x 8192 points -> 65 536 groups) through the
public PointTokenizer against tests/golden/bench_cfg2.npz, which oracle/gen_golden.py recorded from the UNMODIFIED
reference (Group dvae.py:152-181, Encoder dvae.py:184-215, reduce_dim point_encoder.py:133,239) on the same batch
and weights.  bench.py runs the same check on its own timed configuration ("parity_checked")."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp16", "bf16", "fp32"])
def test_tokenizer_cfg2_against_reference_fixture(precision):
    from ppt_b200 import ops
    tok = bench.make_tokenizer(precision).cuda()
    xyz = bench.make_host_batches(0, 1, pin=False)[0].cuda()
    start = torch.zeros(xyz.shape[0], dtype=torch.int64, device="cuda")
    fps_idx, center = ops.fps(xyz, bench.N_GROUP, start, return_centers=True)
    nb, knn_idx = ops.knn_group(xyz, center, bench.GROUP_SIZE, return_idx=True)
    tokens, center2, nb2 = tok(xyz, return_neighborhood=True)   # the public call: same kernels, shared spatial index
    assert torch.equal(center2, center) and torch.equal(nb2, nb)
    r = bench.check_cfg2_parity(fps_idx.cpu(), center.cpu(), knn_idx.cpu(), nb.cpu(), tokens.cpu(), precision)
    assert r["ok"], r


def test_tokenizer_cfg2_halves_agree():
    """Size-independent property: the tokens of a cloud do not depend on which batch it rides in (no cross-cloud
    state in eval mode): the 128-cloud call equals two 64-cloud calls bit for bit."""
    tok = bench.make_tokenizer("fp16").cuda()
    xyz = bench.make_host_batches(0, 1, pin=False)[0].cuda()
    full, _ = tok(xyz)
    a, _ = tok(xyz[:64].contiguous())
    a = a.clone()
    b, _ = tok(xyz[64:].contiguous())
    assert torch.equal(full[:64], a) and torch.equal(full[64:], b)
