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


def test_fp16_token_output_is_the_fp32_output_rounded_once():
    """PPT_TOKENS_F16 (ops.encoder_forward token_dtype=float16): the last kernel's epilogue stores round-to-nearest
    fp16 of exactly the value it would have stored as fp32."""
    tok = bench.make_tokenizer("fp16").cuda()
    xyz = bench.make_host_batches(0, 1, batch=8, pin=False)[0].cuda()
    t32, c32 = tok(xyz)
    t16, c16 = tok(xyz, token_dtype=torch.float16)
    assert t16.dtype == torch.float16 and torch.equal(c16, c32)
    assert torch.equal(t16, t32.half())


def test_two_streams_do_not_share_scratch():
    """ops._workspace is keyed by (device, purpose, stream): two tokenizer calls in flight on two streams give
    the results of the same calls run one after the other (shared scratch raced)."""
    tok = bench.make_tokenizer("fp16").cuda()
    a, b = (t.cuda() for t in bench.make_host_batches(0, 2, batch=32, pin=False))
    want_a, want_b = tok(a)[0].clone(), tok(b)[0].clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            got_a = tok(a)[0]
        with torch.cuda.stream(s2):
            got_b = tok(b)[0]
        torch.cuda.synchronize()
        assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)


def test_host_pipeline_delivers_the_same_tokens():
    from ppt_b200.tokenizer import HostPipeline
    tok = bench.make_tokenizer("fp16").cuda()
    host = bench.make_host_batches(0, 4, batch=16)
    want = [tok(h.cuda())[0].cpu() for h in host]
    for dt in (torch.float32, torch.float16):
        pipe = HostPipeline(tok, 16, bench.N_POINTS, depth=3, device=torch.device("cuda", 0), token_dtype=dt)
        got = {}
        pipe.run(iter(host), on_result=lambda i, t, c: got.__setitem__(i, t.clone()))
        assert sorted(got) == [0, 1, 2, 3]
        for i in range(4):
            assert torch.equal(got[i], want[i].to(dt))
