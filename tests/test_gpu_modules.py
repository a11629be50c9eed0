"""The reference-facing layer (ppt_b200.pointbert / pointnet2 / patch / tokenizer) on the GPU:
same signatures and results as the reference's functions and modules."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import cpu, torch_port
from oracle.inputs import cloud

pytestmark = pytest.mark.gpu

# The SA / FP modules keep their own torch Conv/BatchNorm stacks; compare them against the CPU in full
# fp32 (torch's GPU default lets cuDNN/cuBLAS use TF32, which is not what these tests are about).
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def bits(t):
    return np.ascontiguousarray(t.detach().cpu().numpy()).view(np.uint32)


def test_group_module_matches_oracle_and_consumes_rng_like_the_reference():
    from ppt_b200 import pointbert
    xyz = cloud("U", 3, 2048, 77)
    grp = pointbert.Group(128, 32)
    grp.start_idx = 0
    nb, center = grp(xyz.cuda())
    o_nb, o_c, _, _ = cpu.group_forward(xyz.numpy(), 128, 32, 0)
    assert nb.shape == (3, 128, 32, 3) and center.shape == (3, 128, 3)
    assert np.array_equal(bits(nb), o_nb.view(np.uint32)) and np.array_equal(bits(center), o_c.view(np.uint32))
    # default start: the reference's own torch.randint call on the device (misc.py:59)
    grp.start_idx = None
    torch.manual_seed(11)
    nb2, c2 = grp(xyz.cuda())
    torch.manual_seed(11)
    start = torch.randint(0, 2048, (3,), dtype=torch.long, device="cuda")
    want = np.stack([cpu.group_forward(xyz[i:i + 1].numpy(), 128, 32, int(start[i]))[1][0] for i in range(3)])
    assert np.array_equal(bits(c2), want.view(np.uint32))
    # function forms
    idx = pointbert.farthest_point_sample(xyz.cuda(), 64, start_idx=0)
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), cpu.farthest_point_sample(xyz.numpy(), 64, 0))
    assert torch.equal(pointbert.fps(xyz.cuda(), 64, start_idx=0), pointbert.index_points(xyz.cuda(), idx))
    k = pointbert.knn_point(8, xyz.cuda(), xyz[:, :10].contiguous().cuda())
    assert np.array_equal(k.cpu().numpy(), cpu.knn_point(8, xyz.numpy(), xyz[:, :10].numpy()))
    d = pointbert.square_distance(xyz[:, :10].contiguous().cuda(), xyz.cuda())
    assert np.array_equal(bits(d), cpu.square_distance(xyz[:, :10].numpy(), xyz.numpy()).view(np.uint32))


def test_encoder_module_state_dict_names_and_eval_output():
    from ppt_b200 import pointbert
    enc = pointbert.Encoder(256)
    keys = set(enc.state_dict())
    assert set(torch_port.ENCODER_KEYS) <= keys
    assert keys - set(torch_port.ENCODER_KEYS) == {"first_conv.1.num_batches_tracked",
                                                   "second_conv.1.num_batches_tracked"}
    sd = torch_port.make_encoder_state()
    enc.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    enc = enc.cuda().eval()
    nb = (torch.rand(2, 70, 32, 3, generator=torch.Generator().manual_seed(5)) - 0.5) * 0.5
    with torch.no_grad():
        got = enc(nb.cuda())
        ref = torch_port.encoder_forward(sd, nb)
    assert got.shape == (2, 70, 256)
    assert float((got.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-3
    # weights change -> repack (version counter), not a stale blob
    with torch.no_grad():
        enc.second_conv[3].bias.add_(1.0)
        got2 = enc(nb.cuda())
    assert float((got2 - got - 1.0).abs().max()) < 1e-3
    # train(): batch-statistics BatchNorm through the module's own layers
    enc.train()
    before = enc.first_conv[1].running_mean.clone()
    out = enc(nb.cuda())
    assert out.shape == (2, 70, 256) and not torch.equal(before, enc.first_conv[1].running_mean)


def test_tokenizer_end_to_end_and_fp32_parity_mode():
    from ppt_b200.tokenizer import PointTokenizer
    sd = torch_port.make_encoder_state()
    xyz = cloud("S", 2, 4096, 21)
    o_nb, o_c, _, _ = cpu.group_forward(xyz.numpy(), 256, 32, 0)
    with torch.no_grad():
        ref = torch_port.tokens_forward(sd, torch.from_numpy(o_nb))
    for precision, tol in (("fp16", 1e-3), ("fp32", 2e-5)):
        tok = PointTokenizer(256, 32, precision=precision).cuda().eval().load_reference_state(sd)
        tok.start_idx = 0
        tokens, center, nb = tok(xyz.cuda(), return_neighborhood=True)
        assert tokens.shape == (2, 256, 384)
        assert np.array_equal(bits(nb), o_nb.view(np.uint32)) and np.array_equal(bits(center), o_c.view(np.uint32))
        err = float((tokens.cpu() - ref).abs().max() / ref.abs().max())
        assert err <= tol, (precision, err)


def _ssg_reference(sa, xyz, feats):
    """Same module arithmetic on CPU with the oracle's grouping."""
    f = cpu.farthest_point_sample(xyz.numpy(), sa.npoint, 0)
    c = cpu.index_points(xyz.numpy(), f)
    b = cpu.query_ball_point(sa.radius, sa.nsample, xyz.numpy(), c)
    g = cpu.group_center(xyz.numpy(), b, c)
    if feats is not None:
        g = np.concatenate([g, cpu.index_points(feats.numpy(), b)], -1)
    x = torch.from_numpy(g).permute(0, 3, 2, 1)
    for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
        x = torch.relu(bn(conv(x)))
    return torch.from_numpy(c).permute(0, 2, 1), x.max(2)[0]


def test_set_abstraction_modules():
    from ppt_b200 import pointnet2
    torch.manual_seed(0)
    xyz = cloud("S", 2, 1024, 3)
    feats = torch.randn(2, 1024, 16)
    sa = pointnet2.PointNetSetAbstraction(128, 0.3, 32, 16 + 3, [32, 64], False).eval()
    sa.start_idx = 0
    with torch.no_grad():
        want_xyz, want = _ssg_reference(sa, xyz, feats)
        sa.cuda()
        got_xyz, got = sa(xyz.permute(0, 2, 1).cuda(), feats.permute(0, 2, 1).cuda())
    assert got_xyz.shape == (2, 3, 128) and got.shape == (2, 64, 128)
    assert torch.equal(got_xyz.cpu(), want_xyz)
    assert float((got.cpu() - want).abs().max()) < 1e-4
    # MSG: [feats, xyz - centre] order, one FPS
    msg = pointnet2.PointNetSetAbstractionMsg(64, [0.2, 0.4], [8, 16], 16, [[16, 32], [16, 48]]).eval()
    msg.start_idx = 0
    with torch.no_grad():
        f = cpu.farthest_point_sample(xyz.numpy(), 64, 0)
        c = cpu.index_points(xyz.numpy(), f)
        outs = []
        for i, (r, k) in enumerate(((0.2, 8), (0.4, 16))):
            b = cpu.query_ball_point(r, k, xyz.numpy(), c)
            g = np.concatenate([cpu.index_points(feats.numpy(), b), cpu.group_center(xyz.numpy(), b, c)], -1)
            x = torch.from_numpy(g).permute(0, 3, 2, 1)
            for conv, bn in zip(msg.conv_blocks[i], msg.bn_blocks[i]):
                x = torch.relu(bn(conv(x)))
            outs.append(x.max(2)[0])
        want = torch.cat(outs, 1)
        msg.cuda()
        _, got = msg(xyz.permute(0, 2, 1).cuda(), feats.permute(0, 2, 1).cuda())
    assert got.shape == (2, 80, 64) and float((got.cpu() - want).abs().max()) < 1e-4
    # function form and returnfps
    nx, npts, gx, fi = pointnet2.sample_and_group(128, 0.3, 32, xyz.cuda(), feats.cuda(), returnfps=True, start_idx=0)
    assert npts.shape == (2, 128, 32, 19) and gx.shape == (2, 128, 32, 3) and fi.shape == (2, 128)
    assert torch.equal(npts[..., :3], gx - nx.unsqueeze(2))


def test_feature_propagation_module_forward_and_backward():
    from ppt_b200 import pointnet2
    torch.manual_seed(1)
    xyz1, xyz2 = cloud("S", 2, 512, 8), cloud("S", 2, 64, 9)
    p1, p2 = torch.randn(2, 5, 512), torch.randn(2, 24, 64)
    fp = pointnet2.PointNetFeaturePropagation(5 + 24, [32]).eval()
    with torch.no_grad():
        interp = torch_port.three_nn_interpolate(xyz1, xyz2, p2.permute(0, 2, 1))
        x = torch.cat([p1.permute(0, 2, 1), interp], -1).permute(0, 2, 1)
        want = torch.relu(fp.mlp_bns[0](fp.mlp_convs[0](x)))
    fp.cuda()
    p2g = p2.cuda().requires_grad_(True)
    got = fp(xyz1.permute(0, 2, 1).cuda(), xyz2.permute(0, 2, 1).cuda(), p1.cuda(), p2g)
    assert got.shape == (2, 32, 512) and float((got.detach().cpu() - want).abs().max()) < 1e-4
    got.sum().backward()
    assert p2g.grad is not None and p2g.grad.shape == p2.shape and float(p2g.grad.abs().sum()) > 0
    # S == 1 branch: plain repeat (pointnet2_utils.py:297-298)
    one = pointnet2.three_nn_interpolate(xyz1.cuda(), xyz2[:, :1].contiguous().cuda(), p2[:, :, :1].permute(0, 2, 1).cuda())
    assert one.shape == (2, 512, 24) and torch.equal(one[:, 0], one[:, 100])


def test_patch_reference_diverts_cuda_tensors_on_a_stand_in_tree():
    """The reference tree does not exist on the GPU box: a stand-in with the reference's module
    and attribute names (bodies from the torch-op port) shows the rebinding works on CUDA tensors."""
    from ppt_b200 import patch

    class Group(torch.nn.Module):
        def __init__(self, num_group, group_size):
            super().__init__()
            self.num_group, self.group_size = num_group, group_size

        def forward(self, xyz):
            return torch_port.group_forward(xyz, self.num_group, self.group_size, 0)

    class Encoder(torch.nn.Module):
        def __init__(self, c):
            super().__init__()
            self.encoder_channel = c

        def forward(self, pg):
            raise AssertionError("reference body must not run for CUDA eval input")

    mods = {}
    for name in ("models", "models.pointbert", "models.pointbert.dvae"):
        mods[name] = types.ModuleType(name)
    dvae = mods["models.pointbert.dvae"]
    dvae.knn_point = lambda k, xyz, q: torch_port.knn_indices(k, xyz, q)
    dvae.square_distance = torch_port.pairwise_sqdist
    dvae.Group, dvae.Encoder = Group, Encoder
    saved = {n: sys.modules.get(n) for n in mods}
    sys.modules.update(mods)
    try:
        names = patch.patch_reference(["models.pointbert.dvae"])
        assert "models.pointbert.dvae.knn_point" in names
        xyz = cloud("U", 2, 1024, 1)
        q = xyz[:, :20].contiguous()
        got = dvae.knn_point(16, xyz.cuda(), q.cuda())  # CUDA -> kernels
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), cpu.knn_point(16, xyz.numpy(), q.numpy()))
        cpu_got = dvae.knn_point(16, xyz, q)             # CPU -> the tree's own code
        assert not cpu_got.is_cuda
        torch.manual_seed(0)
        nb, c = dvae.Group(32, 8)(xyz.cuda())
        assert nb.is_cuda and nb.shape == (2, 32, 8, 3)
    finally:
        patch.unpatch_reference()
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_point_transformer_forward_with_fused_front_end_on_a_stand_in():
    """PointTransformer.forward (models/pointbert/point_encoder.py:234-256) with the fused front end, on a stand-in
    that has the reference's attribute names; the unfused body is lines 236-256 written with torch ops."""
    from ppt_b200 import patch, pointbert

    class Blocks(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(384, 384)

        def forward(self, x, pos, task="cls"):
            return self.lin(x + pos)

    class PointTransformer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.group_size, self.num_group = 32, 64
            self.group_divider = pointbert.Group(num_group=64, group_size=32)
            self.group_divider.start_idx = 0
            self.encoder = pointbert.Encoder(256)
            self.reduce_dim = torch.nn.Linear(256, 384)
            self.cls_token = torch.nn.Parameter(torch.randn(1, 1, 384) * 0.1)
            self.cls_pos = torch.nn.Parameter(torch.randn(1, 1, 384))
            self.pos_embed = torch.nn.Sequential(torch.nn.Linear(3, 128), torch.nn.GELU(), torch.nn.Linear(128, 384))
            self.blocks, self.norm = Blocks(), torch.nn.LayerNorm(384)

    def reference_body(self, pts):
        neighborhood, center = self.group_divider(pts)
        tokens = self.reduce_dim(self.encoder._forward_torch(neighborhood))
        x = torch.cat((self.cls_token.expand(tokens.size(0), -1, -1), tokens), dim=1)
        pos = torch.cat((self.cls_pos.expand(tokens.size(0), -1, -1), self.pos_embed(center)), dim=1)
        x = self.norm(self.blocks(x, pos, task="cls"))
        return torch.cat([x[:, 0], x[:, 1:].max(1)[0]], dim=-1)

    torch.manual_seed(3)
    model = PointTransformer()
    sd = torch_port.make_encoder_state()
    model.encoder.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    model = model.cuda().eval()
    xyz = cloud("U", 3, 2048, 9).cuda()
    calls = []
    with torch.no_grad():
        want = reference_body(model, xyz)
        got = patch._point_transformer_forward(model, xyz, lambda s, p: calls.append(1) or reference_body(s, p))
        assert not calls, "the fused path must run for CUDA eval input"
        assert float((got - want).abs().max() / want.abs().max()) < 2e-3
        x, pos = patch.point_transformer_front_end(model, xyz)
        assert x.shape == (3, 65, 384) and torch.equal(x[:, 0], model.cls_token.expand(3, -1, -1)[:, 0])
        assert torch.equal(pos[:, 0], model.cls_pos.expand(3, -1, -1)[:, 0])
        with torch.no_grad():
            model.cls_pos.add_(1.0)   # a parameter update must invalidate the packed blob
        _, pos2 = patch.point_transformer_front_end(model, xyz)
        assert torch.allclose(pos2[:, 0], pos[:, 0] + 1.0)
    model.encoder.train()
    with torch.no_grad():
        patch._point_transformer_forward(model, xyz, lambda s, p: calls.append(1) or want)
    assert calls == [1], "train-mode BatchNorm keeps the reference body"


def test_graph_feature_matches_reference_fixture_and_gradients():
    """DGCNN_Propagation.get_graph_feature (pointnet2_utils.py:392-442, row f4): forward bit-exact against the
    fixture recorded from the unmodified reference; backward against autograd through the torch restatement."""
    from oracle.inputs import digest
    from ppt_b200 import ops, pointnet2
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_feature.npz"))
    for tag in ("small", "partseg"):
        coor_q, x_q, coor_k, x_k = (torch.from_numpy(f[tag + "." + n]).cuda() for n in ("coor_q", "x_q", "coor_k", "x_k"))
        feat = pointnet2.get_graph_feature(coor_q, x_q, coor_k, x_k, 4)
        assert feat.shape == (x_q.shape[0], 2 * x_q.shape[1], x_q.shape[2], 4) and feat.is_contiguous()
        if tag == "small":
            assert np.array_equal(feat.sort(-1)[0].cpu().numpy(), f["small.feature_sorted"])
        else:
            assert digest(feat.sort(-1)[0]) == str(f["partseg.feature_sorted_sha"])
        self_feat = pointnet2.get_graph_feature(coor_q, x_q, coor_q, x_q, 4)
        assert digest(self_feat.sort(-1)[0]) == str(f[tag + ".self_feature_sorted_sha"])
    # same neighbour order as the (distance, index)-ordered kNN kernel: elementwise equal to the restatement on its idx
    coor_q, x_q, coor_k, x_k = (torch.from_numpy(f["small." + n]) for n in ("coor_q", "x_q", "coor_k", "x_k"))
    idx = ops.knn(4, coor_k.permute(0, 2, 1).contiguous().cuda(), coor_q.permute(0, 2, 1).contiguous().cuda())
    xq, xk = x_q.clone().cuda().requires_grad_(True), x_k.clone().cuda().requires_grad_(True)
    out = ops.graph_feature(xq, xk, idx)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    rq, rk = x_q.clone().requires_grad_(True), x_k.clone().requires_grad_(True)
    B, C, Nq = rq.shape
    ic = idx.cpu()
    nb = torch.gather(rk.unsqueeze(2).expand(B, C, Nq, rk.shape[2]), 3, ic.unsqueeze(1).expand(B, C, Nq, 4))
    ref = torch.cat((nb - rq.unsqueeze(-1), rq.unsqueeze(-1).expand(-1, -1, -1, 4)), dim=1)
    assert torch.equal(out.detach().cpu(), ref.detach())
    (ref * w.cpu()).sum().backward()
    assert torch.allclose(xq.grad.cpu(), rq.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(xk.grad.cpu(), rk.grad, rtol=1e-5, atol=1e-5)


def test_loader_fps_drop_in_matches_reference_fixture():
    """ppt_b200.data.farthest_point_sample == data/dataset_3d.py:40-61 (row f4), including the numpy RNG draw."""
    from ppt_b200 import data
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_fps.npz"))
    for tag in ("a", "b"):
        point, npoint, seed = f[tag + ".point"], int(f[tag + ".npoint"]), int(f[tag + ".seed"])
        np.random.seed(seed)
        got = data.farthest_point_sample(point, npoint)             # draws np.random.randint(0, N) like the reference
        assert got.dtype == np.float32 and np.array_equal(got, point[f[tag + ".indices"]])
        idx = data.farthest_point_sample_indices(point, npoint, start=int(f[tag + ".start"]))
        assert np.array_equal(idx, f[tag + ".indices"])
    both = np.stack([f["a.point"][:3000, :3], f["b.point"][:3000, :3]])
    bidx = data.farthest_point_sample_batch(both, 64, [int(f["a.start"]), 5])
    assert np.array_equal(bidx[0], f["a.indices"][:64])
    with pytest.raises(TypeError):
        data.farthest_point_sample(f["a.point"].astype(np.float64), 8)


def test_patch_reference_on_the_real_reference_tree_with_cuda_tensors():
    """`refonly`: runs only where both a GPU and the reference tree (PPT_REFERENCE_ROOT, default /root/reference) exist.
    patch_reference() on the UNMODIFIED modules: Group / Encoder / knn_point / PointTransformer.forward on CUDA tensors go
    through the kernels and agree with the same modules on CPU tensors (which keep the reference's own code)."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip("reference tree not present on this box (PPT_REFERENCE_ROOT)")
    from ppt_b200 import patch
    ns = refimport.load()
    torch.manual_seed(0)
    xyz = cloud("U", 2, 2048, 77)
    grp, enc = ns.dvae.Group(64, 32), ns.dvae.Encoder(256).eval()
    enc.load_state_dict({k: v for k, v in torch_port.make_encoder_state().items() if k in torch_port.ENCODER_KEYS},
                        strict=False)
    with refimport.fixed_fps_start(0), torch.no_grad():
        nb_ref, c_ref = grp(xyz)
        f_ref = enc(nb_ref)
        k_ref = ns.dvae.knn_point(8, xyz, c_ref)
    names = patch.patch_reference()
    try:
        assert "Group.forward" in " ".join(names) and "Encoder.forward" in " ".join(names)
        with refimport.fixed_fps_start(0), torch.no_grad():
            nb, c = grp.cuda()(xyz.cuda())
            f = enc.cuda()(nb)
            k = ns.dvae.knn_point(8, xyz.cuda(), c)
        assert torch.equal(c.cpu(), c_ref)
        assert torch.equal(nb.cpu().sort(2)[0], nb_ref.sort(2)[0])       # neighbour order inside a group is unspecified
        assert torch.equal(k.cpu().sort(-1)[0], k_ref.sort(-1)[0])
        assert float((f.cpu() - f_ref).abs().max() / f_ref.abs().max()) < 1e-3
    finally:
        patch.unpatch_reference()


def test_dgcnn_propagation_fused_matches_the_reference_fixture():
    """Row f4, dense half: DGCNN_Propagation.forward (models/pointbert/pointnet2_utils.py:444-467) with the edge
    convolution split into per-point GEMMs and the fused GroupNorm / LeakyReLU / max kernel, against the output of the
    unmodified reference module (tests/golden/dgcnn.npz); the unfused body (gradient needed) on the same kernels too."""
    from oracle.inputs import digest
    from ppt_b200 import ops, pointnet2
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dgcnn.npz"))
    mod = pointnet2.DGCNN_Propagation(k=4)
    mod.load_state_dict(torch_port.make_dgcnn_state(31))
    mod = mod.cuda().eval()
    calls = []
    real = ops.edge_gn_max
    ops.edge_gn_max = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        for tag, (B, Nk, Nq) in (("cross", (2, 64, 128)), ("up", (1, 96, 50))):
            inputs = torch_port.dgcnn_inputs(4700 + Nk, B, Nk, Nq)
            assert digest(torch.cat([t.reshape(-1) for t in inputs]).numpy()) == str(f[tag + ".inputs_sha"])
            want = torch.from_numpy(f[tag + ".out"])
            for precision, tol in (("fp32", 1e-4), ("fp16", 3e-3)):  # GEMM operands; norm-relative like the Encoder's
                mod.ppt_precision = precision
                n0 = len(calls)
                with torch.no_grad():
                    got = mod(*(t.cuda() for t in inputs))
                assert len(calls) == n0 + 2, "both layers must take the fused path"
                assert got.shape == want.shape
                assert float((got.cpu() - want).abs().max() / want.abs().max()) < tol, precision
            n0 = len(calls)
            got2 = mod(*(t.cuda() for t in inputs))     # parameters require grad: the module's own layers
            assert len(calls) == n0 and got2.requires_grad
            assert float((got2.detach().cpu() - want).abs().max() / want.abs().max()) < 1e-4
    finally:
        ops.edge_gn_max = real


def test_feature_propagation_mlp_folded_in_eval_mode():
    """PointNetFeaturePropagation's Conv1d + BatchNorm1d + ReLU stack with BatchNorm folded (eval, no gradient) equals
    the module's own layers; train mode / gradients keep the layers."""
    from ppt_b200 import pointnet2
    torch.manual_seed(4)
    fp = pointnet2.PointNetFeaturePropagation(19 + 24, [64, 32]).cuda().eval()
    with torch.no_grad():
        for bn in fp.mlp_bns:
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, 43, 300, device="cuda")
    with torch.no_grad():
        want = x
        for conv, bn in zip(fp.mlp_convs, fp.mlp_bns):
            want = torch.relu(bn(conv(want)))
        got = pointnet2.feature_propagation_mlp(fp, x.clone())
    assert "_ppt_fp_folded" in fp.__dict__ and float((got - want).abs().max()) < 1e-4
    out = pointnet2.feature_propagation_mlp(fp, x.clone())      # grad enabled, trainable parameters
    assert out.requires_grad


@pytest.mark.parametrize("shape", [(2, 1000, 77, 19, 384, 1536, 384), (1, 300, 64, 0, 96, 256, 128), (2, 256, 512, 3, 384, 1536, 384)])
def test_feature_propagation_fused_on_tensor_cores(shape):
    """Row f4, dense half: PointNetFeaturePropagation.forward in eval mode -- three_nn + (interpolation + concat fused
    into the operand build) + the two-layer Conv1d/BatchNorm1d/ReLU MLP on tcgen05 (ops.fp_mlp_forward) -- against the
    same module's fp32 torch layers on the torch-op interpolation (the reference's own arithmetic, bit-identical on CPU,
    tests/test_host_cpu.py).  Shapes: the part-seg head's propagation_0 (19 + 384 -> 1536 -> 384, ragged point count),
    no skip features, and propagation_1 (3 + 384)."""
    from ppt_b200 import ops, pointnet2
    B, N, S, D1, D2, C1, C2 = shape
    torch.manual_seed(N)
    fp = pointnet2.PointNetFeaturePropagation(D1 + D2, [C1, C2]).eval()
    with torch.no_grad():
        for bn in fp.mlp_bns:
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.1)
    for p in fp.parameters():
        p.requires_grad_(False)   # the fused path is forward only
    xyz1, xyz2 = cloud("S", B, N, 21), cloud("S", B, S, 22)
    p1 = torch.randn(B, D1, N) if D1 else None
    p2 = torch.randn(B, D2, S)
    with torch.no_grad():
        interp = torch_port.three_nn_interpolate(xyz1, xyz2, p2.permute(0, 2, 1))
        x = interp.permute(0, 2, 1) if p1 is None else torch.cat([p1, interp.permute(0, 2, 1)], dim=1)
        want = x
        for conv, bn in zip(fp.mlp_convs, fp.mlp_bns):
            want = torch.relu(bn(conv(want)))
    fp = fp.cuda()
    calls = []
    real = ops.fp_mlp_forward
    ops.fp_mlp_forward = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        got = fp(xyz1.permute(0, 2, 1).cuda(), xyz2.permute(0, 2, 1).cuda(), None if p1 is None else p1.cuda(), p2.cuda())
        assert calls == [1], "the tensor-core path must have run"
        assert got.shape == want.shape and got.is_contiguous() and bool(torch.isfinite(got).all())
        d = got.double().cpu() - want.double()
        # two chained layers with fp16 operands, norm-relative like the Encoder's tokens
        assert float(d.abs().max() / want.abs().max()) < 2e-3 and float(d.norm() / want.double().norm()) < 1e-3
        fp.train()
        n0 = len(calls)
        fp(xyz1.permute(0, 2, 1).cuda(), xyz2.permute(0, 2, 1).cuda(), None if p1 is None else p1.cuda(), p2.cuda())
        assert len(calls) == n0, "train mode keeps the module's own layers (batch statistics)"
    finally:
        ops.fp_mlp_forward = real


def test_set_abstraction_msg_level3_wide_input_on_tensor_cores():
    """MSG level 3 (models/pointnet2/pointnet2.py:47: group_all over 128 points with 640 + 3 input channels, MLP
    [256, 512, 1024]) was the one level left on the torch layers in round 1 (more input channels than the fused kernel
    keeps in shared memory): its first layer now runs K-blocked.  Against the module's own fp32 layers."""
    from ppt_b200 import ops, pointnet2
    torch.manual_seed(5)
    sa = pointnet2.PointNetSetAbstraction(None, None, None, 640 + 3, [256, 512, 1024], True).eval()
    with torch.no_grad():
        for bn in sa.mlp_bns:
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    for p in sa.parameters():
        p.requires_grad_(False)
    sa = sa.cuda()
    xyz = cloud("S", 3, 128, 31).permute(0, 2, 1).contiguous().cuda()
    feats = torch.randn(3, 640, 128, device="cuda")
    calls = []
    real = ops.sa_mlp_forward
    ops.sa_mlp_forward = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        _, got = sa(xyz, feats)
        assert calls == [1], "the tensor-core path must have run"
        supported = ops.sa_mlp_supported
        ops.sa_mlp_supported = lambda *a: False
        try:
            _, want = sa(xyz, feats)
        finally:
            ops.sa_mlp_supported = supported
    finally:
        ops.sa_mlp_forward = real
    assert got.shape == want.shape == (3, 1024, 1)
    d = (got - want).double()
    assert float(d.abs().max() / want.abs().max()) < 2e-3 and float(d.norm() / want.double().norm()) < 2e-3
