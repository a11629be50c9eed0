"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

The reference ships no golden vectors or unit tests for this path
(SURVEY.md section 4), so these fixtures are the parity pins: outputs of the
reference's own functions on seeded inputs, with FPS start index 0.  They are
valid for the torch build recorded in each file (`torch_version`).  Small cases
store full arrays; BASELINE-size cases store sha256 digests of canonicalised
outputs (a checksum of checksums) so the fixtures stay small.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402
from oracle import torch_port  # noqa: E402
from oracle.inputs import cloud, digest  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")




def small_int(a, hi):
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    return a.astype(np.int16 if hi < 32768 else np.int32)


def knn_tie_rows(sqd, k):
    """Rows whose k-th and (k+1)-th smallest distances are equal."""
    v = torch.topk(sqd, k + 1, dim=-1, largest=False, sorted=True).values
    return (v[..., k - 1] == v[..., k]).numpy()


def canon_group(nb, idx):
    """Reorders each group's rows by point index so unordered top-k output compares."""
    order = idx.argsort(dim=-1)
    return torch.gather(nb, 2, order.unsqueeze(-1).expand_as(nb))


def gen_group(ns, name, kind, B, N, G, K, seed, full):
    xyz = cloud(kind, B, N, seed)
    with refimport.fixed_fps_start(0):
        fps_idx = ns.misc.farthest_point_sample(xyz, G)
        nb, center = ns.dvae.Group(G, K)(xyz)
    sqd = ns.dvae.square_distance(center, xyz)
    knn = ns.dvae.knn_point(K, xyz, center)
    knn_sorted = knn.sort(-1).values
    tie = knn_tie_rows(sqd, K)
    nb_c = canon_group(nb, knn)
    rec = dict(kind=kind, B=B, N=N, G=G, K=K, seed=seed, torch_version=torch.__version__,
               xyz_sha=digest(xyz.numpy()), fps_sha=digest(fps_idx.numpy()),
               knn_sorted_sha=digest(knn_sorted.numpy()), center_sha=digest(center.numpy()),
               nb_canon_sha=digest(nb_c.numpy()), tie_rows=np.argwhere(tie).astype(np.int32),
               n_negative_self=int((sqd.min(-1).values < 0).sum()))
    if full:
        rec.update(xyz=xyz.numpy(), fps_idx=small_int(fps_idx, N), knn_sorted=small_int(knn_sorted, N),
                   center=center.numpy(), nb_canon=nb_c.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "tie rows", int(tie.sum()), "neg self", rec["n_negative_self"])


def gen_sqdist(ns):
    """Bit pattern of square_distance on a small ragged case (F2, F3)."""
    src = cloud("U", 2, 37, 77)
    dst = cloud("U", 2, 1000, 78)
    d = ns.dvae.square_distance(src, dst)
    np.savez_compressed(os.path.join(OUT, "sqdist_small.npz"), src=src.numpy(), dst=dst.numpy(),
                        dist_bits=d.numpy().view(np.uint32), torch_version=torch.__version__)


def gen_sa(ns, name, B, N, seed, full):
    """PointNet++ SSG grouping (models/pointnet2/pointnet2.py:11-12): cfg 3."""
    xyz = cloud("S", B, N, seed)
    g = torch.Generator().manual_seed(seed + 100)
    with refimport.fixed_fps_start(0):
        f1 = ns.pn2.farthest_point_sample(xyz, 512)
        c1 = ns.pn2.index_points(xyz, f1)
        b1 = ns.pn2.query_ball_point(0.2, 32, xyz, c1)
        new_xyz1, new_pts1 = ns.pn2.sample_and_group(512, 0.2, 32, xyz, None)
        feats = torch.randn(B, 512, 128, generator=g)
        f2 = ns.pn2.farthest_point_sample(c1, 128)
        c2 = ns.pn2.index_points(c1, f2)
        b2 = ns.pn2.query_ball_point(0.4, 64, c1, c2)
        new_xyz2, new_pts2 = ns.pn2.sample_and_group(128, 0.4, 64, c1, feats)
    rec = dict(B=B, N=N, seed=seed, torch_version=torch.__version__, xyz_sha=digest(xyz.numpy()),
               fps1_sha=digest(f1.numpy()), ball1_sha=digest(b1.numpy()), grp1_sha=digest(new_pts1.numpy()),
               fps2_sha=digest(f2.numpy()), ball2_sha=digest(b2.numpy()), grp2_sha=digest(new_pts2.numpy()),
               feats_sha=digest(feats.numpy()))
    if full:
        rec.update(xyz=xyz.numpy(), fps1=small_int(f1, N), ball1=small_int(b1, N),
                   fps2=small_int(f2, N), ball2=small_int(b2, N))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    occ = (b1 != b1[..., :1]).sum(-1).float().mean().item() + 1
    print(name, "mean occupancy sa1 ~", round(occ, 1))


def gen_msg_fp(ns, name, B, N, seed, D, full):
    """MSG ball queries (models/pointnet2/pointnet2.py:45-46) and part-seg
    feature propagation (three_nn + interpolate): cfg 4."""
    xyz = cloud("S", B, N, seed)
    g = torch.Generator().manual_seed(seed + 100)
    with refimport.fixed_fps_start(0):
        f1 = ns.pn2.farthest_point_sample(xyz, 512)
    c1 = ns.pn2.index_points(xyz, f1)
    rec = dict(B=B, N=N, D=D, seed=seed, torch_version=torch.__version__, xyz_sha=digest(xyz.numpy()),
               fps1_sha=digest(f1.numpy()))
    for r, k in ((0.1, 16), (0.2, 32), (0.4, 128)):
        b = ns.pn2.query_ball_point(r, k, xyz, c1)
        rec["ball_%g_%d_sha" % (r, k)] = digest(b.numpy())
        if full:
            rec["ball_%g_%d" % (r, k)] = small_int(b, N + 1)
    # three_nn / three_interpolate exactly as PointNetFeaturePropagation.forward does
    # it (models/pointnet2/pointnet2_utils.py:300-307), channel-last.
    feats = torch.randn(B, 512, D, generator=g)
    d, i = ns.pn2.square_distance(xyz, c1).sort(dim=-1)
    d, i = d[:, :, :3], i[:, :, :3]
    r = 1.0 / (d + 1e-8)
    w = r / torch.sum(r, dim=2, keepdim=True)
    interp = torch.sum(ns.pn2.index_points(feats, i) * w.view(B, N, 3, 1), dim=2)
    # rows where the 3rd/4th nearest tie are not uniquely defined
    d4 = torch.topk(ns.pn2.square_distance(xyz, c1), 4, dim=-1, largest=False, sorted=True).values
    tie = ((d4[..., 2] == d4[..., 3]) | (d4[..., 0] == d4[..., 1]) | (d4[..., 1] == d4[..., 2])).numpy()
    rec.update(nn_dist_sha=digest(d.numpy()), nn_idx_sha=digest(i.numpy()), interp_sha=digest(interp.numpy()),
               feats_sha=digest(feats.numpy()), nn_tie_rows=np.argwhere(tie).astype(np.int32),
               n_negative=int((d < 0).sum()))
    if full:
        rec.update(xyz=xyz.numpy(), fps1=small_int(f1, N), nn_dist_bits=d.numpy().view(np.uint32),
                   nn_idx=small_int(i, 512), feats=feats.numpy(), interp_bits=interp.numpy().view(np.uint32))
    # the whole module's interpolation path through the reference class (no MLP: mlp=[])
    fp = ns.pn2.PointNetFeaturePropagation(D, [])
    out = fp(xyz.permute(0, 2, 1), c1.permute(0, 2, 1), None, feats.permute(0, 2, 1))
    assert torch.equal(out.permute(0, 2, 1), interp)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "nn tie rows", int(tie.sum()), "negative d", rec["n_negative"])


def gen_encoder(ns):
    """Reference Encoder (models/pointbert/dvae.py:184-215) + reduce_dim
    (models/pointbert/point_encoder.py:133,239), eval mode, seeded weights."""
    sd = torch_port.make_encoder_state()
    enc = ns.dvae.Encoder(256).eval()
    missing = enc.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    reduce_dim = torch.nn.Linear(256, 384)
    reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
    xyz = cloud("U", 2, 1024, 4242)
    with refimport.fixed_fps_start(0):
        nb, _ = ns.dvae.Group(64, 32)(xyz)
    with torch.no_grad():
        feat = enc(nb)
        tok = reduce_dim(feat)
    wsum = hashlib.sha256(b"".join(np.ascontiguousarray(sd[k].numpy()).tobytes() for k in sorted(sd))).hexdigest()
    np.savez_compressed(os.path.join(OUT, "encoder_small.npz"), neighborhood=nb.numpy(), features=feat.numpy(),
                        tokens=tok.numpy(), weights_sha=wsum, torch_version=torch.__version__)
    print("encoder tokens", tuple(tok.shape), "absmax", float(tok.abs().max()))


def gen_encoder_train(ns):
    """Reference Encoder under .train() (dvae.py:184-215; main_cls.py:169 puts the frozen module in this mode):
    features, reduce_dim tokens and the running statistics after ONE forward, seeded weights."""
    sd = torch_port.make_encoder_state()
    enc = ns.dvae.Encoder(256).train()
    missing = enc.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    reduce_dim = torch.nn.Linear(256, 384)
    reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
    xyz = cloud("U", 3, 1024, 4444)
    with refimport.fixed_fps_start(0):
        nb, _ = ns.dvae.Group(50, 32)(xyz)   # 150 groups: a ragged last tile
    with torch.no_grad():
        feat = enc(nb)
        tok = reduce_dim(feat)
    after = {k: v.clone() for k, v in enc.state_dict().items() if "running_" in k or "num_batches" in k}
    port_feat, port_stats = torch_port.encoder_forward_train(sd, nb)
    assert (port_feat - feat).abs().max() <= 1e-6 * feat.abs().max()
    for k, v in port_stats.items():
        assert (v - after[k]).abs().max() <= 1e-6 * after[k].abs().max(), k
    rec = {"neighborhood": nb.numpy(), "features": feat.numpy(), "tokens": tok.numpy(),
           "torch_version": torch.__version__}
    for k, v in after.items():
        rec["after." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "encoder_train_small.npz"), **rec)
    print("encoder train-mode features", tuple(feat.shape), "absmax", float(feat.abs().max()))


def gen_graph_feature(ns):
    """DGCNN_Propagation.get_graph_feature of the unmodified reference (pointnet2_utils.py:392-442), k = 4:
    the cross-level call (queries 128, keys 64) in full and the part-seg point counts (512 <- 256; C = 96 to keep the fixture small) as digests."""
    from oracle.inputs import digest
    dg = ns.pb_pn2.DGCNN_Propagation(k=4)
    g = torch.Generator().manual_seed(4545)
    rec = {}
    for tag, (B, C, Nq, Nk), full in (("small", (2, 24, 128, 64), True), ("partseg", (2, 96, 512, 256), False)):
        coor_k = cloud("S", B, Nk, 4545 + Nk).permute(0, 2, 1).contiguous()
        coor_q = cloud("S", B, Nq, 4546 + Nq).permute(0, 2, 1).contiguous()
        x_q, x_k = torch.randn(B, C, Nq, generator=g), torch.randn(B, C, Nk, generator=g)
        with torch.no_grad():
            feat = dg.get_graph_feature(coor_q, x_q, coor_k, x_k)
        port, idx = torch_port.graph_feature(coor_q, x_q, coor_k, x_k, 4)
        # neighbour order inside the k is unspecified in the reference (topk sorted=False): compare per-(q) sets
        assert torch.equal(feat.sort(-1)[0], port.sort(-1)[0])
        rec.update({tag + ".coor_q": coor_q.numpy(), tag + ".coor_k": coor_k.numpy(), tag + ".x_q": x_q.numpy(),
                    tag + ".x_k": x_k.numpy()})
        if full:
            rec[tag + ".feature_sorted"] = feat.sort(-1)[0].numpy()
            rec[tag + ".idx_sorted"] = idx.sort(-1)[0].numpy()
        else:
            rec[tag + ".feature_sorted_sha"] = digest(feat.sort(-1)[0])
        # self-graph call of the second layer (keys = queries)
        with torch.no_grad():
            feat2 = dg.get_graph_feature(coor_q, x_q, coor_q, x_q)
        rec[tag + ".self_feature_sorted_sha"] = digest(feat2.sort(-1)[0])
    np.savez_compressed(os.path.join(OUT, "graph_feature.npz"), **rec)
    print("graph_feature fixtures", sorted(rec))


def gen_loader_fps(ns):
    """The data loader's numpy FPS (data/dataset_3d.py:40-61).  The module itself cannot be imported (F12), so the
    function's own source text is executed; np.random.seed pins its np.random.randint draw."""
    import re
    src = open(os.path.join(refimport.REFERENCE_ROOT, "data", "dataset_3d.py")).read()
    m = re.search(r"^def farthest_point_sample\(point, npoint\):.*?^    return point\n", src, re.S | re.M)
    scope = {"np": np}
    exec(m.group(0), scope)  # noqa: S102 -- the reference's function, unmodified
    from oracle import cpu
    rec = {}
    for tag, (N, D, npoint, seed) in (("a", (3000, 6, 256, 7)), ("b", (10000, 3, 1024, 8))):
        rng = np.random.RandomState(seed)
        point = rng.uniform(-1, 1, size=(N, D)).astype(np.float32)
        np.random.seed(seed)
        start = np.random.randint(0, N)
        np.random.seed(seed)
        ref = scope["farthest_point_sample"](point, npoint)
        idx = cpu.loader_fps_indices(point, npoint, start)
        assert np.array_equal(point[idx], ref)
        rec.update({tag + ".point": point, tag + ".npoint": npoint, tag + ".seed": seed, tag + ".start": start,
                    tag + ".indices": idx})
    np.savez_compressed(os.path.join(OUT, "loader_fps.npz"), **rec)
    print("loader fps fixtures", {k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items()})


def gen_sa_mlp(ns):
    """The unmodified reference PointNetSetAbstraction / ...Msg in eval mode with seeded weights (three-layer
    shared MLPs as in models/pointnet2/pointnet2.py:11-13, 45-47): outputs for the fused tensor-core path (row f1)."""
    rec = {}
    # SSG level 2: 131 -> 128 -> 128 -> 256, ball (0.4, 64); 512 points with 128-d features, 128 centres
    xyz = cloud("S", 1, 512, 4646)
    feats = torch.randn(1, 512, 128, generator=torch.Generator().manual_seed(4646))
    sa = ns.pn2.PointNetSetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False).eval()
    sa.load_state_dict(torch_port.make_sa_state(131, [128, 128, 256], 11), strict=False)
    with refimport.fixed_fps_start(0), torch.no_grad():
        nx, out = sa(xyz.permute(0, 2, 1), feats.permute(0, 2, 1))
    rec.update({"ssg2.xyz": xyz.numpy(), "ssg2.feats": feats.numpy(), "ssg2.new_xyz": nx.numpy(), "ssg2.out": out.numpy()})
    # SSG level 3 (group_all): 259 -> 256 -> 512 -> 1024 over 128 points
    xyz3 = cloud("S", 2, 128, 4647)
    feats3 = torch.randn(2, 128, 256, generator=torch.Generator().manual_seed(4647))
    sa3 = ns.pn2.PointNetSetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True).eval()
    sa3.load_state_dict(torch_port.make_sa_state(259, [256, 512, 1024], 12), strict=False)
    with torch.no_grad():
        _, out3 = sa3(xyz3.permute(0, 2, 1), feats3.permute(0, 2, 1))
    rec.update({"ssg3.xyz": xyz3.numpy(), "ssg3.feats": feats3.numpy(), "ssg3.out": out3.numpy()})
    # MSG level 1: xyz only, three radii / nsample 16, 32, 128, widths of pointnet2.py:45
    xyzm = cloud("S", 2, 1024, 4648)
    widths = [[32, 32, 64], [64, 64, 128], [64, 96, 128]]
    msg = ns.pn2.PointNetSetAbstractionMsg(128, [0.1, 0.2, 0.4], [16, 32, 128], 0, widths).eval()
    sd = {}
    for j, w in enumerate(widths):
        sd.update(torch_port.make_sa_state(3, w, 20 + j, "conv_blocks.%d." % j, "bn_blocks.%d." % j))
    msg.load_state_dict(sd, strict=False)
    with refimport.fixed_fps_start(0), torch.no_grad():
        _, outm = msg(xyzm.permute(0, 2, 1), None)
    rec.update({"msg1.xyz": xyzm.numpy(), "msg1.out": outm.numpy()})
    np.savez_compressed(os.path.join(OUT, "sa_mlp.npz"), **rec)
    print("sa_mlp fixtures", {k: v.shape for k, v in rec.items()})


def gen_front_end(ns):
    """The unmodified reference PointTransformer (models/pointbert/point_encoder.py:111-256) up to the call of
    self.blocks: a forward pre-hook records the (x, pos) it is given (:241-249).  depth 1 keeps the unused
    transformer small; group_divider / encoder / reduce_dim / cls rows / pos_embed are the real modules."""
    import importlib
    import types
    pe = importlib.import_module("models.pointbert.point_encoder")
    cfg = types.SimpleNamespace(trans_dim=384, depth=1, drop_path_rate=0.1, cls_dim=40, num_heads=6, group_size=32,
                                num_group=64, encoder_dims=256)
    model = pe.PointTransformer(cfg, args=types.SimpleNamespace()).eval()
    sd = torch_port.make_encoder_state()
    front = torch_port.make_front_end_state()
    missing = model.encoder.load_state_dict({k: v for k, v in sd.items() if k in torch_port.ENCODER_KEYS}, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    model.reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
    with torch.no_grad():
        model.cls_token.copy_(front["cls_token"])
        model.cls_pos.copy_(front["cls_pos"])
    model.pos_embed.load_state_dict({k[len("pos_embed."):]: v for k, v in front.items() if k.startswith("pos_embed.")})
    seen = {}
    model.blocks.register_forward_pre_hook(lambda m, a: seen.update(x=a[0].clone(), pos=a[1].clone()))
    xyz = cloud("U", 3, 1024, 4343)
    with refimport.fixed_fps_start(0), torch.no_grad():
        model(xyz)
        nb, center = model.group_divider(xyz)
    # the restatement agrees with the reference on the same tokens
    x2, pos2 = torch_port.assemble_forward(front, seen["x"][:, 1:], center)
    assert torch.equal(x2, seen["x"]) and (pos2 - seen["pos"]).abs().max() <= 1e-6 * seen["pos"].abs().max()
    np.savez_compressed(os.path.join(OUT, "front_end_small.npz"), xyz=xyz.numpy(), neighborhood=nb.numpy(),
                        center=center.numpy(), x=seen["x"].numpy(), pos=seen["pos"].numpy(),
                        torch_version=torch.__version__)
    print("front end x", tuple(seen["x"].shape), "pos absmax", float(seen["pos"].abs().max()))


def gen_dgcnn(ns):
    """The unmodified reference DGCNN_Propagation(k=4) (models/pointbert/pointnet2_utils.py:371-467) with seeded
    weights on seeded inputs (regenerated by the tests from the seeds): its output, for the fused edge-conv path."""
    dg = ns.pb_pn2.DGCNN_Propagation(k=4).eval()
    sd = torch_port.make_dgcnn_state(31)
    dg.load_state_dict(sd)
    rec = {}
    for tag, (B, Nk, Nq) in (("cross", (2, 64, 128)), ("up", (1, 96, 50))):
        coor, f, coor_q, f_q = torch_port.dgcnn_inputs(4700 + Nk, B, Nk, Nq)
        with torch.no_grad():
            out = dg(coor, f, coor_q, f_q)
            port = torch_port.dgcnn_forward(sd, coor, f, coor_q, f_q)
        assert (out - port).abs().max() <= 1e-5 * out.abs().max()
        rec[tag + ".out"] = out.numpy()
        rec[tag + ".inputs_sha"] = digest(torch.cat([t.reshape(-1) for t in (coor, f, coor_q, f_q)]).numpy())
    np.savez_compressed(os.path.join(OUT, "dgcnn.npz"), torch_version=torch.__version__, **rec)
    print("dgcnn fixtures", {k: getattr(v, "shape", v) for k, v in rec.items()})


def gen_bench_cfg2(ns):
    """BASELINE configs[1] at bench.py's own size and weights: rank 0's first batch (128 clouds x 8192 points,
    seed 1234) through the UNMODIFIED reference Group (dvae.py:152-181) + Encoder (dvae.py:184-215) + reduce_dim
    (point_encoder.py:133,239) carrying bench.make_tokenizer's seeded weights.  Stores digests of the index work
    (k-boundary tie rows zeroed, F6), the fp32 tokens of two clouds (every 4th group) and a per-group token
    checksum for all 65 536 groups; bench.check_cfg2_parity consumes it on the GPU box."""
    import bench
    tok = bench.make_tokenizer("fp16")
    xyz = bench.make_host_batches(0, 1, pin=False)[0]
    B, G, K = xyz.shape[0], bench.N_GROUP, bench.GROUP_SIZE
    enc = ns.dvae.Encoder(256).eval()
    enc.load_state_dict(tok.encoder.state_dict())
    reduce_dim = torch.nn.Linear(256, 384)
    reduce_dim.load_state_dict(tok.reduce_dim.state_dict())
    fps_all, center_all, knn_all, nb_all, tok_all, tie_all = [], [], [], [], [], []
    group = ns.dvae.Group(G, K)
    for lo in range(0, B, 16):
        part = xyz[lo:lo + 16]
        with refimport.fixed_fps_start(0), torch.no_grad():
            fps_all.append(ns.misc.farthest_point_sample(part, G))
            nb, center = group(part)
            sqd = ns.dvae.square_distance(center, part)
            knn = ns.dvae.knn_point(K, part, center)
            tok_all.append(reduce_dim(enc(nb)))
        tie_all.append(torch.from_numpy(knn_tie_rows(sqd, K)))
        center_all.append(center)
        knn_all.append(knn)
        nb_all.append(nb)
        print("bench_cfg2 clouds", lo + 16, flush=True)
    fps_idx, center, knn, nb, tokens, tie = (torch.cat(v) for v in (fps_all, center_all, knn_all, nb_all, tok_all, tie_all))
    assert torch.equal(center, torch_port.take_rows(xyz, fps_idx))
    knn_sorted = knn.sort(-1).values.clone()
    nb_c = canon_group(nb, knn).clone()
    knn_sorted[tie] = 0
    nb_c[tie] = 0
    clouds, step = np.array([0, B - 1]), 4
    wsum = hashlib.sha256(b"".join(np.ascontiguousarray(v.detach().numpy()).tobytes()
                                   for _, v in sorted(tok.state_dict().items()))).hexdigest()
    np.savez_compressed(os.path.join(OUT, "bench_cfg2.npz"), B=B, N=xyz.shape[1], G=G, K=K,
                        torch_version=torch.__version__, xyz_sha=digest(xyz.numpy()), weights_sha=wsum,
                        fps_sha=digest(fps_idx.numpy()), center_sha=digest(center.numpy()),
                        knn_sorted_sha=digest(knn_sorted.numpy()), nb_canon_sha=digest(nb_c.numpy()),
                        tie_rows=np.argwhere(tie.numpy()).astype(np.int32),
                        token_clouds=clouds, token_group_step=step, tokens=tokens[clouds][:, ::step].numpy(),
                        group_sum=tokens.double().sum(-1).float().numpy(), group_abs=tokens.double().abs().sum(-1).float().numpy())
    print("bench_cfg2: tie rows", int(tie.sum()), "token absmax", float(tokens.abs().max()))


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:  # regenerate one fixture only: front_end | encoder_train
        {"front_end": gen_front_end, "encoder_train": gen_encoder_train, "graph_feature": gen_graph_feature,
         "loader_fps": gen_loader_fps, "sa_mlp": gen_sa_mlp, "bench_cfg2": gen_bench_cfg2, "dgcnn": gen_dgcnn}[sys.argv[1]](refimport.load())
        return
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    ns = refimport.load()
    gen_sqdist(ns)
    gen_group(ns, "group_u1024", "U", 2, 1024, 128, 32, 1001, full=True)     # unordered-topk regime (F5, F11)
    gen_group(ns, "group_u1000_ragged", "U", 3, 1000, 37, 24, 1002, full=True)  # odd sizes, PointMLP k=24
    gen_group(ns, "group_s2048", "S", 1, 2048, 512, 32, 1003, full=True)      # part-seg shape
    gen_group(ns, "group_cfg1_u8192", "U", 32, 8192, 512, 32, 1234, full=False)  # BASELINE config 1 (U)
    gen_group(ns, "group_cfg1_s8192", "S", 32, 8192, 512, 32, 1235, full=False)  # BASELINE config 1 (S)
    gen_group(ns, "group_stress_s32768", "S", 1, 32768, 512, 32, 1239, full=False)  # cfg 5 stress
    gen_sa(ns, "sa_ssg_small", 2, 1024, 1237, full=True)
    gen_sa(ns, "sa_ssg_cfg3", 32, 1024, 1237, full=False)
    gen_msg_fp(ns, "msg_fp_small", 1, 2048, 1238, 16, full=True)
    gen_msg_fp(ns, "msg_fp_cfg4", 8, 2048, 1238, 384, full=False)
    gen_encoder(ns)
    gen_front_end(ns)
    gen_encoder_train(ns)
    gen_graph_feature(ns)
    gen_loader_fps(ns)
    gen_sa_mlp(ns)
    gen_bench_cfg2(ns)
    gen_dgcnn(ns)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("fixtures total bytes", tot)


if __name__ == "__main__":
    main()
