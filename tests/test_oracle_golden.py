"""The CPU oracle (oracle/ppt_oracle.c, oracle/torch_port.py) against the
fixtures generated from the unmodified reference (oracle/gen_golden.py).
This is what makes the oracle's parity PINNED."""
import numpy as np
import pytest
import torch

from oracle import cpu, torch_port
from oracle.inputs import cloud, digest

SMALL_GROUP = ["group_u1024", "group_u1000_ragged", "group_s2048"]
BIG_GROUP = ["group_cfg1_u8192", "group_cfg1_s8192", "group_stress_s32768"]


def _mask_rows(shape, rows):
    m = np.zeros(shape, dtype=bool)
    for r in rows:
        m[tuple(r)] = True
    return m


def _canon(nb, idx):
    order = np.argsort(idx, axis=-1, kind="stable")
    return np.take_along_axis(nb, order[..., None], axis=2)


@pytest.mark.parametrize("name", SMALL_GROUP)
def test_group_small_cases_match_reference(golden, name):
    f = golden(name)
    xyz, G, K = f["xyz"], int(f["G"]), int(f["K"])
    assert digest(xyz) == str(f["xyz_sha"])
    nb, center, fidx, kidx = cpu.group_forward(xyz, G, K, start=0)
    assert np.array_equal(fidx, f["fps_idx"].astype(np.int64))
    assert np.array_equal(center.view(np.uint32), f["center"].view(np.uint32))
    tie = _mask_rows(kidx.shape[:2], f["tie_rows"])
    ks = np.sort(kidx, -1)
    assert np.array_equal(ks[~tie], f["knn_sorted"].astype(np.int64)[~tie])
    nbc = _canon(nb, kidx)
    assert np.array_equal(nbc[~tie].view(np.uint32), f["nb_canon"][~tie].view(np.uint32))


@pytest.mark.parametrize("name", BIG_GROUP)
def test_group_baseline_sizes_match_reference_digests(golden, name):
    f = golden(name)
    B, N, G, K = (int(f[k]) for k in "BNGK")
    xyz = cloud(str(f["kind"]), B, N, int(f["seed"])).numpy()
    assert digest(xyz) == str(f["xyz_sha"]), "torch RNG stream differs from the fixture's"
    nb, center, fidx, kidx = cpu.group_forward(xyz, G, K, start=0)
    assert digest(fidx) == str(f["fps_sha"])
    assert digest(center) == str(f["center_sha"])
    tie = _mask_rows(kidx.shape[:2], f["tie_rows"])
    if not tie.any():
        assert digest(np.sort(kidx, -1)) == str(f["knn_sorted_sha"])
        assert digest(_canon(nb, kidx)) == str(f["nb_canon_sha"])
    else:
        # F6: the reference resolves k-boundary ties arbitrarily; on those rows
        # compare the selected distance multiset, elsewhere demand the digest
        # after substituting the reference-agnostic rows out of both sides.
        t = torch.from_numpy(xyz)
        ref_idx = torch_port.knn_indices(K, t, torch.from_numpy(center)).numpy()
        ks, rs = np.sort(kidx, -1), np.sort(ref_idx, -1)
        assert digest(rs) == str(f["knn_sorted_sha"])  # the port reproduces the reference here
        assert np.array_equal(ks[~tie], rs[~tie])
        sd = cpu.square_distance(center, xyz)
        for b, g in np.argwhere(tie):
            assert np.array_equal(np.sort(sd[b, g, kidx[b, g]]), np.sort(sd[b, g, ref_idx[b, g]]))


def test_square_distance_bits(golden):
    f = golden("sqdist_small")
    d = cpu.square_distance(f["src"], f["dst"])
    assert np.array_equal(d.view(np.uint32), f["dist_bits"])
    dp = torch_port.pairwise_sqdist(torch.from_numpy(f["src"]), torch.from_numpy(f["dst"])).numpy()
    assert np.array_equal(dp.view(np.uint32), f["dist_bits"])
    assert (d < 0).any() or True  # negatives are legal (F3)


def _ssg(xyz, feats):
    f1 = cpu.farthest_point_sample(xyz, 512, 0)
    c1 = cpu.index_points(xyz, f1)
    b1 = cpu.query_ball_point(0.2, 32, xyz, c1)
    g1 = cpu.group_center(xyz, b1, c1)
    f2 = cpu.farthest_point_sample(c1, 128, 0)
    c2 = cpu.index_points(c1, f2)
    b2 = cpu.query_ball_point(0.4, 64, c1, c2)
    g2 = np.concatenate([cpu.group_center(c1, b2, c2), cpu.index_points(feats, b2)], -1)
    return f1, b1, g1, f2, b2, g2


@pytest.mark.parametrize("name", ["sa_ssg_small", "sa_ssg_cfg3"])
def test_set_abstraction_grouping_matches_reference(golden, name):
    f = golden(name)
    B, N, seed = int(f["B"]), int(f["N"]), int(f["seed"])
    xyz = cloud("S", B, N, seed).numpy()
    assert digest(xyz) == str(f["xyz_sha"])
    feats = torch.randn(B, 512, 128, generator=torch.Generator().manual_seed(seed + 100)).numpy()
    assert digest(feats) == str(f["feats_sha"])
    f1, b1, g1, f2, b2, g2 = _ssg(xyz, feats)
    for got, key in ((f1, "fps1"), (b1, "ball1"), (g1, "grp1"), (f2, "fps2"), (b2, "ball2"), (g2, "grp2")):
        assert digest(got) == str(f[key + "_sha"]), key
    if "ball1" in f.files:
        assert np.array_equal(b1, f["ball1"].astype(np.int64))


@pytest.mark.parametrize("name", ["msg_fp_small", "msg_fp_cfg4"])
def test_msg_ball_and_feature_propagation_match_reference(golden, name):
    f = golden(name)
    B, N, D, seed = int(f["B"]), int(f["N"]), int(f["D"]), int(f["seed"])
    xyz = cloud("S", B, N, seed).numpy()
    assert digest(xyz) == str(f["xyz_sha"])
    f1 = cpu.farthest_point_sample(xyz, 512, 0)
    assert digest(f1) == str(f["fps1_sha"])
    c1 = cpu.index_points(xyz, f1)
    for r, k in ((0.1, 16), (0.2, 32), (0.4, 128)):
        assert digest(cpu.query_ball_point(r, k, xyz, c1)) == str(f["ball_%g_%d_sha" % (r, k)])
    feats = torch.randn(B, 512, D, generator=torch.Generator().manual_seed(seed + 100)).numpy()
    assert digest(feats) == str(f["feats_sha"])
    dist, idx = cpu.three_nn(xyz, c1)
    assert len(f["nn_tie_rows"]) == 0
    assert digest(dist) == str(f["nn_dist_sha"])
    assert digest(idx) == str(f["nn_idx_sha"])
    out = cpu.three_interpolate(feats, idx, dist)
    assert digest(out) == str(f["interp_sha"])  # bit-exact (F8), including negative-d rows (F3)
    assert int(f["n_negative"]) > 0
    outp = torch_port.three_nn_interpolate(torch.from_numpy(xyz), torch.from_numpy(c1), torch.from_numpy(feats))
    assert digest(outp) == str(f["interp_sha"])


def test_ball_query_without_survivor_pads_with_N():
    xyz = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    q = np.array([[[10, 10, 10]]], dtype=np.float32)
    assert (cpu.query_ball_point(0.1, 4, xyz, q) == 3).all()  # reference sentinel N


def test_fps_tie_breaks_on_first_index():
    # four corners of a square: after picking corner 0, corners 1 and 2 tie; torch.max keeps the first
    xyz = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]], dtype=np.float32)
    got = cpu.farthest_point_sample(xyz, 4, 0)
    ref = torch_port.fps_indices(torch.from_numpy(xyz), 4, 0).numpy()
    assert np.array_equal(got, ref)
    assert got[0, 1] == 3 and got[0, 2] == 1


def test_encoder_port_matches_reference_tokens(golden):
    f = golden("encoder_small")
    sd = torch_port.make_encoder_state()
    import hashlib
    wsum = hashlib.sha256(b"".join(np.ascontiguousarray(sd[k].numpy()).tobytes() for k in sorted(sd))).hexdigest()
    assert wsum == str(f["weights_sha"]), "seeded weights differ from the fixture's (torch build changed?)"
    with torch.no_grad():
        nb = torch.from_numpy(f["neighborhood"])
        feat = torch_port.encoder_forward(sd, nb).numpy()
        tok = torch_port.tokens_forward(sd, nb).numpy()
    for got, ref in ((feat, f["features"]), (tok, f["tokens"])):
        rel = np.abs(got - ref).max() / np.abs(ref).max()
        assert rel <= 1e-6, rel


def test_token_assembly_port_matches_reference_point_transformer(golden):
    """x / pos handed to self.blocks by the unmodified PointTransformer (point_encoder.py:241-249)."""
    f = golden("front_end_small")
    sd, front = torch_port.make_encoder_state(), torch_port.make_front_end_state()
    with torch.no_grad():
        nb, center = torch.from_numpy(f["neighborhood"]), torch.from_numpy(f["center"])
        B, G = center.shape[:2]
        tok = torch_port.tokens_forward(sd, nb)
        x, pos = torch_port.assemble_forward(front, tok, center)
    assert x.shape == (B, G + 1, 384) and pos.shape == (B, G + 1, 384)
    for got, ref in ((x.numpy(), f["x"]), (pos.numpy(), f["pos"])):
        assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-6
    assert np.array_equal(x[:, 0].numpy(), f["x"][:, 0]) and np.array_equal(pos[:, 0].numpy(), f["pos"][:, 0])


def test_train_mode_encoder_port_matches_reference(golden):
    """Reference Encoder under .train(): batch-statistics BatchNorm and its running-stat update (F9)."""
    f = golden("encoder_train_small")
    sd = torch_port.make_encoder_state()
    with torch.no_grad():
        feat, stats = torch_port.encoder_forward_train(sd, torch.from_numpy(f["neighborhood"]))
    assert np.abs(feat.numpy() - f["features"]).max() / np.abs(f["features"]).max() <= 1e-6
    for k, v in stats.items():
        ref = f["after." + k]
        assert np.abs(v.numpy() - ref).max() <= 1e-6 * np.abs(ref).max(), k
    assert int(f["after.first_conv.1.num_batches_tracked"]) == 1


def test_graph_feature_port_matches_reference(golden):
    """DGCNN_Propagation.get_graph_feature (row f4): neighbour order inside k is unspecified -> per-row sorted."""
    from oracle.inputs import digest
    f = golden("graph_feature")
    for tag in ("small", "partseg"):
        args = [torch.from_numpy(f[tag + "." + n]) for n in ("coor_q", "x_q", "coor_k", "x_k")]
        feat, idx = torch_port.graph_feature(*args, 4)
        if tag == "small":
            assert np.array_equal(feat.sort(-1)[0].numpy(), f["small.feature_sorted"])
            assert np.array_equal(idx.sort(-1)[0].numpy(), f["small.idx_sorted"])
        else:
            assert digest(feat.sort(-1)[0]) == str(f["partseg.feature_sorted_sha"])
        self_feat, _ = torch_port.graph_feature(args[0], args[1], args[0], args[1], 4)
        assert digest(self_feat.sort(-1)[0]) == str(f[tag + ".self_feature_sorted_sha"])


def test_loader_fps_port_matches_reference(golden):
    """data/dataset_3d.py:40-61 (row f4): the fixture was recorded by executing the reference function's own source."""
    f = golden("loader_fps")
    for tag in ("a", "b"):
        idx = cpu.loader_fps_indices(f[tag + ".point"], int(f[tag + ".npoint"]), int(f[tag + ".start"]))
        assert np.array_equal(idx, f[tag + ".indices"])
        # and the C oracle's FPS (the kernels' checker) agrees with it: one semantics for all FPS copies (F13)
        xyz = np.ascontiguousarray(f[tag + ".point"][None, :, :3])
        c_idx = cpu.farthest_point_sample(xyz, int(f[tag + ".npoint"]), np.array([int(f[tag + ".start"])]))
        assert np.array_equal(np.asarray(c_idx)[0], f[tag + ".indices"])


def test_sa_mlp_port_matches_reference(golden):
    """Set-abstraction shared MLP + max-pool (row f1) of the unmodified reference modules in eval mode."""
    f = golden("sa_mlp")
    # SSG level 2 through the oracle's geometry + the restated layer stack
    xyz, feats = f["ssg2.xyz"], f["ssg2.feats"]
    fi = cpu.farthest_point_sample(xyz, 128, 0)
    c = cpu.index_points(xyz, fi)
    b = cpu.query_ball_point(0.4, 64, xyz, c)
    grouped = np.concatenate([cpu.group_center(xyz, b, c), cpu.index_points(feats, b)], -1)
    out = torch_port.sa_mlp_max(torch.from_numpy(grouped), torch_port.make_sa_state(131, [128, 128, 256], 11), 3)
    assert np.array_equal(np.transpose(c, (0, 2, 1)), f["ssg2.new_xyz"])
    assert np.abs(out.numpy() - f["ssg2.out"]).max() <= 1e-5 * np.abs(f["ssg2.out"]).max()
    # group_all
    g3 = np.concatenate([f["ssg3.xyz"], f["ssg3.feats"]], -1)[:, None]
    out3 = torch_port.sa_mlp_max(torch.from_numpy(g3), torch_port.make_sa_state(259, [256, 512, 1024], 12), 3)
    assert np.abs(out3.numpy() - f["ssg3.out"]).max() <= 1e-5 * np.abs(f["ssg3.out"]).max()
