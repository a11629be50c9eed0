"""Torch-op restatement of the reference's CPU path -- TEST INFRASTRUCTURE ONLY.

The reference implements the tokenizer with ATen ops (matmul, topk, sort,
max, advanced indexing, Conv1d/BatchNorm1d), so the honest "reference on CPU"
timing is those same ops on CPU tensors.  /root/reference does not exist on
the GPU box, hence this port: the same op sequence, written independently,
validated bit-for-bit against the imported reference in the build container
(tests/test_oracle_vs_reference.py) and against tests/golden/.

Used for: bench.py's cpu_baseline / --impl reference legs (kind "port"), and
as the fp32 reference for the Encoder (a floating-point kernel keeps a torch
fp32 reference; tolerance is norm-relative).
"""
import torch
import torch.nn.functional as F


def pairwise_sqdist(src, dst):
    """models/pointbert/dvae.py:130-149 -- -2*src@dst^T + |src|^2 + |dst|^2, that order."""
    d = torch.matmul(src, dst.transpose(1, 2))
    d = d * -2
    d += (src ** 2).sum(-1).unsqueeze(2)
    d += (dst ** 2).sum(-1).unsqueeze(1)
    return d


def take_rows(points, idx):
    """index_points, models/pointbert/misc.py:26-42: points[b, idx[b, ...], :]."""
    B = points.shape[0]
    bsel = torch.arange(B, device=points.device).reshape((B,) + (1,) * (idx.dim() - 1))
    return points[bsel.expand_as(idx), idx, :]


def fps_indices(xyz, npoint, start):
    """models/pointbert/misc.py:44-69 with the start index supplied (the
    reference draws it with torch.randint, misc.py:59)."""
    B, N, _ = xyz.shape
    picked = xyz.new_zeros((B, npoint), dtype=torch.long)
    mind = xyz.new_full((B, N), 1e10)
    far = torch.as_tensor(start, dtype=torch.long, device=xyz.device).expand(B).clone()
    rows = torch.arange(B, device=xyz.device)
    for g in range(npoint):
        picked[:, g] = far
        c = xyz[rows, far].unsqueeze(1)
        d = ((xyz - c) ** 2).sum(-1)
        mind = torch.minimum(mind, d)
        far = mind.max(dim=-1).indices
    return picked


def knn_indices(k, xyz, query):
    """models/pointbert/dvae.py:116-127."""
    return pairwise_sqdist(query, xyz).topk(k, dim=-1, largest=False, sorted=False).indices


def group_forward(xyz, num_group, group_size, start=0):
    """Group.forward, models/pointbert/dvae.py:159-181 -> (neighborhood, center)."""
    B, N, _ = xyz.shape
    center = take_rows(xyz, fps_indices(xyz, num_group, start))
    idx = knn_indices(group_size, xyz, center)
    flat = (idx + torch.arange(B, device=xyz.device).view(B, 1, 1) * N).reshape(-1)
    nb = xyz.reshape(B * N, 3)[flat].reshape(B, num_group, group_size, 3)
    return nb - center.unsqueeze(2), center


def ball_indices(radius, nsample, xyz, query):
    """query_ball_point, models/pointnet2/pointnet2_utils.py:87-107."""
    B, N, _ = xyz.shape
    S = query.shape[1]
    ids = torch.arange(N, device=xyz.device).expand(B, S, N).clone()
    ids[pairwise_sqdist(query, xyz) > radius ** 2] = N
    ids = ids.sort(dim=-1).values[:, :, :nsample]
    first = ids[:, :, :1].expand(-1, -1, nsample)
    return torch.where(ids == N, first, ids)


def three_nn_interpolate(xyz1, xyz2, feats2):
    """Interpolation part of PointNetFeaturePropagation.forward,
    models/pointnet2/pointnet2_utils.py:297-307 (channel-last tensors)."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        return feats2.repeat(1, N, 1)
    d, i = pairwise_sqdist(xyz1, xyz2).sort(dim=-1)
    d, i = d[:, :, :3], i[:, :, :3]
    r = 1.0 / (d + 1e-8)
    w = r / r.sum(dim=2, keepdim=True)
    return (take_rows(feats2, i) * w.unsqueeze(-1)).sum(dim=2)


# ---- mini-PointNet patch Encoder (+ reduce_dim) ------------------------------

ENCODER_KEYS = (
    "first_conv.0.weight", "first_conv.0.bias",
    "first_conv.1.weight", "first_conv.1.bias", "first_conv.1.running_mean", "first_conv.1.running_var",
    "first_conv.3.weight", "first_conv.3.bias",
    "second_conv.0.weight", "second_conv.0.bias",
    "second_conv.1.weight", "second_conv.1.bias", "second_conv.1.running_mean", "second_conv.1.running_var",
    "second_conv.3.weight", "second_conv.3.bias",
)


def make_encoder_state(encoder_channel=256, trans_dim=384, seed=0):
    """Seeded weights in the reference Encoder's state_dict naming
    (models/pointbert/dvae.py:188-199) plus reduce_dim
    (models/pointbert/point_encoder.py:133).  Same init distributions torch
    uses for Conv1d/Linear, BN running stats perturbed so folding is exercised
    (SURVEY.md section 8d).  Deterministic for a given torch build; the
    fixture stores a checksum."""
    g = torch.Generator().manual_seed(seed)

    def conv(cout, cin):
        bound = 1.0 / (cin ** 0.5)
        w = (torch.rand(cout, cin, 1, generator=g) * 2 - 1) * bound
        b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        return w, b

    def bn(c):
        return (0.75 + 0.5 * torch.rand(c, generator=g), 0.1 * torch.randn(c, generator=g),
                0.1 * torch.randn(c, generator=g), 0.5 + torch.rand(c, generator=g))

    sd = {}
    sd["first_conv.0.weight"], sd["first_conv.0.bias"] = conv(128, 3)
    (sd["first_conv.1.weight"], sd["first_conv.1.bias"],
     sd["first_conv.1.running_mean"], sd["first_conv.1.running_var"]) = bn(128)
    sd["first_conv.3.weight"], sd["first_conv.3.bias"] = conv(256, 128)
    sd["second_conv.0.weight"], sd["second_conv.0.bias"] = conv(512, 512)
    (sd["second_conv.1.weight"], sd["second_conv.1.bias"],
     sd["second_conv.1.running_mean"], sd["second_conv.1.running_var"]) = bn(512)
    sd["second_conv.3.weight"], sd["second_conv.3.bias"] = conv(encoder_channel, 512)
    w, b = conv(trans_dim, encoder_channel)
    sd["reduce_dim.weight"], sd["reduce_dim.bias"] = w.squeeze(-1), b
    return sd


def encoder_forward(sd, point_groups, eps=1e-5):
    """Encoder.forward in eval mode, models/pointbert/dvae.py:201-215:
    (B,G,n,3) -> (B,G,C)."""
    bs, g, n, _ = point_groups.shape
    x = point_groups.reshape(bs * g, n, 3).transpose(2, 1)
    x = F.conv1d(x, sd["first_conv.0.weight"], sd["first_conv.0.bias"])
    x = F.batch_norm(x, sd["first_conv.1.running_mean"], sd["first_conv.1.running_var"],
                     sd["first_conv.1.weight"], sd["first_conv.1.bias"], False, 0.0, eps)
    x = F.conv1d(F.relu(x), sd["first_conv.3.weight"], sd["first_conv.3.bias"])
    glob = x.max(dim=2, keepdim=True).values
    x = torch.cat([glob.expand(-1, -1, n), x], dim=1)
    x = F.conv1d(x, sd["second_conv.0.weight"], sd["second_conv.0.bias"])
    x = F.batch_norm(x, sd["second_conv.1.running_mean"], sd["second_conv.1.running_var"],
                     sd["second_conv.1.weight"], sd["second_conv.1.bias"], False, 0.0, eps)
    x = F.conv1d(F.relu(x), sd["second_conv.3.weight"], sd["second_conv.3.bias"])
    return x.max(dim=2).values.reshape(bs, g, -1)


def tokens_forward(sd, point_groups):
    """Encoder then reduce_dim (models/pointbert/point_encoder.py:239): (B,G,n,3) -> (B,G,384)."""
    return F.linear(encoder_forward(sd, point_groups), sd["reduce_dim.weight"], sd["reduce_dim.bias"])


# ---- pos_embed + token assembly (models/pointbert/point_encoder.py:135-142, 241-247) --------------------
FRONT_KEYS = ("cls_token", "cls_pos", "pos_embed.0.weight", "pos_embed.0.bias", "pos_embed.2.weight", "pos_embed.2.bias")


def make_front_end_state(seed=1):
    """Seeded random-init cls_token / cls_pos / pos_embed with the reference's shapes (trans_dim 384)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: (torch.rand(*s, generator=g) * 2 - 1) * scale
    return {"cls_token": r(1, 1, 384, scale=0.5), "cls_pos": torch.randn(1, 1, 384, generator=g),
            "pos_embed.0.weight": r(128, 3, scale=3 ** -0.5), "pos_embed.0.bias": r(128, scale=3 ** -0.5),
            "pos_embed.2.weight": r(384, 128, scale=128 ** -0.5), "pos_embed.2.bias": r(384, scale=128 ** -0.5)}


def assemble_forward(front, tokens, center):
    """x = cat(cls_token, tokens), pos = cat(cls_pos, pos_embed(center)) -- point_encoder.py:241-247 restated."""
    B = tokens.shape[0]
    h = torch.nn.functional.linear(center, front["pos_embed.0.weight"], front["pos_embed.0.bias"])
    h = torch.nn.functional.gelu(h)  # nn.GELU() default: exact erf form
    p = torch.nn.functional.linear(h, front["pos_embed.2.weight"], front["pos_embed.2.bias"])
    x = torch.cat((front["cls_token"].expand(B, -1, -1), tokens), dim=1)
    pos = torch.cat((front["cls_pos"].expand(B, -1, -1), p), dim=1)
    return x, pos


# ---- Encoder under model.train(): batch-statistics BatchNorm (dvae.py:190,196; main_cls.py:169; F9) ---------
def encoder_forward_train(sd, neighborhood, momentum=0.1, eps=1e-5):
    """dvae.py:201-215 with both BatchNorm1d layers in training mode.  Returns (features [B,G,256], updated
    running statistics as a dict) without modifying `sd`."""
    F = torch.nn.functional
    bs, g, n, _ = neighborhood.shape
    new = {k: sd[k].clone() for k in sd if "running_" in k}
    x = neighborhood.reshape(bs * g, n, 3).transpose(2, 1)
    y = F.conv1d(x, sd["first_conv.0.weight"], sd["first_conv.0.bias"])
    y = F.batch_norm(y, new["first_conv.1.running_mean"], new["first_conv.1.running_var"], sd["first_conv.1.weight"],
                     sd["first_conv.1.bias"], True, momentum, eps)
    f = F.conv1d(F.relu(y), sd["first_conv.3.weight"], sd["first_conv.3.bias"])
    glob = torch.max(f, dim=2, keepdim=True)[0]
    y = F.conv1d(torch.cat([glob.expand(-1, -1, n), f], dim=1), sd["second_conv.0.weight"], sd["second_conv.0.bias"])
    y = F.batch_norm(y, new["second_conv.1.running_mean"], new["second_conv.1.running_var"], sd["second_conv.1.weight"],
                     sd["second_conv.1.bias"], True, momentum, eps)
    f = F.conv1d(F.relu(y), sd["second_conv.3.weight"], sd["second_conv.3.bias"])
    return torch.max(f, dim=2)[0].reshape(bs, g, -1), new


# ---- DGCNN_Propagation.get_graph_feature (models/pointbert/pointnet2_utils.py:392-442) ---------------------
def graph_feature(coor_q, x_q, coor_k, x_k, k):
    """Restated with explicit indexing: cat(x_k[b, :, idx] - x_q, x_q) -> [B, 2C, Nq, k]; returns (feature, idx)."""
    idx = knn_indices(k, coor_k.permute(0, 2, 1), coor_q.permute(0, 2, 1))          # [B, Nq, k]
    B, C, Nq = x_q.shape
    nb = torch.gather(x_k.unsqueeze(2).expand(B, C, Nq, x_k.shape[2]), 3, idx.unsqueeze(1).expand(B, C, Nq, k))
    xq = x_q.unsqueeze(-1).expand(-1, -1, -1, k)
    return torch.cat((nb - xq, xq), dim=1), idx


# ---- PointNet++ set-abstraction shared MLP (models/pointnet2/pointnet2_utils.py:161-266) ----------------------
def make_sa_state(in_channel, mlp, seed, conv_prefix="mlp_convs.", bn_prefix="mlp_bns."):
    """Seeded weights for one Conv2d(1x1)+BatchNorm2d stack with the reference's parameter names, including
    non-trivial running statistics (fresh modules have mean 0 / var 1, which would hide folding mistakes)."""
    g = torch.Generator().manual_seed(seed)
    sd, last = {}, in_channel
    for i, out in enumerate(mlp):
        bound = last ** -0.5
        sd["%s%d.weight" % (conv_prefix, i)] = (torch.rand(out, last, 1, 1, generator=g) * 2 - 1) * bound
        sd["%s%d.bias" % (conv_prefix, i)] = (torch.rand(out, generator=g) * 2 - 1) * bound
        sd["%s%d.weight" % (bn_prefix, i)] = torch.rand(out, generator=g) + 0.5
        sd["%s%d.bias" % (bn_prefix, i)] = torch.randn(out, generator=g) * 0.1
        sd["%s%d.running_mean" % (bn_prefix, i)] = torch.randn(out, generator=g) * 0.1
        sd["%s%d.running_var" % (bn_prefix, i)] = torch.rand(out, generator=g) + 0.5
        last = out
    return sd


def sa_mlp_max(grouped, sd, n_layers, conv_prefix="mlp_convs.", bn_prefix="mlp_bns.", eps=1e-5):
    """grouped [B, S, K, C0] -> 3 x relu(bn_eval(conv1x1)) -> max over K -> [B, C3, S]  (:196-201 restated)."""
    F = torch.nn.functional
    x = grouped.permute(0, 3, 2, 1)
    for i in range(n_layers):
        x = F.conv2d(x, sd["%s%d.weight" % (conv_prefix, i)], sd["%s%d.bias" % (conv_prefix, i)])
        x = F.batch_norm(x, sd["%s%d.running_mean" % (bn_prefix, i)], sd["%s%d.running_var" % (bn_prefix, i)],
                         sd["%s%d.weight" % (bn_prefix, i)], sd["%s%d.bias" % (bn_prefix, i)], False, 0.1, eps)
        x = F.relu(x)
    return torch.max(x, 2)[0]


# ---- DGCNN_Propagation (models/pointbert/pointnet2_utils.py:371-467) -------------------------------------------
def make_dgcnn_state(seed):
    """Seeded weights with the reference module's parameter names (layer1.0 / layer1.1 / layer2.0 / layer2.1);
    non-trivial GroupNorm affine parameters (fresh modules have weight 1 / bias 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cout, cin in (("layer1", 512, 768), ("layer2", 384, 1024)):
        sd[name + ".0.weight"] = (torch.rand(cout, cin, 1, 1, generator=g) * 2 - 1) * cin ** -0.5
        sd[name + ".1.weight"] = torch.rand(cout, generator=g) + 0.5
        sd[name + ".1.bias"] = torch.randn(cout, generator=g) * 0.1
    sd["layer1.1.weight"][::7] *= -1.0   # negative scales: max over neighbours must come after the affine map
    return sd


def dgcnn_inputs(seed, B, Nk, Nq, C=384):
    g = torch.Generator().manual_seed(seed)
    unit = lambda n: torch.nn.functional.normalize(torch.randn(B, 3, n, generator=g), dim=1)
    return unit(Nk), torch.randn(B, C, Nk, generator=g), unit(Nq), torch.randn(B, C, Nq, generator=g)


def dgcnn_forward(sd, coor, f, coor_q, f_q, k=4):
    """DGCNN_Propagation.forward restated with torch ops (edge features -> conv -> GroupNorm(4) -> LeakyReLU(0.2) ->
    max over k, twice; the second graph is the queries' own)."""
    F = torch.nn.functional

    def layer(name, cq, xq, ck, xk):
        feat, _ = graph_feature(cq, xq, ck, xk, k)
        y = F.conv2d(feat, sd[name + ".0.weight"])
        y = F.group_norm(y, 4, sd[name + ".1.weight"], sd[name + ".1.bias"], 1e-5)
        return F.leaky_relu(y, 0.2).max(dim=-1)[0]

    h = layer("layer1", coor_q, f_q, coor, f)
    return layer("layer2", coor_q, h, coor_q, h)
