"""Drop-in counterparts of models/pointnet2/pointnet2_utils.py (and the copy in
models/pointbert/pointnet2_utils.py) of the reference tree.

The geometry (FPS, ball query, grouping, three_nn / three_interpolate, DGCNN edge features) runs in the
sm_100a kernels.  The shared MLP + max-pool of the set-abstraction modules runs on the tensor cores in eval mode
(ops.sa_mlp_forward) and falls back to the module's own Conv/BatchNorm layers in train
mode or when a gradient is needed; the feature-propagation MLPs stay torch layers.  State dicts and outputs keep
the reference's layout: channel-first in, channel-first out.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoder_pack, ops
from .pointbert import _as_start, index_points, knn_point, square_distance  # noqa: F401  (same functions)


def farthest_point_sample(xyz, npoint, start_idx=None):
    """models/pointnet2/pointnet2_utils.py:63-84.  This copy draws the start index on the
    CPU and moves it (:75); the draw is repeated here so seeded runs stay aligned."""
    B, N, _ = xyz.shape
    if start_idx is None:
        start = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    else:
        start = _as_start(start_idx, xyz)
    return ops.fps(xyz, npoint, start)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """models/pointnet2/pointnet2_utils.py:87-107.  -> group_idx [B,S,nsample] int64."""
    from .pointbert import _cops
    c = _cops()
    return c.ball_query(float(radius), nsample, xyz, new_xyz) if c else ops.ball_query(radius, nsample, xyz, new_xyz)


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, start_idx=None):
    """models/pointnet2/pointnet2_utils.py:110-138.
    -> new_xyz [B,npoint,3], new_points [B,npoint,nsample,3+D] (SSG order: [xyz - centre, feats])."""
    fps_idx = farthest_point_sample(xyz, npoint, start_idx)
    new_xyz = ops.gather(xyz, fps_idx)
    idx = ops.ball_query(radius, nsample, xyz, new_xyz)
    new_points = ops.group_concat(xyz, new_xyz, points, idx, xyz_first=True)
    if returnfps:
        return new_xyz, new_points, ops.gather(xyz, idx), fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """models/pointnet2/pointnet2_utils.py:141-158 (no geometry: a view and a concat)."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        return new_xyz, torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped_xyz


def _fused_mlp_max(owner, key, xyz, points, new_xyz, idx, convs, bns, xyz_first):
    """The shared MLP + max-pool of a set-abstraction level on the tensor cores (ops.sa_mlp_forward), gather fused.
    Returns None when the fused path does not apply (train-mode BatchNorm, gradients needed, unsupported widths):
    the caller then runs the reference's own layer stack."""
    if owner.training or len(convs) != 3 or not xyz.is_cuda:
        return None
    params = [p for m in list(convs) + list(bns) for p in m.parameters()]
    if torch.is_grad_enabled() and (xyz.requires_grad or (points is not None and points.requires_grad)
                                    or any(p.requires_grad for p in params)):
        return None
    c0 = 3 + (0 if points is None else points.shape[-1])
    widths = [c.out_channels for c in convs]
    if convs[0].in_channels != c0 or not ops.sa_mlp_supported(c0, *widths, idx.shape[2]):
        return None
    mode = ops.ENC_MODES[getattr(owner, "ppt_precision", "fp16")]
    if mode not in (ops.ENC_FP16, ops.ENC_BF16):
        return None
    tensors = params + [b for m in bns for b in m.buffers()]
    ckey = (mode, str(xyz.device)) + tuple((t.data_ptr(), t._version) for t in tensors)
    cache = owner.__dict__.setdefault("_ppt_sa_cache", {})
    if key not in cache or cache[key][0] != ckey:
        blob, dims = encoder_pack.pack_sa_mlp(convs, bns, xyz_first, mode)
        cache[key] = (ckey, blob.to(xyz.device), dims)
    _, blob, dims = cache[key]
    return ops.sa_mlp_forward(xyz, points, new_xyz, idx, blob, dims, mode=mode,
                              per_layer=bool(getattr(owner, "ppt_sa_per_layer", False)))


def _shared_mlp_max(new_points, convs, bns):
    # [B, S, K, C] -> [B, C, K, S] -> Conv2d/BN/ReLU stack -> max over K  (:196-201)
    x = new_points.permute(0, 3, 2, 1)
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return torch.max(x, 2)[0]


class PointNetSetAbstraction(nn.Module):
    """models/pointnet2/pointnet2_utils.py:161-206."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, remove_last=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel
        self.group_all = group_all
        self.remove_last = remove_last
        self.start_idx = None

    def forward(self, xyz, points):
        """xyz [B,3,N], points [B,D,N] or None -> new_xyz [B,3,S], new_points [B,D',S]."""
        xyz = xyz.permute(0, 2, 1).contiguous()
        if points is not None:
            points = points.permute(0, 2, 1).contiguous()
        # eval mode: FPS / ball query, then gather + shared MLP + max-pool as one tensor-core pipeline (row f1)
        if self.group_all:
            new_xyz = torch.zeros(xyz.shape[0], 1, 3, device=xyz.device)
            idx = torch.arange(xyz.shape[1], device=xyz.device).view(1, 1, -1).expand(xyz.shape[0], 1, -1).contiguous()
        else:
            new_xyz = ops.gather(xyz, farthest_point_sample(xyz, self.npoint, self.start_idx))
            idx = ops.ball_query(self.radius, self.nsample, xyz, new_xyz)
        new_points = _fused_mlp_max(self, 0, xyz, points, new_xyz, idx, self.mlp_convs, self.mlp_bns, xyz_first=True)
        if new_points is None:
            if self.group_all:
                new_xyz, grouped = sample_and_group_all(xyz, points)
            else:
                grouped = ops.group_concat(xyz, new_xyz, points, idx, xyz_first=True)
            new_points = _shared_mlp_max(grouped, self.mlp_convs, self.mlp_bns)
        if getattr(self, "remove_last", False):  # the models/pointbert copy of this class has no such attribute
            return new_points
        return new_xyz.permute(0, 2, 1), new_points


class PointNetSetAbstractionMsg(nn.Module):
    """models/pointnet2/pointnet2_utils.py:209-266: one FPS, one ball query + grouping + MLP per radius;
    grouped features come first, centred coordinates last (:252)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for widths in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel + 3
            for out_channel in widths:
                convs.append(nn.Conv2d(last, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.start_idx = None

    def forward(self, xyz, points):
        xyz = xyz.permute(0, 2, 1).contiguous()
        if points is not None:
            points = points.permute(0, 2, 1).contiguous()
        new_xyz = ops.gather(xyz, farthest_point_sample(xyz, self.npoint, self.start_idx))
        outs = []
        for i, radius in enumerate(self.radius_list):
            idx = ops.ball_query(radius, self.nsample_list[i], xyz, new_xyz)
            out = _fused_mlp_max(self, i, xyz, points, new_xyz, idx, self.conv_blocks[i], self.bn_blocks[i],
                                 xyz_first=False)
            if out is None:
                grouped = ops.group_concat(xyz, new_xyz, points, idx, xyz_first=False)
                out = _shared_mlp_max(grouped, self.conv_blocks[i], self.bn_blocks[i])
            outs.append(out)
        return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)


def three_nn_interpolate(xyz1, xyz2, points2):
    """Interpolation part of PointNetFeaturePropagation.forward (:297-307), channel-last:
    xyz1 [B,N,3], xyz2 [B,S,3], points2 [B,S,D] -> [B,N,D].  Differentiable w.r.t. points2."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        return points2.repeat(1, N, 1)
    from .pointbert import _cops
    c = _cops()
    if c:
        dist, idx = c.three_nn(xyz1, xyz2)
        return c.three_interpolate(points2, idx, dist)
    dist, idx = ops.three_nn(xyz1, xyz2)
    return ops.three_interpolate(points2, idx, dist)


class PointNetFeaturePropagation(nn.Module):
    """models/pointnet2/pointnet2_utils.py:269-319."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        """xyz1 [B,3,N], xyz2 [B,3,S], points1 [B,D1,N] or None, points2 [B,D2,S] -> [B,D',N]."""
        fused = _fused_fp_mlp(self, xyz1, xyz2, points1, points2)
        if fused is not None:
            return fused
        interpolated = three_nn_interpolate(xyz1.permute(0, 2, 1).contiguous(), xyz2.permute(0, 2, 1).contiguous(),
                                            points2.permute(0, 2, 1).contiguous())
        if points1 is not None:
            new_points = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)
        else:
            new_points = interpolated
        new_points = new_points.permute(0, 2, 1)
        return feature_propagation_mlp(self, new_points)


def _fused_fp_mlp(module, xyz1, xyz2, points1, points2):
    """three_nn, then interpolation + concat + the two-layer MLP on the tensor cores (ops.fp_mlp_forward; SURVEY.md section
    8 row f4, dense half).  Eval mode with nothing to differentiate, two layers, <= 512 input and output channels, at
    least three known points; otherwise None and the caller runs the unfused path."""
    convs, bns = list(module.mlp_convs), list(module.mlp_bns)
    if module.training or len(convs) != 2 or not xyz1.is_cuda or xyz2.shape[2] < 3:
        return None
    params = [p for m in convs + bns for p in m.parameters()]
    tensors_in = [t for t in (xyz1, xyz2, points1, points2) if t is not None]
    if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors_in) or any(p.requires_grad for p in params)):
        return None
    if any(t.dtype != torch.float32 for t in tensors_in):
        return None
    d1 = 0 if points1 is None else points1.shape[1]
    c0, c1, c2 = d1 + points2.shape[1], convs[0].out_channels, convs[1].out_channels
    mode = ops.ENC_MODES[getattr(module, "ppt_precision", "fp16")]
    if convs[0].in_channels != c0 or mode not in (ops.ENC_FP16, ops.ENC_BF16) or not ops.fp_mlp_supported(c0, c1, c2):
        return None
    tensors = params + [b for m in bns for b in m.buffers()]
    key = (mode, str(xyz1.device)) + tuple((t.data_ptr(), t._version) for t in tensors)
    cache = module.__dict__.get("_ppt_fp_packed")
    if cache is None or cache[0] != key:
        blob, dims = encoder_pack.pack_fp_mlp(convs, bns, d1, mode)
        cache = (key, blob.to(xyz1.device), dims)
        module.__dict__["_ppt_fp_packed"] = cache
    dist, idx = ops.three_nn(xyz1.permute(0, 2, 1).contiguous(), xyz2.permute(0, 2, 1).contiguous())
    return ops.fp_mlp_forward(None if points1 is None else points1.contiguous(), points2.permute(0, 2, 1).contiguous(),
                              idx, dist, cache[1], cache[2], mode=mode)


def feature_propagation_mlp(module, x):
    """The Conv1d + BatchNorm1d + ReLU stack of PointNetFeaturePropagation (models/pointnet2/pointnet2_utils.py:316-319)
    on x [B, C, N].  Eval mode with nothing to differentiate: BatchNorm is folded into the convolution (one library
    GEMM with a fused bias per layer, then ReLU in place -- no normalisation pass, no extra activation tensor); the
    folded weights are cached per weight version.  Otherwise (training: these layers are trainable in the part-seg
    head, models/ULIP_models.py:550-565) the module's own layers run."""
    convs, bns = list(module.mlp_convs), list(module.mlp_bns)
    params = [p for m in convs + bns for p in m.parameters()]
    if module.training or (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))):
        for conv, bn in zip(convs, bns):
            x = F.relu(bn(conv(x)))
        return x
    tensors = params + [b for m in bns for b in m.buffers()]
    key = (str(x.device),) + tuple((t.data_ptr(), t._version) for t in tensors)
    cache = module.__dict__.get("_ppt_fp_folded")
    if cache is None or cache[0] != key:
        folded = []
        for conv, bn in zip(convs, bns):
            w, b = encoder_pack.fold_conv_bn(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                             bn.eps)
            folded.append((w.to(torch.float32).unsqueeze(-1).to(x.device), b.to(torch.float32).to(x.device)))
        cache = (key, folded)
        module.__dict__["_ppt_fp_folded"] = cache
    for w, b in cache[1]:
        x = F.relu_(F.conv1d(x, w, b))
    return x


class DGCNN_Propagation(nn.Module):
    """models/pointbert/pointnet2_utils.py:371-467 -- same sub-module names (layer1 / layer2: Conv2d, GroupNorm(4),
    LeakyReLU(0.2)), so part-seg checkpoints load."""

    def __init__(self, k=16):
        super().__init__()
        self.k = k
        self.layer1 = nn.Sequential(nn.Conv2d(768, 512, kernel_size=1, bias=False), nn.GroupNorm(4, 512),
                                    nn.LeakyReLU(negative_slope=0.2))
        self.layer2 = nn.Sequential(nn.Conv2d(1024, 384, kernel_size=1, bias=False), nn.GroupNorm(4, 384),
                                    nn.LeakyReLU(negative_slope=0.2))

    def get_graph_feature(self, coor_q, x_q, coor_k, x_k):
        return get_graph_feature(coor_q, x_q, coor_k, x_k, self.k)

    def forward(self, coor, f, coor_q, f_q):
        """coor [B,3,Nk], f [B,C,Nk], coor_q [B,3,Nq], f_q [B,C,Nq] -> [B,384,Nq]."""
        if dgcnn_fusable(self, coor, f, coor_q, f_q):
            return dgcnn_propagation_forward(self, coor, f, coor_q, f_q)
        x = self.layer1(self.get_graph_feature(coor_q, f_q, coor, f)).max(dim=-1, keepdim=False)[0]
        return self.layer2(self.get_graph_feature(coor_q, x, coor_q, x)).max(dim=-1, keepdim=False)[0]


def dgcnn_fusable(module, coor, f, coor_q, f_q):
    """The fused edge-conv path is forward only (no BatchNorm here, so train / eval does not matter)."""
    try:
        conv1, gn1, act1 = module.layer1
        conv2, gn2, act2 = module.layer2
    except Exception:
        return False
    params = list(module.parameters())
    if torch.is_grad_enabled() and (f.requires_grad or f_q.requires_grad or any(p.requires_grad for p in params)):
        return False
    ok_types = all(isinstance(c, nn.Conv2d) and c.bias is None and c.kernel_size == (1, 1) for c in (conv1, conv2)) and \
        all(isinstance(g, nn.GroupNorm) and g.affine for g in (gn1, gn2)) and \
        all(isinstance(a, nn.LeakyReLU) for a in (act1, act2))
    return (ok_types and f.is_cuda and f.dtype == torch.float32 and coor.shape[1] == 3 and coor_q.shape[1] == 3
            and conv1.in_channels == 2 * f.shape[1] and f_q.shape[1] == f.shape[1]
            and conv2.in_channels == 2 * conv1.out_channels and 1 <= module.k <= min(16, coor.shape[2], coor_q.shape[2]))


def _edge_layer(conv, gn, act, x_q, x_k, idx, half):
    """Conv2d(cat(x_k[idx] - x_q, x_q)) -> GroupNorm -> LeakyReLU -> max over neighbours, with the linear convolution
    split into U = Wa x_k and V = (Wb - Wa) x_q (two per-point library GEMMs; ops.edge_gn_max does the rest).
    half: fp16 operands / fp32 accumulate on the tensor cores (the fast mode, like the Encoder's; the reference's own
    cuDNN convolution runs TF32 by default), else plain fp32 GEMMs."""
    C = x_k.shape[1]
    w = conv.weight.reshape(conv.out_channels, 2 * C)
    wa, wd = w[:, :C], w[:, C:] - w[:, :C]
    cast = (lambda t: t.half()) if half else (lambda t: t)
    if x_q is x_k:
        uv = torch.matmul(cast(torch.cat([wa, wd], dim=0)), cast(x_k)).float()   # one GEMM for both
        U, V = uv[:, :conv.out_channels].contiguous(), uv[:, conv.out_channels:].contiguous()
    else:
        U, V = torch.matmul(cast(wa), cast(x_k)).float(), torch.matmul(cast(wd), cast(x_q)).float()
    return ops.edge_gn_max(U, V, idx, gn.weight, gn.bias, gn.num_groups, gn.eps, act.negative_slope)


@torch.no_grad()
def dgcnn_propagation_forward(module, coor, f, coor_q, f_q):
    """DGCNN_Propagation.forward (models/pointbert/pointnet2_utils.py:444-467) without the [B, 2C, Nq, k] edge tensor
    and with k times fewer convolution FLOPs: kNN kernels for the two graphs, per-point GEMMs, fused
    GroupNorm / LeakyReLU / max kernel (dense half; forward only)."""
    half = getattr(module, "ppt_precision", "fp16") != "fp32"
    cq = coor_q.permute(0, 2, 1).contiguous()
    idx1 = ops.knn(module.k, coor.permute(0, 2, 1).contiguous(), cq)
    h = _edge_layer(*module.layer1, f_q.contiguous(), f.contiguous(), idx1, half)
    idx2 = ops.knn(module.k, cq, cq)
    return _edge_layer(*module.layer2, h, h, idx2, half)


def get_graph_feature(coor_q, x_q, coor_k, x_k, k):
    """DGCNN_Propagation.get_graph_feature (models/pointbert/pointnet2_utils.py:392-442): kNN of every query
    among the keys (k = 4 in the part-seg head, point_encoder.py:303-304), then the edge features
    cat(x_k[idx] - x_q, x_q) -> [B, 2C, Nq, k].  coor_* [B,3,N*], x_* [B,C,N*] channel-first."""
    with torch.no_grad():
        idx = ops.knn(k, coor_k.permute(0, 2, 1).contiguous(), coor_q.permute(0, 2, 1).contiguous())
    return ops.graph_feature(x_q, x_k, idx)
