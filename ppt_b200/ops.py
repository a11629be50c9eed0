"""Tensor-level front end of the C ABI: validates, allocates outputs with torch,
launches on torch's current stream.  One function per entry point of
include/ppt_b200.h.  CUDA tensors only -- there is no CPU path here.
"""
import torch

from . import _lib

ENC_FP16, ENC_BF16, ENC_FP16X3 = 0, 1, 2
ENC_MODES = {"fp16": ENC_FP16, "bf16": ENC_BF16, "fp32": ENC_FP16X3, "fp16x3": ENC_FP16X3}


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ppt_b200 ops need CUDA tensors (got device %s); there is no CPU fallback" % t.device)


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _i64(t):
    return t.contiguous() if t.dtype == torch.int64 else t.long().contiguous()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


AUTO = "auto"


def _xyz3(name, t):
    """The geometry kernels are written for 3-D coordinates (stride 3 in the C ABI): anything else is an error here,
    never a silently wrong answer (feature-space kNN etc. must stay on the caller's torch code; patch.py does that)."""
    if t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError("%s must be [B, N, 3] (got %s)" % (name, tuple(t.shape)))


def spatial_index(xyz):
    """Builds the per-cloud spatial index used by fps / knn / knn_group to skip far-away rows of the
    cloud (results are bit-identical with and without it).  Returns None where it does not apply
    (N outside [512, 32768]; the bucketed FPS uses it up to 8192 points, the pruned kNN search up to 32768).
    The buffer is reused by the next call on the same device and stream: build, use, discard."""
    _need_cuda(xyz)
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    lib = _lib.load()
    nbytes = lib.ppt_spatial_index_bytes(B, N)
    if nbytes <= 0:
        return None
    buf = _workspace((xyz.device, "index"), nbytes)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.ppt_spatial_index_build(_ptr(xyz), _ptr(buf), B, N, _stream(xyz)), "ppt_spatial_index_build")
    return buf


FPS_INDEX_MAX_N = 8192  # csrc/spatial_index.cuh MAX_N: the bucketed FPS keeps a cloud's state in shared memory


def _resolve_index(index, xyz):
    return spatial_index(xyz) if isinstance(index, str) and index == AUTO else index


def fps(xyz, npoint, start, return_centers=False, index=AUTO):
    """farthest_point_sample with a given start index tensor [B] -> idx [B,npoint] (int64).
    index: AUTO (build one if it applies), None (plain kernel), or the result of spatial_index(xyz)."""
    _need_cuda(xyz, start)
    _xyz3("xyz", xyz)
    xyz = _f32(xyz)
    B, N, C = xyz.shape
    start = _i64(start)
    idx = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    centers = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if return_centers else None
    # AUTO: an index pays for itself from a handful of samples on, and only up to FPS_INDEX_MAX_N points (above that
    # ppt_fps runs its plain cluster kernel and would ignore it)
    use = npoint > 8 and N <= FPS_INDEX_MAX_N and npoint <= N
    index = (_resolve_index(index, xyz) if use else None) if isinstance(index, str) else index
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().ppt_fps(_ptr(xyz), _ptr(start), _ptr(idx), _ptr(centers), _ptr(index), B, N, npoint,
                                       _stream(xyz)), "ppt_fps")
    return (idx, centers) if return_centers else idx


def square_distance(src, dst):
    _need_cuda(src, dst)
    _xyz3("src", src), _xyz3("dst", dst)
    src, dst = _f32(src), _f32(dst)
    B, S, _ = src.shape
    N = dst.shape[1]
    out = torch.empty((B, S, N), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().ppt_square_distance(_ptr(src), _ptr(dst), _ptr(out), B, S, N, _stream(src)),
                   "ppt_square_distance")
    return out


def knn(k, xyz, query, return_dist=False, index=AUTO):
    _need_cuda(xyz, query)
    _xyz3("xyz", xyz), _xyz3("query", query)
    xyz, query = _f32(xyz), _f32(query)
    B, N, _ = xyz.shape
    S = query.shape[1]
    idx = torch.empty((B, S, k), dtype=torch.int64, device=xyz.device)
    dist = torch.empty((B, S, k), dtype=torch.float32, device=xyz.device) if return_dist else None
    index = _resolve_index(index, xyz)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().ppt_knn(_ptr(xyz), _ptr(query), _ptr(idx), _ptr(dist), _ptr(index), B, N, S, k,
                                       _stream(xyz)), "ppt_knn")
    return (idx, dist) if return_dist else idx


def knn_group(xyz, center, k, return_idx=False, index=AUTO):
    _need_cuda(xyz, center)
    _xyz3("xyz", xyz), _xyz3("center", center)
    xyz, center = _f32(xyz), _f32(center)
    B, N, _ = xyz.shape
    G = center.shape[1]
    nb = torch.empty((B, G, k, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.empty((B, G, k), dtype=torch.int64, device=xyz.device) if return_idx else None
    index = _resolve_index(index, xyz)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().ppt_knn_group(_ptr(xyz), _ptr(center), _ptr(nb), _ptr(idx), _ptr(index), B, N, G, k,
                                             _stream(xyz)), "ppt_knn_group")
    return (nb, idx) if return_idx else nb


def ball_query(radius, nsample, xyz, new_xyz):
    _need_cuda(xyz, new_xyz)
    _xyz3("xyz", xyz), _xyz3("new_xyz", new_xyz)
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    # torch compares the fp32 distances with the Python double radius**2 cast to fp32
    thr = float(torch.tensor(float(radius) ** 2, dtype=torch.float32))
    idx = torch.empty((B, S, nsample), dtype=torch.int64, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().ppt_ball_query(_ptr(xyz), _ptr(new_xyz), _ptr(idx), thr, B, N, S, nsample, _stream(xyz)),
                   "ppt_ball_query")
    return idx


def _gather_fwd(points, idx):
    B, N, C = points.shape
    M = idx[0].numel() if B else 0
    out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().ppt_gather(_ptr(points), _ptr(idx), _ptr(out), B, N, C, M, _stream(points)),
                   "ppt_gather")
    return out


def _scatter_rows(grad_rows, idx, N):
    """Adjoint of a row gather: grad_rows [B, M, C] accumulated into [B, N, C] at rows idx [B, M] (torch
    scatter_add: the backward is not on PPT's hot path -- its backbones are frozen -- it only has to be right)."""
    B, M, C = grad_rows.shape
    out = torch.zeros((B, N, C), dtype=grad_rows.dtype, device=grad_rows.device)
    return out.scatter_add_(1, idx.reshape(B, M, 1).expand(B, M, C), grad_rows)


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.N = points.shape[1]
        return _gather_fwd(points, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C = grad_out.shape[0], grad_out.shape[-1]
        return _scatter_rows(_f32(grad_out).reshape(B, -1, C), idx.reshape(B, -1), ctx.N), None


def gather(points, idx):
    """index_points: points [B,N,C], idx [B,...] -> [B,...,C].  Differentiable w.r.t. points."""
    _need_cuda(points, idx)
    points, idx = _f32(points), _i64(idx)
    if points.dim() != 3 or idx.dim() < 2 or idx.shape[0] != points.shape[0]:
        raise ValueError("gather: points [B,N,C], idx [B,...]")
    if points.requires_grad and torch.is_grad_enabled():
        return _Gather.apply(points, idx)
    return _gather_fwd(points, idx)


def _group_concat_fwd(xyz, new_xyz, points, idx, xyz_first):
    B, N, _ = xyz.shape
    _, S, K = idx.shape
    D = 0 if points is None else points.shape[2]
    out = torch.empty((B, S, K, 3 + D), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().ppt_group_concat(_ptr(xyz), _ptr(new_xyz), _ptr(points), _ptr(idx), _ptr(out), B, N, S,
                                                K, D, 1 if xyz_first else 0, _stream(xyz)), "ppt_group_concat")
    return out


class _GroupConcat(torch.autograd.Function):
    """out[b,s,j] = cat(xyz[b,idx] - new_xyz[b,s], points[b,idx]) (or features first): gradients for xyz, new_xyz
    and points, so an unfrozen PointNet++ backbone trains correctly through the grouping."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, points, idx, xyz_first):
        ctx.save_for_backward(idx)
        ctx.N, ctx.xyz_first, ctx.has_points = xyz.shape[1], xyz_first, points is not None
        return _group_concat_fwd(xyz, new_xyz, points, idx, xyz_first)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = _f32(grad_out)
        B, S, K, C = g.shape
        lo = 0 if ctx.xyz_first else C - 3
        g_xyz = g[..., lo:lo + 3]
        flat = idx.reshape(B, S * K)
        d_xyz = _scatter_rows(g_xyz.reshape(B, S * K, 3).contiguous(), flat, ctx.N)
        d_new = -g_xyz.sum(dim=2)
        d_pts = None
        if ctx.has_points:
            g_f = g[..., 3:] if ctx.xyz_first else g[..., :C - 3]
            d_pts = _scatter_rows(g_f.reshape(B, S * K, C - 3).contiguous(), flat, ctx.N)
        return d_xyz, d_new, d_pts, None, None


def group_concat(xyz, new_xyz, points, idx, xyz_first=True):
    _need_cuda(xyz, new_xyz, points, idx)
    _xyz3("xyz", xyz), _xyz3("new_xyz", new_xyz)
    xyz, new_xyz, idx = _f32(xyz), _f32(new_xyz), _i64(idx)
    if points is not None:
        points = _f32(points)
    if torch.is_grad_enabled() and (xyz.requires_grad or new_xyz.requires_grad
                                    or (points is not None and points.requires_grad)):
        return _GroupConcat.apply(xyz, new_xyz, points, idx, xyz_first)
    return _group_concat_fwd(xyz, new_xyz, points, idx, xyz_first)


def three_nn(unknown, known):
    _need_cuda(unknown, known)
    _xyz3("unknown", unknown), _xyz3("known", known)
    unknown, known = _f32(unknown), _f32(known)
    B, N, _ = unknown.shape
    S = known.shape[1]
    dist = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, N, 3), dtype=torch.int64, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.check(_lib.load().ppt_three_nn(_ptr(unknown), _ptr(known), _ptr(dist), _ptr(idx), B, N, S,
                                            _stream(unknown)), "ppt_three_nn")
    return dist, idx


def _interp_fwd(feats, idx, dist):
    B, S, D = feats.shape
    N = idx.shape[1]
    out = torch.empty((B, N, D), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.load().ppt_three_interpolate(_ptr(feats), _ptr(idx), _ptr(dist), _ptr(out), B, N, S, D,
                                                     _stream(feats)), "ppt_three_interpolate")
    return out


class _ThreeInterpolate(torch.autograd.Function):
    """Differentiable w.r.t. feats only (the part-seg head trains through it)."""

    @staticmethod
    def forward(ctx, feats, idx, dist):
        ctx.save_for_backward(idx, dist)
        ctx.S = feats.shape[1]
        return _interp_fwd(feats, idx, dist)

    @staticmethod
    def backward(ctx, grad_out):
        idx, dist = ctx.saved_tensors
        grad_out = _f32(grad_out)
        B, N, D = grad_out.shape
        grad_feats = torch.zeros((B, ctx.S, D), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _lib.check(_lib.load().ppt_three_interpolate_grad(_ptr(grad_out), _ptr(idx), _ptr(dist), _ptr(grad_feats),
                                                              B, N, ctx.S, D, _stream(grad_out)),
                       "ppt_three_interpolate_grad")
        return grad_feats, None, None


def three_interpolate(feats, idx, dist):
    """feats [B,S,D], idx/dist [B,N,3] -> [B,N,D]."""
    _need_cuda(feats, idx, dist)
    feats, idx, dist = _f32(feats), _i64(idx), _f32(dist)
    if feats.requires_grad and torch.is_grad_enabled():
        return _ThreeInterpolate.apply(feats, idx, dist)
    return _interp_fwd(feats, idx, dist)


def _graph_feature_fwd(x_q, x_k, idx):
    B, C, Nq = x_q.shape
    Nk, k = x_k.shape[2], idx.shape[2]
    out = torch.empty((B, 2 * C, Nq, k), dtype=torch.float32, device=x_q.device)
    if out.numel():
        with torch.cuda.device(x_q.device):
            _lib.check(_lib.load().ppt_graph_feature(_ptr(x_q), _ptr(x_k), _ptr(idx), _ptr(out), B, C, Nq, Nk, k,
                                                     _stream(x_q)), "ppt_graph_feature")
    return out


class _GraphFeature(torch.autograd.Function):
    """The part-seg head trains through the edge features (DGCNN_Propagation, point_encoder.py:409-411)."""

    @staticmethod
    def forward(ctx, x_q, x_k, idx):
        ctx.save_for_backward(idx)
        ctx.Nk = x_k.shape[2]
        return _graph_feature_fwd(x_q, x_k, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = _f32(grad_out)
        B, C2, Nq, k = grad_out.shape
        C = C2 // 2
        grad_xq = torch.empty((B, C, Nq), dtype=torch.float32, device=grad_out.device)
        grad_xk = torch.zeros((B, C, ctx.Nk), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _lib.check(_lib.load().ppt_graph_feature_grad(_ptr(grad_out), _ptr(idx), _ptr(grad_xq), _ptr(grad_xk), B, C,
                                                          Nq, ctx.Nk, k, _stream(grad_out)), "ppt_graph_feature_grad")
        return grad_xq, grad_xk, None


def graph_feature(x_q, x_k, idx):
    """Edge features of DGCNN_Propagation.get_graph_feature (models/pointbert/pointnet2_utils.py:418-442):
    x_q [B,C,Nq], x_k [B,C,Nk] channel-first, idx [B,Nq,k] int64 (neighbours of each query among the keys)
    -> [B, 2C, Nq, k] = cat(x_k[idx] - x_q, x_q).  Differentiable w.r.t. x_q and x_k."""
    _need_cuda(x_q, x_k, idx)
    x_q, x_k, idx = _f32(x_q), _f32(x_k), _i64(idx)
    if x_q.dim() != 3 or x_k.dim() != 3 or idx.dim() != 3 or x_q.shape[:2] != x_k.shape[:2] or \
            idx.shape[:2] != (x_q.shape[0], x_q.shape[2]):
        raise ValueError("graph_feature: x_q [B,C,Nq], x_k [B,C,Nk], idx [B,Nq,k]")
    if torch.is_grad_enabled() and (x_q.requires_grad or x_k.requires_grad):
        return _GraphFeature.apply(x_q, x_k, idx)
    return _graph_feature_fwd(x_q, x_k, idx)


def edge_gn_max(U, V, idx, gn_weight, gn_bias, num_groups, eps=1e-5, negative_slope=0.2):
    """The tail of a DGCNN_Propagation layer (models/pointbert/pointnet2_utils.py:382-390, 444-467) behind its two
    per-point GEMMs: U [B,Co,Nk] = Wa x_k, V [B,Co,Nq] = (Wb - Wa) x_q, idx [B,Nq,k] int64 ->
    max_j LeakyReLU(GroupNorm(U[:, :, idx[:, :, j]] + V)) : [B,Co,Nq].  Forward only."""
    _need_cuda(U, V, idx, gn_weight, gn_bias)
    U, V, idx = _f32(U), _f32(V), _i64(idx)
    B, C, Nk = U.shape
    Nq, k = idx.shape[1], idx.shape[2]
    if tuple(V.shape) != (B, C, Nq) or idx.shape[0] != B or gn_weight.numel() != C or gn_bias.numel() != C:
        raise ValueError("edge_gn_max: U [B,C,Nk], V [B,C,Nq], idx [B,Nq,k], weight / bias [C]")
    lib = _lib.load()
    out = torch.empty((B, C, Nq), dtype=torch.float32, device=U.device)
    if B == 0:
        return out
    ws = _workspace((U.device, "edge_gn"), lib.ppt_edge_gn_workspace_bytes(B, num_groups))
    with torch.cuda.device(U.device):
        _lib.check(lib.ppt_edge_gn_max_forward(_ptr(U), _ptr(V), _ptr(idx), _ptr(_f32(gn_weight)), _ptr(_f32(gn_bias)),
                                               _ptr(ws), _ptr(out), B, C, Nq, Nk, k, num_groups, float(eps),
                                               float(negative_slope), _stream(U)), "ppt_edge_gn_max_forward")
    return out


def fp_mlp_supported(c0, c1, c2):
    """Whether ppt_fp_mlp_forward covers a two-layer feature-propagation MLP c0 -> c1 -> c2 (c0 <= 512, c2 <= 512)."""
    return _lib.load().ppt_fp_mlp_packed_bytes(c0, c1, c2) > 0


def fp_mlp_forward(points1, feats2, idx, dist, packed, dims, mode=ENC_FP16):
    """three_interpolate + concat + two-layer Conv1d/BatchNorm1d/ReLU MLP of PointNetFeaturePropagation (eval mode,
    models/pointnet2/pointnet2_utils.py:304-319).  points1 [B,D1,N] channel-first or None, feats2 [B,S,D2]
    channel-last, idx / dist [B,N,3] from three_nn; `packed`, `dims` from encoder_pack.pack_fp_mlp -> [B, c2, N]."""
    _need_cuda(feats2, idx, dist, packed)
    feats2, idx, dist = _f32(feats2), _i64(idx), _f32(dist)
    B, S, D2 = feats2.shape
    N = idx.shape[1]
    c0, c1, c2 = dims
    D1 = c0 - D2
    if D1 > 0:
        _need_cuda(points1)
        points1 = _f32(points1)
        if tuple(points1.shape) != (B, D1, N):
            raise ValueError("points1 must be [B, %d, N] (channel-first)" % D1)
    elif D1 < 0:
        raise ValueError("packed MLP expects fewer input channels than feats2 has")
    if tuple(idx.shape) != (B, N, 3) or tuple(dist.shape) != (B, N, 3):
        raise ValueError("idx / dist must be [B, N, 3]")
    lib = _lib.load()
    if packed.dtype != torch.uint8 or packed.numel() != lib.ppt_fp_mlp_packed_bytes(c0, c1, c2):
        raise ValueError("packed FP-MLP blob does not match its dims")
    out = torch.empty((B, c2, N), dtype=torch.float32, device=feats2.device)
    ws = _workspace((feats2.device, "fp_mlp"), lib.ppt_fp_mlp_workspace_bytes(B * N, c0, c1, c2))
    with torch.cuda.device(feats2.device):
        _lib.check(lib.ppt_fp_mlp_forward(_ptr(points1) if D1 > 0 else None, _ptr(feats2), _ptr(idx), _ptr(dist), _ptr(packed),
                                          _ptr(ws), _ptr(out), B, N, S, D1, D2, c1, c2, mode, _stream(feats2)),
                   "ppt_fp_mlp_forward")
    return out


def sa_mlp_supported(c0, c1, c2, c3, nsample):
    """Whether ppt_sa_mlp_forward covers this layer stack (<= 512 input channels per layer, nsample 16/32/64/128)."""
    return nsample in (16, 32, 64, 128) and _lib.load().ppt_sa_mlp_packed_bytes(c0, c1, c2, c3) > 0


SA_PER_LAYER = 0x100  # PPT_SA_PER_LAYER


def sa_mlp_forward(xyz, feats, new_xyz, idx, packed, dims, mode=ENC_FP16, per_layer=False):
    """Grouping gather + 3 x (Conv 1x1 + BN + ReLU) + max over nsample of a PointNet++ set-abstraction level
    (models/pointnet2/pointnet2_utils.py:196-201, 256-261; eval mode).  xyz [B,N,3], feats [B,N,D] or None,
    new_xyz [B,S,3], idx [B,S,nsample] int64; `packed`, `dims` from encoder_pack.pack_sa_mlp -> [B, c3, S]."""
    _need_cuda(xyz, new_xyz, idx, packed)
    xyz, new_xyz, idx = _f32(xyz), _f32(new_xyz), _i64(idx)
    B, N, _ = xyz.shape
    S, ns = idx.shape[1], idx.shape[2]
    c0, c1, c2, c3 = dims
    D = c0 - 3
    if D > 0:
        feats = _f32(feats)
        if tuple(feats.shape) != (B, N, D):
            raise ValueError("feats must be [B, N, %d]" % D)
    lib = _lib.load()
    if packed.dtype != torch.uint8 or packed.numel() != lib.ppt_sa_mlp_packed_bytes(c0, c1, c2, c3):
        raise ValueError("packed SA-MLP blob does not match its dims")
    out = torch.empty((B, c3, S), dtype=torch.float32, device=xyz.device)
    ws = _workspace((xyz.device, "sa_mlp"), lib.ppt_sa_mlp_workspace_bytes(B * S * ns, c0, c1, c2, c3))
    with torch.cuda.device(xyz.device):
        _lib.check(lib.ppt_sa_mlp_forward(_ptr(xyz), _ptr(feats) if D > 0 else None, _ptr(new_xyz), _ptr(idx), _ptr(packed),
                                          _ptr(ws), _ptr(out), B, N, S, ns, D, c1, c2, c3,
                                          mode | (SA_PER_LAYER if per_layer else 0), _stream(xyz)),
                   "ppt_sa_mlp_forward")
    return out


def selftest_umma(a, b, mode=ENC_FP16, b_mn_major=False, a_packed=None, a_in_tmem=False):
    """D[128,N] = A[128,K] @ B[N,K]^T through the Encoder's tcgen05 building blocks."""
    _need_cuda(a, b)
    a, b = _f32(a), _f32(b)
    N, K = b.shape
    d = torch.empty((128, N), dtype=torch.float32, device=a.device)
    flags = mode | (4 if b_mn_major else 0) | (8 if a_packed is not None else 0) | (16 if a_in_tmem else 0)
    src = a if a_packed is None else a_packed
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().ppt_selftest_umma(_ptr(src), _ptr(b), _ptr(d), N, K, flags, _stream(a)),
                   "ppt_selftest_umma")
    return d


_workspaces = {}


def _workspace(key, nbytes):
    """Scratch owned by the host layer (the C ABI never allocates); grown on demand, one per
    (device, purpose, STREAM): kernels of one stream run in order, so reuse within a stream is safe, and two
    streams (or two threads driving their own streams) never share a buffer.  A buffer that is outgrown is
    released to torch's caching allocator, which keeps it reserved for this stream until its queued work is done."""
    dev = key[0]
    key = key + (torch.cuda.current_stream(dev).cuda_stream,)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def release_workspaces():
    """Drops every cached scratch buffer (e.g. after a stream has been destroyed)."""
    _workspaces.clear()


def encoder_packed_bytes(mode):
    return int(_lib.load().ppt_encoder_packed_bytes(mode))


TOKENS_F16 = 1  # PPT_TOKENS_F16
ENC_PHASE_NAMES = ("stage1", "group_linear_c", "stage2", "group_linear_tokens")


def encoder_forward(neighborhood, packed, mode=ENC_FP16, return_features=False, want_tokens=True, phase_events=None,
                    token_dtype=torch.float32, clock_acc=None, only_phase=None):
    """neighborhood [..., 32, 3] fp32 (CUDA) -> tokens [..., 384] (and Encoder features [..., 256]).

    `packed` is the uint8 CUDA blob from ppt_b200.encoder_pack.pack_encoder(state_dict, mode).
    `phase_events`: optional list; when given, the four launches are issued one by one and a
    (name, start_event, end_event) triple per launch is appended (for per-kernel timing on
    the launching stream).  `only_phase` (with phase_events): bracket just that launch (e.g. "stage2") and issue the
    launches before / after it together, so that timing the dominant kernel perturbs the step as little as possible.
    `token_dtype`: torch.float32 (the reference's dtype) or torch.float16 (PPT_TOKENS_F16: the same values
    rounded once more to fp16 in the last kernel's epilogue -- half the bytes to ship).
    `clock_acc`: optional int64[2] CUDA tensor (zeroed by the caller) that receives ns / SM cycles of the
    stage-2 kernel's CTA 0 (measurement aid, Stage2ClockTrace)."""
    _need_cuda(neighborhood, packed)
    nb = _f32(neighborhood)
    if nb.dim() < 3 or nb.shape[-1] != 3 or nb.shape[-2] != 32:
        raise ValueError("neighborhood must be [..., 32, 3]; the kernels are specialised for group_size 32 "
                         "(models/pointbert/PointTransformer_8192point.yaml:17-24)")
    lead = tuple(nb.shape[:-2])
    groups = 1
    for d in lead:
        groups *= d
    lib = _lib.load()
    if packed.dtype != torch.uint8 or packed.numel() != lib.ppt_encoder_packed_bytes(mode):
        raise ValueError("packed weight blob does not match mode %d" % mode)
    if not (want_tokens or return_features):
        raise ValueError("nothing to compute")
    if token_dtype not in (torch.float32, torch.float16):
        raise ValueError("token_dtype must be torch.float32 or torch.float16")
    flags = TOKENS_F16 if token_dtype == torch.float16 else 0
    tokens = torch.empty(lead + (384,), dtype=token_dtype, device=nb.device) if want_tokens else None
    feats = torch.empty(lead + (256,), dtype=torch.float32, device=nb.device) if return_features else None
    if groups == 0:
        return (tokens, feats) if return_features else tokens
    if clock_acc is not None and not (clock_acc.is_cuda and clock_acc.dtype == torch.int64 and clock_acc.numel() >= 2):
        raise ValueError("clock_acc must be a CUDA int64[2] tensor")
    ws = _workspace((nb.device, "encoder"), lib.ppt_encoder_workspace_bytes(groups, mode))
    with torch.cuda.device(nb.device):
        if phase_events is None:
            _lib.check(lib.ppt_encoder_forward_ex(_ptr(nb), _ptr(packed), _ptr(ws), _ptr(feats), _ptr(tokens), groups,
                                                  mode, 15, flags, _ptr(clock_acc), _stream(nb)), "ppt_encoder_forward")
        elif only_phase is not None:
            bit = ENC_PHASE_NAMES.index(only_phase)
            for mask, name in (((1 << bit) - 1, None), (1 << bit, only_phase), (15 & ~((2 << bit) - 1), None)):
                if not mask:
                    continue
                if name:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                _lib.check(lib.ppt_encoder_forward_ex(_ptr(nb), _ptr(packed), _ptr(ws), _ptr(feats), _ptr(tokens), groups,
                                                      mode, mask, flags, _ptr(clock_acc), _stream(nb)), "ppt_encoder_forward")
                if name:
                    e1.record()
                    phase_events.append((name, e0, e1))
        else:
            for bit, name in enumerate(ENC_PHASE_NAMES):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.ppt_encoder_forward_ex(_ptr(nb), _ptr(packed), _ptr(ws), _ptr(feats), _ptr(tokens), groups,
                                                      mode, 1 << bit, flags, _ptr(clock_acc), _stream(nb)),
                           "ppt_encoder_" + name)
                e1.record()
                phase_events.append((name, e0, e1))
    return (tokens, feats) if return_features else tokens


def encoder_forward_train(neighborhood, packed_train, bn, mode=ENC_FP16, return_features=False, want_tokens=True):
    """Encoder.forward under model.train() (batch-statistics BatchNorm, forward only).

    `packed_train`: the caller's own uint8 CUDA copy of encoder_pack.pack_encoder_train(state_dict, mode) -- it is
    written to.  `bn`: dict of the module's CUDA tensors conv1_weight, conv1_bias, bn{1,2}_{weight,bias,
    running_mean,running_var,num_batches_tracked} plus floats momentum, eps; the running statistics and counters
    are updated in place like torch.nn.BatchNorm1d does."""
    _need_cuda(neighborhood, packed_train)
    nb = _f32(neighborhood)
    if nb.dim() < 3 or nb.shape[-1] != 3 or nb.shape[-2] != 32:
        raise ValueError("neighborhood must be [..., 32, 3]")
    lead = tuple(nb.shape[:-2])
    groups = 1
    for d in lead:
        groups *= d
    lib = _lib.load()
    if packed_train.dtype != torch.uint8 or packed_train.numel() != lib.ppt_encoder_packed_bytes(mode):
        raise ValueError("packed weight blob does not match mode %d" % mode)
    if groups == 0 or not (want_tokens or return_features):
        raise ValueError("nothing to compute")
    st = _lib.EncoderBn()
    for name, _ in _lib.EncoderBn._fields_:
        v = bn[name]
        if name in ("momentum", "eps"):
            setattr(st, name, float(v))
            continue
        if v is None and name.endswith("num_batches_tracked"):
            setattr(st, name, None)
            continue
        want = torch.int64 if name.endswith("num_batches_tracked") else torch.float32
        if not (v.is_cuda and v.dtype == want and v.is_contiguous()):
            raise ValueError("bn[%r] must be a contiguous CUDA %s tensor" % (name, want))
        setattr(st, name, v.data_ptr())
    tokens = torch.empty(lead + (384,), dtype=torch.float32, device=nb.device) if want_tokens else None
    feats = torch.empty(lead + (256,), dtype=torch.float32, device=nb.device) if return_features else None
    ws = _workspace((nb.device, "encoder"), lib.ppt_encoder_train_workspace_bytes(groups, mode))
    import ctypes
    with torch.cuda.device(nb.device):
        _lib.check(lib.ppt_encoder_forward_train(_ptr(nb), _ptr(packed_train), ctypes.byref(st), _ptr(ws), _ptr(feats),
                                                 _ptr(tokens), groups, mode, _stream(nb)), "ppt_encoder_forward_train")
    return (tokens, feats) if return_features else tokens


def tokenizer_forward(neighborhood, center, enc_packed, pos_packed, mode=ENC_FP16, want_x=True):
    """The tokenizer's tail in one call (models/pointbert/point_encoder.py:239-247):

        x   = cat(cls_token, reduce_dim(encoder(neighborhood)), dim=1)     [B, G+1, 384]
        pos = cat(cls_pos,   pos_embed(center),                 dim=1)     [B, G+1, 384]

    neighborhood [B, G, 32, 3], center [B, G, 3] fp32 CUDA; `enc_packed` / `pos_packed` are the blobs of
    encoder_pack.pack_encoder / pack_pos_embed for `mode`.  want_x=False computes only pos."""
    _need_cuda(center, pos_packed)
    ct = _f32(center)
    if ct.dim() != 3 or ct.shape[-1] != 3:
        raise ValueError("center must be [B, G, 3]")
    B, G = int(ct.shape[0]), int(ct.shape[1])
    if G < 32:
        raise ValueError("token assembly needs num_group >= 32")
    lib = _lib.load()
    if pos_packed.dtype != torch.uint8 or pos_packed.numel() != lib.ppt_posembed_packed_bytes(mode):
        raise ValueError("packed pos_embed blob does not match mode %d" % mode)
    nb = None
    if want_x:
        _need_cuda(neighborhood, enc_packed)
        nb = _f32(neighborhood)
        if tuple(nb.shape) != (B, G, 32, 3):
            raise ValueError("neighborhood must be [B, G, 32, 3] matching center")
        if enc_packed.dtype != torch.uint8 or enc_packed.numel() != lib.ppt_encoder_packed_bytes(mode):
            raise ValueError("packed weight blob does not match mode %d" % mode)
    x = torch.empty((B, G + 1, 384), dtype=torch.float32, device=ct.device) if want_x else None
    pos = torch.empty((B, G + 1, 384), dtype=torch.float32, device=ct.device)
    if B == 0:
        return x, pos
    ws = _workspace((ct.device, "encoder"), lib.ppt_tokenizer_workspace_bytes(B * G, mode))
    with torch.cuda.device(ct.device):
        _lib.check(lib.ppt_tokenizer_forward(_ptr(nb), _ptr(ct), _ptr(enc_packed) if want_x else None, _ptr(pos_packed),
                                             _ptr(ws), _ptr(x), _ptr(pos), B * G, G, mode, _stream(ct)),
                   "ppt_tokenizer_forward")
    return x, pos


class Stage2ClockTrace:
    """SM clock inside the Encoder's stage-2 kernel, without perturbing it: pass `.acc` as encoder_forward's
    `clock_acc` and CTA 0 of every such launch adds its lifetime in ns and in SM cycles to it.  Use as a context
    manager; .mhz after.  (No library state: the accumulator travels with the call.)"""

    def __init__(self, device):
        self.acc = torch.zeros(2, dtype=torch.int64, device=device)
        self.mhz = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        torch.cuda.synchronize(self.acc.device)
        ns, cyc = (int(v) for v in self.acc.cpu())
        self.mhz = cyc / ns * 1e3 if ns > 0 else None
        return False


class ClockProbe:
    """SM clock seen INSIDE kernels: a one-thread kernel on a side stream samples (globaltimer, clock64) every
    `period_us`; mhz() after the work of interest has been synchronised.  CAUTION: the probe's CTA holds the 1 KB of
    reserved shared memory that a 227 KB CTA needs, so on its SM the Encoder stage kernels cannot be resident and
    run their last CTA as a second wave (about 1.7x slower): use it for clocks (tools/kernel_clocks.py), never
    inside a timed region -- bench.py uses Stage2ClockTrace instead."""

    def __init__(self, device, duration_ms=50.0, period_us=20.0):
        self.samples = max(2, int(duration_ms * 1e3 / period_us))
        self.buf = torch.zeros((self.samples, 2), dtype=torch.int64, device=device)
        self.stream = torch.cuda.Stream(device)
        self.period_ns = int(period_us * 1e3)

    def start(self):
        with torch.cuda.device(self.buf.device):
            _lib.check(_lib.load().ppt_clock_probe(_ptr(self.buf), self.samples, self.period_ns,
                                                   self.stream.cuda_stream), "ppt_clock_probe")
        return self

    def mhz(self):
        """-> (mean MHz, min MHz over 10-sample windows, number of samples) over the probe's lifetime."""
        self.stream.synchronize()
        b = self.buf.cpu()
        b = b[b[:, 0] > 0]
        if b.shape[0] < 12:
            return None
        mean = float(b[-1, 1] - b[0, 1]) / float(b[-1, 0] - b[0, 0]) * 1e3
        win = (b[10:, 1] - b[:-10, 1]).double() / (b[10:, 0] - b[:-10, 0]).double() * 1e3
        return mean, float(win.min()), float(win.max()), int(b.shape[0])
