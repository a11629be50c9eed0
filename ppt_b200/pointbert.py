"""Drop-in counterparts of models/pointbert/misc.py and models/pointbert/dvae.py (reference tree).

Same names, argument order, dtypes and return shapes as the reference; the bodies launch
the sm_100a kernels through ppt_b200.ops.  Additive keywords (never required): `start_idx`
to pin the FPS start instead of drawing it.
"""
import torch
import torch.nn as nn

from . import encoder_pack, ops


def _cops():
    """torch.ops.ppt_b200.* while a tracer (torch.compile) is recording, else None: eager calls go straight to ctypes."""
    from . import custom_ops
    return custom_ops if custom_ops.use_custom_ops() else None


def _draw_start(xyz):
    """The reference's own draw (models/pointbert/misc.py:59), so a seeded run consumes
    the RNG stream exactly as it does with the reference."""
    B, N, _ = xyz.shape
    return torch.randint(0, N, (B,), dtype=torch.long, device=xyz.device)


def farthest_point_sample(xyz, npoint, start_idx=None):
    """models/pointbert/misc.py:44-69.  xyz [B,N,3] -> centroids [B,npoint] int64."""
    start = _draw_start(xyz) if start_idx is None else _as_start(start_idx, xyz)
    c = _cops()
    return c.fps(xyz, npoint, start) if c else ops.fps(xyz, npoint, start)


def _as_start(start_idx, xyz):
    if isinstance(start_idx, int):
        # a fill kernel, not a host-to-device copy: torch.as_tensor(int, device=cuda) is a synchronous pageable copy that
        # would stall the host behind everything queued on the stream (it cost the host->host pipeline its overlap)
        return torch.full((xyz.shape[0],), start_idx, dtype=torch.long, device=xyz.device)
    s = torch.as_tensor(start_idx, dtype=torch.long, device=xyz.device)
    return s.expand(xyz.shape[0]).contiguous() if s.dim() == 0 else s


def index_points(points, idx):
    """models/pointbert/misc.py:26-42.  points [B,N,C], idx [B,S] or [B,S,K] -> [B,S(,K),C]."""
    c = _cops()
    return c.gather(points, idx) if c else ops.gather(points, idx)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def fps(data, number, start_idx=None):
    """models/pointbert/misc.py:12-24.  data [B,N,3] -> fps_data [B,number,3]."""
    start = _draw_start(data) if start_idx is None else _as_start(start_idx, data)
    if _needs_grad(data):  # the gather is the differentiable part (index_points, misc.py:23)
        return ops.gather(data, ops.fps(data.detach(), number, start))
    return ops.fps(data, number, start, return_centers=True)[1]


def square_distance(src, dst):
    """models/pointbert/dvae.py:130-149.  src [B,N,C], dst [B,M,C] -> [B,N,M] (C = 3)."""
    if _needs_grad(src, dst):  # the kernel is forward only: keep autograd's graph with the reference's own formula
        dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
        dist = dist + torch.sum(src ** 2, -1).unsqueeze(-1)
        return dist + torch.sum(dst ** 2, -1).unsqueeze(1)
    return ops.square_distance(src, dst)


def knn_point(nsample, xyz, new_xyz):
    """models/pointbert/dvae.py:116-127.  -> group_idx [B,S,nsample] int64.  The reference's
    topk(sorted=False) leaves the order unspecified; here it is ascending (distance, index)."""
    c = _cops()
    return c.knn(nsample, xyz, new_xyz) if c else ops.knn(nsample, xyz, new_xyz)


def group_forward(xyz, num_group, group_size, start):
    """Group.forward behind the start-index draw (dvae.py:163-180).  Indices come from the kernels either way; when
    the coordinates need a gradient the two gathers go through the differentiable ops.gather."""
    if _needs_grad(xyz):
        x = xyz.detach()
        index = ops.spatial_index(x)
        center = ops.gather(xyz, ops.fps(x, num_group, start, index=index))
        idx = ops.knn(group_size, x, center.detach(), index=index)
        return ops.gather(xyz, idx) - center.unsqueeze(2), center
    c = _cops()
    if c:
        return c.group(xyz, num_group, group_size, start)
    index = ops.spatial_index(xyz)  # one index serves both FPS and kNN
    _, center = ops.fps(xyz, num_group, start, return_centers=True, index=index)
    return ops.knn_group(xyz, center, group_size, index=index), center


class Group(nn.Module):
    """models/pointbert/dvae.py:152-181: FPS centres, kNN patches, centred coordinates."""

    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size
        self.start_idx = None  # additive: set to an int / [B] tensor for a deterministic start

    def forward(self, xyz):
        """xyz [B,N,3] -> neighborhood [B,G,M,3], center [B,G,3]."""
        start = _draw_start(xyz) if self.start_idx is None else _as_start(self.start_idx, xyz)
        return group_forward(xyz, self.num_group, self.group_size, start)


def train_forward_fusable(module, point_groups, reduce_dim=None):
    """The fused batch-statistics path is forward only: usable when no tensor involved needs a gradient."""
    bn1, bn2 = module.first_conv[1], module.second_conv[1]
    params = list(module.parameters()) + (list(reduce_dim.parameters()) if reduce_dim is not None else [])
    needs_grad = torch.is_grad_enabled() and (point_groups.requires_grad or any(p.requires_grad for p in params))
    return (point_groups.is_cuda and not needs_grad and point_groups.dim() == 4 and point_groups.shape[2] == 32
            and module.encoder_channel == 256 and point_groups.shape[0] * point_groups.shape[1] * 32 > 1
            and all(isinstance(b, nn.BatchNorm1d) and b.track_running_stats and b.momentum is not None and b.affine
                    for b in (bn1, bn2)) and bn1.eps == bn2.eps and bn1.momentum == bn2.momentum)


def train_forward(module, point_groups, state_for_pack, mode, want_tokens):
    """Encoder.forward (models/pointbert/dvae.py:201-215) with its BatchNorms in batch-statistics mode, on any
    module with the reference's layout (first_conv / second_conv Sequentials).  Returns tokens [B,G,384]
    (reduce_dim fused) or the Encoder's own features [B,G,256]."""
    conv1, bn1, bn2 = module.first_conv[0], module.first_conv[1], module.second_conv[1]
    static = [p for n, p in module.named_parameters() if not n.startswith(("first_conv.1.", "second_conv.1."))]
    key = (mode, str(point_groups.device)) + tuple((t.data_ptr(), t._version) for t in static)
    if getattr(module, "_ppt_train_key", None) != key:
        blob = encoder_pack.pack_encoder_train(state_for_pack(), mode).to(point_groups.device)
        object.__setattr__(module, "_ppt_train_blob", blob)
        object.__setattr__(module, "_ppt_train_key", key)
    bn = {"conv1_weight": conv1.weight, "conv1_bias": conv1.bias,
          "bn1_weight": bn1.weight, "bn1_bias": bn1.bias, "bn1_running_mean": bn1.running_mean,
          "bn1_running_var": bn1.running_var, "bn1_num_batches_tracked": bn1.num_batches_tracked,
          "bn2_weight": bn2.weight, "bn2_bias": bn2.bias, "bn2_running_mean": bn2.running_mean,
          "bn2_running_var": bn2.running_var, "bn2_num_batches_tracked": bn2.num_batches_tracked,
          "momentum": bn1.momentum, "eps": bn1.eps}
    out = ops.encoder_forward_train(point_groups, module._ppt_train_blob, bn, mode=mode, return_features=not want_tokens,
                                    want_tokens=want_tokens)
    object.__setattr__(module, "_bn_epoch", getattr(module, "_bn_epoch", 0) + 1)  # eval blobs are stale now
    return out if want_tokens else out[1]


class Encoder(nn.Module):
    """models/pointbert/dvae.py:184-215 -- same sub-module / state_dict names, so the ULIP
    checkpoints keep loading (models/ULIP_models.py:487-507).

    eval(): BatchNorm is folded and the whole stack runs in the tcgen05 kernels.
    train(): the reference puts the frozen Encoder's BatchNorm in batch-statistics mode
    (main_cls.py:169).  When nothing needs a gradient (PPT freezes the Encoder,
    models/ULIP_models.py:505) that runs fused too (ops.encoder_forward_train: batch statistics, running-stat
    momentum update in place); otherwise the module's own torch layers run on the GPU.
    """

    def __init__(self, encoder_channel, precision="fp16"):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                                        nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, self.encoder_channel, 1))
        self.precision = precision
        self._reduce_dim = None  # optional nn.Linear fused behind the Encoder (point_encoder.py:133)
        self._packed = None
        self._packed_key = None

    # -- additive API -----------------------------------------------------------------------------
    def attach_reduce_dim(self, linear):
        """Registers the caller's reduce_dim Linear (not as a sub-module: its parameters stay
        owned by the caller) so forward_tokens() can fuse it."""
        object.__setattr__(self, "_reduce_dim", linear)
        self._packed = None
        return self

    def _state_for_pack(self):
        sd = {k: v for k, v in self.state_dict().items()}
        if self._reduce_dim is not None:
            sd["reduce_dim.weight"], sd["reduce_dim.bias"] = self._reduce_dim.weight, self._reduce_dim.bias
        else:
            ref = sd["first_conv.0.weight"]
            sd["reduce_dim.weight"] = ref.new_zeros((384, 256))
            sd["reduce_dim.bias"] = ref.new_zeros((384,))
        return sd

    def _blob(self, device):
        mode = ops.ENC_MODES[self.precision]
        tensors = list(self.parameters()) + list(self.buffers())
        if self._reduce_dim is not None:
            tensors += [self._reduce_dim.weight, self._reduce_dim.bias]
        # _bn_epoch: the fused train path updates the running statistics from a kernel (no version bump)
        key = (mode, str(device), getattr(self, "_bn_epoch", 0)) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is None or self._packed_key != key:
            self._packed = encoder_pack.pack_encoder(self._state_for_pack(), mode).to(device)
            self._packed_key = key
        return self._packed, mode

    def _grad_needed(self, point_groups):
        params = list(self.parameters()) + (list(self._reduce_dim.parameters()) if self._reduce_dim is not None else [])
        return _needs_grad(point_groups, *params)

    def invalidate_packed(self):
        """Drops the cached packed weights.  The cache is keyed on (data_ptr, _version) of every parameter and
        buffer, which misses writes through `param.data` (the reference's loading idiom, models/ULIP_models.py:507):
        call this after such a load if a forward has already run (load_state_dict does it by itself)."""
        self._packed = None
        self._packed_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def _train_fusable(self, point_groups):
        return train_forward_fusable(self, point_groups, self._reduce_dim)

    def _forward_torch(self, point_groups):
        bs, g, n, _ = point_groups.shape
        feature = self.first_conv(point_groups.reshape(bs * g, n, 3).transpose(2, 1))
        glob = torch.max(feature, dim=2, keepdim=True)[0]
        feature = self.second_conv(torch.cat([glob.expand(-1, -1, n), feature], dim=1))
        return torch.max(feature, dim=2, keepdim=False)[0].reshape(bs, g, self.encoder_channel)

    def forward_tokens(self, point_groups, token_dtype=torch.float32):
        """(B,G,32,3) -> reduce_dim(Encoder(point_groups)) : (B,G,384), one fused pipeline.
        token_dtype=torch.float16 stores the tokens as fp16 (eval mode: in the last kernel's epilogue)."""
        if self._reduce_dim is None:
            raise RuntimeError("attach_reduce_dim() first")
        if self.training:
            if self._train_fusable(point_groups):
                out = train_forward(self, point_groups, self._state_for_pack, ops.ENC_MODES[self.precision],
                                    want_tokens=True)
            else:
                out = self._reduce_dim(self._forward_torch(point_groups))
            return out if token_dtype == torch.float32 else out.to(token_dtype)
        if self._grad_needed(point_groups):  # the fused path is forward only
            out = self._reduce_dim(self._forward_torch(point_groups))
            return out if token_dtype == torch.float32 else out.to(token_dtype)
        blob, mode = self._blob(point_groups.device)
        return ops.encoder_forward(point_groups, blob, mode=mode, token_dtype=token_dtype)

    # -- reference API ----------------------------------------------------------------------------
    def forward(self, point_groups):
        """point_groups [B,G,N,3] -> feature_global [B,G,C]."""
        if self.training:
            if self._train_fusable(point_groups):
                return train_forward(self, point_groups, self._state_for_pack, ops.ENC_MODES[self.precision],
                                     want_tokens=False)
            return self._forward_torch(point_groups)
        if self._grad_needed(point_groups) or point_groups.shape[2] != 32 or self.encoder_channel != 256:
            # a gradient is wanted (unfrozen fine-tune) or a shape the kernels are not specialised for
            # (models/pointbert/PointTransformer_8192point.yaml:17-24): the module's own torch layers
            return self._forward_torch(point_groups)
        blob, mode = self._blob(point_groups.device)
        return ops.encoder_forward(point_groups, blob, mode=mode, return_features=True, want_tokens=False)[1]
