"""PointTokenizer: the whole data-parallel front end of PPT's Point-BERT backbone on one GPU.

    xyz [B,N,3]  --FPS-->  centres  --kNN+gather+centre-->  patches [B,G,32,3]
                 --mini-PointNet (tcgen05)--> features [B,G,256] --reduce_dim--> tokens [B,G,384]

which is group_divider + encoder + reduce_dim of PointTransformer.forward
(models/pointbert/point_encoder.py:236-239).  `forward_assembled` adds the next step (:241-247): the cls
token row, pos_embed(center) and the cls_pos row, i.e. the two tensors `self.blocks(x, pos)` consumes.  Clouds are independent, so multi-GPU is a
contiguous batch shard per rank with no collective in the loop;
`gather_tokens` is the one validation-time all-gather.
"""
import torch
import torch.nn as nn

from . import encoder_pack, ops
from .pointbert import Encoder, _as_start, _draw_start


class PointTokenizer(nn.Module):
    def __init__(self, num_group=512, group_size=32, encoder_dims=256, trans_dim=384, precision="fp16"):
        super().__init__()
        self.num_group, self.group_size = num_group, group_size
        self.encoder = Encoder(encoder_dims, precision=precision)
        self.reduce_dim = nn.Linear(encoder_dims, trans_dim)
        self.encoder.attach_reduce_dim(self.reduce_dim)
        # same names and shapes as PointTransformer's (models/pointbert/point_encoder.py:135-142)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, trans_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, trans_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, trans_dim))
        self.start_idx = None
        self._pos_packed = None
        self._pos_key = None

    def load_reference_state(self, sd):
        """Accepts the torch_port / reference naming: first_conv.*, second_conv.*, reduce_dim.*"""
        enc = {k: v for k, v in sd.items() if not k.startswith("reduce_dim.")}
        self.encoder.load_state_dict(enc, strict=False)
        self.reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
        self.encoder._packed = None
        return self

    def load_front_end_state(self, module):
        """Copies cls_token, cls_pos and pos_embed from a reference PointTransformer (or its state dict)."""
        sd = module if isinstance(module, dict) else module.state_dict()
        with torch.no_grad():
            self.cls_token.copy_(sd["cls_token"])
            self.cls_pos.copy_(sd["cls_pos"])
        self.pos_embed.load_state_dict({k[len("pos_embed."):]: v for k, v in sd.items() if k.startswith("pos_embed.")})
        self._pos_packed = None
        return self

    def _pos_blob(self, device):
        mode = ops.ENC_MODES[self.encoder.precision]
        tensors = [self.cls_token, self.cls_pos] + list(self.pos_embed.parameters())
        key = (mode, str(device)) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._pos_packed is None or self._pos_key != key:
            self._pos_packed = encoder_pack.pack_pos_embed(self.pos_embed.state_dict(), self.cls_token, self.cls_pos,
                                                           mode).to(device)
            self._pos_key = key
        return self._pos_packed

    @torch.no_grad()
    def forward_assembled(self, xyz):
        """xyz [B,N,3] (CUDA) -> x [B,G+1,384], pos [B,G+1,384], center [B,G,3]: the arguments of
        `self.blocks(x, pos)` in PointTransformer.forward (models/pointbert/point_encoder.py:241-249)."""
        neighborhood, center = self.group(xyz)
        blob, mode = self.encoder._blob(xyz.device)
        x, pos = ops.tokenizer_forward(neighborhood, center, blob, self._pos_blob(xyz.device), mode=mode)
        return x, pos, center

    def group(self, xyz):
        start = _draw_start(xyz) if self.start_idx is None else _as_start(self.start_idx, xyz)
        index = ops.spatial_index(xyz)  # one index serves both FPS and kNN
        _, center = ops.fps(xyz, self.num_group, start, return_centers=True, index=index)
        return ops.knn_group(xyz, center, self.group_size, index=index), center

    @torch.no_grad()
    def forward(self, xyz, return_neighborhood=False, token_dtype=torch.float32):
        """xyz [B,N,3] (CUDA) -> tokens [B,G,384], center [B,G,3] (, neighborhood [B,G,32,3]).
        token_dtype: torch.float32 (the reference's dtype) or torch.float16 (half the bytes; adds one fp16 rounding)."""
        neighborhood, center = self.group(xyz)
        tokens = self.encoder.forward_tokens(neighborhood, token_dtype=token_dtype)
        if return_neighborhood:
            return tokens, center, neighborhood
        return tokens, center


class HostPipeline:
    """Host-to-host tokenization of a stream of batches: pinned host clouds in, pinned host tokens out.

    Three CUDA streams (H2D copy, kernels, D2H copy) and `depth` rotating slots, so the copy of
    batch i+1 and the read-back of batch i-1 overlap the kernels of batch i.  This is the path a
    data-loader-fed caller uses (pc.to(gpu) ... features.cpu(), main_cls.py:188-189,
    lp_feat_extractor.py:53-56), and what bench.py reports as `e2e`."""

    def __init__(self, tokenizer, batch, points, depth=2, device=None, token_dtype=torch.float32, numa_local=True):
        """token_dtype: dtype of the tokens that reach the host (fp32 = the reference's; fp16 halves the D2H bytes,
        which bound this path on PCIe).  numa_local: allocate the pinned output buffers on the NUMA node the GPU
        hangs off (hostmem.pinned_empty), so the DMA does not cross the socket interconnect."""
        from . import hostmem
        self.tok = tokenizer
        self.device = device or next(tokenizer.parameters()).device
        self.depth = depth
        self.token_dtype = token_dtype
        G = tokenizer.num_group
        D = tokenizer.reduce_dim.out_features
        alloc = (lambda shape, dt: hostmem.pinned_empty(shape, dt, self.device)) if numa_local else \
            (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory())
        self.dev_in = [torch.empty((batch, points, 3), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.out_tokens = [alloc((batch, G, D), token_dtype) for _ in range(depth)]
        self.out_center = [alloc((batch, G, 3), torch.float32) for _ in range(depth)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_consumed = [torch.cuda.Event() for _ in range(depth)]  # kernels done reading dev_in[slot]

    def run(self, host_batches, on_result=None):
        """host_batches: iterable of pinned [B,N,3] fp32 tensors.  `on_result(i, tokens, center)` is
        called with the pinned output buffers of batch i once they are complete (they are reused
        `depth` batches later).  Returns the number of batches processed."""
        pending = []
        n = 0
        for i, h in enumerate(host_batches):
            slot = i % self.depth
            if i >= self.depth:  # slot reuse: its previous outputs must have reached the host
                self.ev_out[slot].synchronize()
                if on_result is not None:
                    j = pending.pop(0)
                    on_result(j, self.out_tokens[slot], self.out_center[slot])
            with torch.cuda.stream(self.s_in):
                if i >= self.depth:
                    self.s_in.wait_event(self.ev_consumed[slot])
                self.dev_in[slot].copy_(h, non_blocking=True)
                self.ev_in[slot].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[slot])
                tokens, center = self.tok(self.dev_in[slot], token_dtype=self.token_dtype)
                self.ev_consumed[slot].record(self.s_run)
                self.ev_run[slot].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[slot])
                tokens.record_stream(self.s_out)
                center.record_stream(self.s_out)
                self.out_tokens[slot].copy_(tokens, non_blocking=True)
                self.out_center[slot].copy_(center, non_blocking=True)
                self.ev_out[slot].record(self.s_out)
            pending.append(i)
            n += 1
        for j in pending:
            slot = j % self.depth
            self.ev_out[slot].synchronize()
            if on_result is not None:
                on_result(j, self.out_tokens[slot], self.out_center[slot])
        return n


def shard_bounds(total, rank, world_size):
    """Contiguous [lo, hi) slice of `total` clouds for `rank` (mirrors DistributedSampler's even split,
    main_cls.py:74-76, but contiguous and without padding)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tokens(tokens, group=None):
    """Validation-time all-gather of per-rank token shards (equal shard sizes) -> [world*B_local, G, D]."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    tokens = tokens.contiguous()
    if dist.get_backend(group) == "nccl":
        out = tokens.new_empty((world * tokens.shape[0],) + tuple(tokens.shape[1:]))
        dist.all_gather_into_tensor(out, tokens, group=group)  # one NCCL all-gather over NVLink
        return out
    parts = [torch.empty_like(tokens) for _ in range(world)]  # gloo (CPU tests)
    dist.all_gather(parts, tokens, group=group)
    return torch.cat(parts, dim=0)
