"""PointTokenizer: the whole data-parallel front end of PPT's Point-BERT backbone on one GPU.

    xyz [B,N,3]  --FPS-->  centres  --kNN+gather+centre-->  patches [B,G,32,3]
                 --mini-PointNet (tcgen05)--> features [B,G,256] --reduce_dim--> tokens [B,G,384]

which is group_divider + encoder + reduce_dim of PointTransformer.forward
(models/pointbert/point_encoder.py:236-239).  Clouds are independent, so multi-GPU is a
contiguous batch shard per rank with no collective in the loop;
`gather_tokens` is the one validation-time all-gather.
"""
import torch
import torch.nn as nn

from . import ops
from .pointbert import Encoder, _as_start, _draw_start


class PointTokenizer(nn.Module):
    def __init__(self, num_group=512, group_size=32, encoder_dims=256, trans_dim=384, precision="fp16"):
        super().__init__()
        self.num_group, self.group_size = num_group, group_size
        self.encoder = Encoder(encoder_dims, precision=precision)
        self.reduce_dim = nn.Linear(encoder_dims, trans_dim)
        self.encoder.attach_reduce_dim(self.reduce_dim)
        self.start_idx = None

    def load_reference_state(self, sd):
        """Accepts the torch_port / reference naming: first_conv.*, second_conv.*, reduce_dim.*"""
        enc = {k: v for k, v in sd.items() if not k.startswith("reduce_dim.")}
        self.encoder.load_state_dict(enc, strict=False)
        self.reduce_dim.load_state_dict({"weight": sd["reduce_dim.weight"], "bias": sd["reduce_dim.bias"]})
        self.encoder._packed = None
        return self

    def group(self, xyz):
        start = _draw_start(xyz) if self.start_idx is None else _as_start(self.start_idx, xyz)
        _, center = ops.fps(xyz, self.num_group, start, return_centers=True)
        return ops.knn_group(xyz, center, self.group_size), center

    @torch.no_grad()
    def forward(self, xyz, return_neighborhood=False):
        """xyz [B,N,3] (CUDA) -> tokens [B,G,384], center [B,G,3] (, neighborhood [B,G,32,3])."""
        neighborhood, center = self.group(xyz)
        tokens = self.encoder.forward_tokens(neighborhood)
        if return_neighborhood:
            return tokens, center, neighborhood
        return tokens, center


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
