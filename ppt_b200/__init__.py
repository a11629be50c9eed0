"""ppt_b200 -- B200-native point-cloud tokenizer behind auniquesun/PPT's Python hot path.

Host side in Python (the reference's language), compute in hand-written sm_100a
CUDA behind the C ABI of libppt_b200.so (include/ppt_b200.h).  The modules
mirror the reference's own names:

    ppt_b200.pointbert   <->  models/pointbert/{misc,dvae}.py   (fps, knn_point, Group, Encoder, ...)
    ppt_b200.pointnet2   <->  models/pointnet2/pointnet2_utils.py (query_ball_point, sample_and_group, ...)
    ppt_b200.patch       --   patch_reference(): rebinds those names on an imported reference tree
    ppt_b200.tokenizer   --   PointTokenizer: Group -> Encoder -> reduce_dim on one GPU, batch-sharded across ranks

There is no CPU path: ops raise if the CUDA library is missing or a tensor is not on a CUDA device.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
