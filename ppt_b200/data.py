"""Data-loader side of the hot path: the numpy farthest point sampling the reference runs per sample when a
stored cloud has more points than `npoints` (data/dataset_3d.py:40-61, called from the ModelNet / ScanObjectNN /
ShapeNet loaders).  Same signature, same random draw (`np.random.randint(0, N)`), same result -- on the GPU.

The reference's arithmetic on float32 input: `np.sum((xyz - centroid) ** 2, -1)` is (dx*dx + dy*dy) + dz*dz in
float32, the running `distance` array is float64 but only ever holds 1e10 (exact in float32) or float32 values,
`argmax` takes the first maximum -- exactly the FPS kernel's semantics.  float64 clouds would be
evaluated in float64 by the reference, so they are refused here rather than answered differently.
"""
import numpy as np
import torch

from . import ops


def farthest_point_sample(point, npoint, start=None, device=None):
    """point [N, D] float32 numpy (xyz in the first three columns) -> the `npoint` sampled rows [npoint, D].
    `start`: optional first index (additive keyword); by default drawn with np.random.randint(0, N) like the
    reference, so seeded loaders consume numpy's RNG identically."""
    idx = farthest_point_sample_indices(point, npoint, start, device)
    return point[idx.astype(np.int32)]


def farthest_point_sample_indices(point, npoint, start=None, device=None):
    if not isinstance(point, np.ndarray) or point.ndim != 2 or point.shape[1] < 3:
        raise ValueError("point must be a numpy array [N, D>=3]")
    if point.dtype != np.float32:
        raise TypeError("ppt_b200.data.farthest_point_sample reproduces the reference on float32 clouds only "
                        "(got %s: the reference would evaluate it in that precision)" % point.dtype)
    N = point.shape[0]
    if start is None:
        start = np.random.randint(0, N)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    xyz = torch.from_numpy(np.ascontiguousarray(point[:, :3])).to(dev, non_blocking=True).unsqueeze(0)
    st = torch.tensor([int(start)], dtype=torch.int64, device=dev)
    return ops.fps(xyz, int(npoint), st)[0].cpu().numpy()


def farthest_point_sample_batch(points, npoint, starts):
    """Many stored clouds of one size at once: points [B, N, D] float32 numpy / tensor, starts [B] -> indices
    [B, npoint] (numpy int64).  One kernel launch for the whole batch (the per-sample call above is bound by the
    launch + copy latency, not by the sampling)."""
    pts = torch.as_tensor(points)
    if pts.dtype != torch.float32 or pts.dim() != 3 or pts.shape[2] < 3:
        raise TypeError("points must be float32 [B, N, D>=3]")
    dev = pts.device if pts.is_cuda else torch.device("cuda", torch.cuda.current_device())
    xyz = pts[:, :, :3].contiguous().to(dev)
    st = torch.as_tensor(starts, dtype=torch.int64).to(dev)
    return ops.fps(xyz, int(npoint), st).cpu().numpy()
