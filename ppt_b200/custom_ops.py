"""torch.library registration of the kernels: `torch.ops.ppt_b200.*`.

The reference-facing functions (ppt_b200.pointbert / pointnet2) call the C ABI directly through ctypes, which is
the cheapest path in eager mode but opaque to a tracer.  These custom ops wrap the same calls with fake-tensor
(shape / dtype) implementations, so `torch.compile` over a patched model sees opaque, correctly-shaped ops instead
of failing on `data_ptr()`; the wrappers switch to them automatically while compiling (`use_custom_ops()`).
Index-valued ops carry no gradient; three_interpolate and gather register theirs.
"""
import torch

from . import ops

_lib_def = torch.library.custom_op


def use_custom_ops():
    try:
        return torch.compiler.is_compiling()
    except Exception:
        return False


@_lib_def("ppt_b200::fps", mutates_args=())
def fps(xyz: torch.Tensor, npoint: int, start: torch.Tensor) -> torch.Tensor:
    return ops.fps(xyz, npoint, start)


@fps.register_fake
def _(xyz, npoint, start):
    return xyz.new_empty((xyz.shape[0], npoint), dtype=torch.int64)


@_lib_def("ppt_b200::fps_centers", mutates_args=())
def fps_centers(xyz: torch.Tensor, npoint: int, start: torch.Tensor) -> torch.Tensor:
    return ops.fps(xyz, npoint, start, return_centers=True)[1]


@fps_centers.register_fake
def _(xyz, npoint, start):
    return xyz.new_empty((xyz.shape[0], npoint, 3), dtype=torch.float32)


@_lib_def("ppt_b200::square_distance", mutates_args=())
def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    return ops.square_distance(src, dst)


@square_distance.register_fake
def _(src, dst):
    return src.new_empty((src.shape[0], src.shape[1], dst.shape[1]), dtype=torch.float32)


@_lib_def("ppt_b200::knn", mutates_args=())
def knn(k: int, xyz: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
    return ops.knn(k, xyz, query)


@knn.register_fake
def _(k, xyz, query):
    return xyz.new_empty((xyz.shape[0], query.shape[1], k), dtype=torch.int64)


@_lib_def("ppt_b200::group", mutates_args=())
def group(xyz: torch.Tensor, num_group: int, group_size: int, start: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    index = ops.spatial_index(xyz)
    _, center = ops.fps(xyz, num_group, start, return_centers=True, index=index)
    return ops.knn_group(xyz, center, group_size, index=index), center


@group.register_fake
def _(xyz, num_group, group_size, start):
    B = xyz.shape[0]
    return (xyz.new_empty((B, num_group, group_size, 3), dtype=torch.float32),
            xyz.new_empty((B, num_group, 3), dtype=torch.float32))


@_lib_def("ppt_b200::ball_query", mutates_args=())
def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    return ops.ball_query(radius, nsample, xyz, new_xyz)


@ball_query.register_fake
def _(radius, nsample, xyz, new_xyz):
    return xyz.new_empty((xyz.shape[0], new_xyz.shape[1], nsample), dtype=torch.int64)


@_lib_def("ppt_b200::gather", mutates_args=())
def gather(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return ops._gather_fwd(ops._f32(points), ops._i64(idx))


@gather.register_fake
def _(points, idx):
    return points.new_empty(tuple(idx.shape) + (points.shape[2],), dtype=torch.float32)


def _gather_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.N = inputs[0].shape[1]


def _gather_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    B, C = grad.shape[0], grad.shape[-1]
    return ops._scatter_rows(grad.contiguous().reshape(B, -1, C), idx.reshape(B, -1), ctx.N), None


gather.register_autograd(_gather_backward, setup_context=_gather_setup)


@_lib_def("ppt_b200::three_nn", mutates_args=())
def three_nn(unknown: torch.Tensor, known: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    return ops.three_nn(unknown, known)


@three_nn.register_fake
def _(unknown, known):
    B, N = unknown.shape[0], unknown.shape[1]
    return unknown.new_empty((B, N, 3), dtype=torch.float32), unknown.new_empty((B, N, 3), dtype=torch.int64)


@_lib_def("ppt_b200::three_interpolate", mutates_args=())
def three_interpolate(feats: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor) -> torch.Tensor:
    return ops._interp_fwd(ops._f32(feats), ops._i64(idx), ops._f32(dist))


@three_interpolate.register_fake
def _(feats, idx, dist):
    return feats.new_empty((feats.shape[0], idx.shape[1], feats.shape[2]), dtype=torch.float32)


@_lib_def("ppt_b200::three_interpolate_grad", mutates_args=())
def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor, S: int) -> torch.Tensor:
    from . import _lib
    g = ops._f32(grad_out)
    B, N, D = g.shape
    out = torch.zeros((B, S, D), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().ppt_three_interpolate_grad(g.data_ptr(), idx.data_ptr(), dist.data_ptr(), out.data_ptr(),
                                                          B, N, S, D, ops._stream(g)), "ppt_three_interpolate_grad")
    return out


@three_interpolate_grad.register_fake
def _(grad_out, idx, dist, S):
    return grad_out.new_empty((grad_out.shape[0], S, grad_out.shape[2]), dtype=torch.float32)


def _interp_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1], inputs[2])
    ctx.S = inputs[0].shape[1]


def _interp_backward(ctx, grad):
    idx, dist = ctx.saved_tensors
    return three_interpolate_grad(grad, idx, dist, ctx.S), None, None


three_interpolate.register_autograd(_interp_backward, setup_context=_interp_setup)


@_lib_def("ppt_b200::encoder_tokens", mutates_args=())
def encoder_tokens(neighborhood: torch.Tensor, packed: torch.Tensor, mode: int) -> torch.Tensor:
    return ops.encoder_forward(neighborhood, packed, mode=mode)


@encoder_tokens.register_fake
def _(neighborhood, packed, mode):
    return neighborhood.new_empty(tuple(neighborhood.shape[:-2]) + (384,), dtype=torch.float32)


@_lib_def("ppt_b200::encoder_features", mutates_args=())
def encoder_features(neighborhood: torch.Tensor, packed: torch.Tensor, mode: int) -> torch.Tensor:
    return ops.encoder_forward(neighborhood, packed, mode=mode, return_features=True, want_tokens=False)[1]


@encoder_features.register_fake
def _(neighborhood, packed, mode):
    return neighborhood.new_empty(tuple(neighborhood.shape[:-2]) + (256,), dtype=torch.float32)
