"""patch_reference(): install the B200 kernels behind an imported auniquesun/PPT tree.

The reference has no operator registry; its boundary is Python names.
Module-level functions are rebound on the reference modules (intra-module calls resolve
through module globals at call time, so `sample_and_group` picks up the new
`farthest_point_sample`), and the nn.Module classes get their `forward` replaced in place,
which is import-order-proof (`from models.pointbert.dvae import Group` elsewhere keeps working).

CPU tensors keep running the reference's own code: the wrappers only divert CUDA tensors.
"""
import functools
import importlib
import sys

import torch

from . import encoder_pack, ops, pointbert, pointnet2

_FUNCTIONS = {
    # reference module -> {attribute: replacement}
    "models.pointbert.misc": {
        "fps": pointbert.fps, "farthest_point_sample": pointbert.farthest_point_sample,
        "index_points": pointbert.index_points,
    },
    "models.pointbert.dvae": {"knn_point": pointbert.knn_point, "square_distance": pointbert.square_distance},
    "models.pointbert.pointnet2_utils": {
        "farthest_point_sample": pointnet2.farthest_point_sample, "index_points": pointbert.index_points,
        "knn_point": pointbert.knn_point, "square_distance": pointbert.square_distance,
        "query_ball_point": pointnet2.query_ball_point, "sample_and_group": pointnet2.sample_and_group,
    },
    "models.pointnet2.pointnet2_utils": {
        "farthest_point_sample": pointnet2.farthest_point_sample, "index_points": pointbert.index_points,
        "square_distance": pointbert.square_distance, "query_ball_point": pointnet2.query_ball_point,
        "sample_and_group": pointnet2.sample_and_group,
    },
    "models.pointbert.point_encoder": {},  # PointTransformer.forward (class patch below)
    "models.pointmlp.pointMLP": {
        "furthest_point_sample": pointbert.farthest_point_sample,  # sic: pointMLP.py:64 spells it "furthest"
        "index_points": pointbert.index_points,
        "knn_point": pointbert.knn_point, "square_distance": pointbert.square_distance,
        "query_ball_point": pointnet2.query_ball_point,
    },
}

_installed = []  # (owner, attribute, original)
missing = []     # 'module.attribute' entries of the table that the imported tree did not have (last patch_reference call)


# ---- the envelope the kernels cover; anything else keeps the reference's own code ----------------------------
def _f32_cuda(*tensors):
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors)


def _is_xyz(*tensors):
    return all(t.dim() == 3 and t.shape[-1] == 3 for t in tensors)


def _ok_fps(xyz, npoint, *a, **k):
    return _f32_cuda(xyz) and _is_xyz(xyz) and 1 <= xyz.shape[1] <= 65536 and npoint >= 1


def _ok_fps_data(data, number, *a, **k):
    return _ok_fps(data, number)


def _ok_index_points(points, idx, *a, **k):
    return (_f32_cuda(points) and points.dim() == 3 and idx.dtype == torch.int64 and idx.dim() in (2, 3)
            and points.shape[0] <= 65535)


def _ok_sqdist(src, dst, *a, **k):
    return _f32_cuda(src, dst) and _is_xyz(src, dst) and src.shape[1] <= 65535 and src.shape[0] <= 65535


def _ok_knn(nsample, xyz, new_xyz, *a, **k):
    return _f32_cuda(xyz, new_xyz) and _is_xyz(xyz, new_xyz) and 1 <= nsample <= min(32, xyz.shape[1])


def _ok_ball(radius, nsample, xyz, new_xyz, *a, **k):
    return _f32_cuda(xyz, new_xyz) and _is_xyz(xyz, new_xyz) and nsample >= 1


def _ok_sample_and_group(npoint, radius, nsample, xyz, points, *a, **k):
    return (_ok_fps(xyz, npoint) and nsample >= 1 and xyz.shape[0] <= 65535
            and (points is None or (_f32_cuda(points) and points.dim() == 3)))


_SUPPORTED = {"fps": _ok_fps_data, "farthest_point_sample": _ok_fps, "furthest_point_sample": _ok_fps,
              "index_points": _ok_index_points, "square_distance": _ok_sqdist, "knn_point": _ok_knn,
              "query_ball_point": _ok_ball, "sample_and_group": _ok_sample_and_group}


def _first_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            return a
    return None


def _divert_cuda(original, replacement, supported=None):
    """CUDA fp32 arguments inside the kernels' envelope (3-D coordinates, k <= 32, N <= 65536, ...) go to
    `replacement`; CPU tensors, other dtypes, feature-space inputs and out-of-envelope sizes keep running the
    reference's own code -- never an error or a wrong answer where the reference had a right one."""
    @functools.wraps(original)
    def wrapper(*args, **kwargs):
        t = _first_tensor(args, kwargs)
        if t is not None and t.is_cuda:
            try:
                ok = supported is None or supported(*args, **kwargs)
            except Exception:
                ok = False
            if ok:
                return replacement(*args, **kwargs)
        return original(*args, **kwargs)

    wrapper.__ppt_b200_original__ = original
    return wrapper


def _set(owner, name, value):
    _installed.append((owner, name, getattr(owner, name)))
    setattr(owner, name, value)


def _group_forward(self, xyz):
    """Group.forward, models/pointbert/dvae.py:159-181."""
    return pointbert.group_forward(xyz, self.num_group, self.group_size, pointbert._draw_start(xyz))


def _encoder_forward(self, point_groups):
    """Encoder.forward (eval), models/pointbert/dvae.py:201-215, on the reference's own module instance."""
    mode = ops.ENC_MODES[getattr(self, "ppt_precision", "fp16")]
    tensors = list(self.parameters()) + list(self.buffers())
    key = (mode, str(point_groups.device), getattr(self, "_bn_epoch", 0)) + \
        tuple((t.data_ptr(), t._version) for t in tensors)
    if getattr(self, "_ppt_key", None) != key:
        sd = _encoder_state_zero_reduce(self)
        object.__setattr__(self, "_ppt_blob", encoder_pack.pack_encoder(sd, mode).to(point_groups.device))
        object.__setattr__(self, "_ppt_key", key)
    return ops.encoder_forward(point_groups, self._ppt_blob, mode=mode, return_features=True, want_tokens=False)[1]


def point_transformer_front_end(model, pts):
    """Lines 236-247 of PointTransformer.forward (models/pointbert/point_encoder.py) on the reference's own module
    instance: returns the (x, pos) that `model.blocks(x, pos)` consumes.  Grouping, Encoder, reduce_dim, the
    cls rows and pos_embed run as one kernel pipeline (ops.tokenizer_forward); the packed weights are rebuilt
    when any of the parameters involved changes."""
    neighborhood, center = model.group_divider(pts)
    mode = ops.ENC_MODES[getattr(model, "ppt_precision", "fp16")]
    enc_t = list(model.encoder.parameters()) + list(model.encoder.buffers()) + list(model.reduce_dim.parameters())
    pos_t = [model.cls_token, model.cls_pos] + list(model.pos_embed.parameters())
    key = (mode, str(pts.device), getattr(model.encoder, "_bn_epoch", 0)) + \
        tuple((t.data_ptr(), t._version) for t in enc_t + pos_t)
    if getattr(model, "_ppt_front_key", None) != key:
        sd = dict(model.encoder.state_dict())
        sd["reduce_dim.weight"], sd["reduce_dim.bias"] = model.reduce_dim.weight, model.reduce_dim.bias
        blobs = (encoder_pack.pack_encoder(sd, mode).to(pts.device),
                 encoder_pack.pack_pos_embed(model.pos_embed.state_dict(), model.cls_token, model.cls_pos, mode)
                 .to(pts.device))
        object.__setattr__(model, "_ppt_front_blobs", blobs)
        object.__setattr__(model, "_ppt_front_key", key)
    enc_blob, pos_blob = model._ppt_front_blobs
    return ops.tokenizer_forward(neighborhood, center, enc_blob, pos_blob, mode=mode)


def _front_end_fusable(model, pts):
    enc = model.encoder
    # forward only: fine under no_grad, or when everything involved is frozen (PPT: models/ULIP_models.py:505)
    involved = list(enc.parameters()) + list(model.reduce_dim.parameters()) + list(model.pos_embed.parameters()) + \
        [model.cls_token, model.cls_pos]
    frozen = not torch.is_grad_enabled() or not (pts.requires_grad or any(p.requires_grad for p in involved))
    return (pts.is_cuda and not enc.training and frozen and model.group_size == 32
            and model.num_group >= 32 and getattr(enc, "encoder_channel", 0) == 256
            and tuple(model.reduce_dim.weight.shape) == (384, 256) and tuple(model.pos_embed[0].weight.shape) == (128, 3))


def _point_transformer_forward(self, pts, _original):
    """PointTransformer.forward, models/pointbert/point_encoder.py:234-256: fused front end, then the
    reference's own transformer blocks, norm and [cls, max] readout."""
    if not _front_end_fusable(self, pts):
        return _original(self, pts)
    x, pos = point_transformer_front_end(self, pts)
    x = self.norm(self.blocks(x, pos, task="cls"))
    return torch.cat([x[:, 0], x[:, 1:].max(1)[0]], dim=-1)


def _encoder_state_zero_reduce(self):
    sd = dict(self.state_dict())
    ref = sd["first_conv.0.weight"]
    sd["reduce_dim.weight"], sd["reduce_dim.bias"] = ref.new_zeros((384, 256)), ref.new_zeros((384,))
    return sd


def _encoder_forward_train(self, point_groups):
    """Encoder.forward under model.train() (main_cls.py:169): batch-statistics BatchNorm, fused, forward only."""
    mode = ops.ENC_MODES[getattr(self, "ppt_precision", "fp16")]
    return pointbert.train_forward(self, point_groups, lambda: _encoder_state_zero_reduce(self), mode, want_tokens=False)


def _fp_forward(self, xyz1, xyz2, points1, points2):
    return pointnet2.PointNetFeaturePropagation.forward(self, xyz1, xyz2, points1, points2)


def patch_reference(modules=None):
    """Rebinds the hot-path names on every reference module that is importable.
    Returns the list of patched 'module.attribute' names; entries of the table that the tree does not have are
    listed in `patch.missing` and reported with a warning (a silently unpatched function would leave the Python
    loop in place).  Idempotent."""
    if _installed:
        return [getattr(o, "__name__", repr(o)) + "." + n for o, n, _ in _installed]
    del missing[:]
    names = modules if modules is not None else list(_FUNCTIONS)
    for modname in names:
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue  # optional backbone whose imports are unavailable
        for attr, repl in _FUNCTIONS.get(modname, {}).items():
            if hasattr(mod, attr):
                _set(mod, attr, _divert_cuda(getattr(mod, attr), repl, _SUPPORTED.get(attr)))
            else:
                missing.append(modname + "." + attr)
        if modname == "models.pointbert.dvae":
            orig_g, orig_e = mod.Group.forward, mod.Encoder.forward

            def group_fwd(self, xyz, _o=orig_g):
                ok = _f32_cuda(xyz) and _is_xyz(xyz) and xyz.shape[1] <= 65536 and \
                    1 <= self.group_size <= min(32, xyz.shape[1])
                return _group_forward(self, xyz) if ok else _o(self, xyz)

            def enc_fwd(self, pg, _o=orig_e):
                if pg.is_cuda and self.training and pointbert.train_forward_fusable(self, pg):
                    return _encoder_forward_train(self, pg)
                # forward-only kernels: an unfrozen fine-tune (anything here wants a gradient) keeps the torch layers
                frozen = not pointbert._needs_grad(pg, *self.parameters())
                fused = (_f32_cuda(pg) and not self.training and frozen and pg.dim() == 4 and pg.shape[2] == 32
                         and pg.shape[3] == 3 and self.encoder_channel == 256)
                return _encoder_forward(self, pg) if fused else _o(self, pg)

            _set(mod.Group, "forward", group_fwd)
            _set(mod.Encoder, "forward", enc_fwd)
        if modname == "models.pointbert.point_encoder" and hasattr(mod, "PointTransformer"):
            orig_pt = mod.PointTransformer.forward

            def pt_fwd(self, pts, _o=orig_pt):
                return _point_transformer_forward(self, pts, _o)

            _set(mod.PointTransformer, "forward", pt_fwd)
        if modname in ("models.pointnet2.pointnet2_utils", "models.pointbert.pointnet2_utils"):
            cls = getattr(mod, "PointNetFeaturePropagation", None)
            if cls is not None:
                orig_f = cls.forward

                def fp_fwd(self, xyz1, xyz2, points1, points2, _o=orig_f):
                    if _f32_cuda(xyz1, xyz2, points2) and xyz1.shape[1] == 3 and (xyz2.shape[2] == 1 or xyz2.shape[2] >= 3):
                        return _fp_forward(self, xyz1, xyz2, points1, points2)
                    return _o(self, xyz1, xyz2, points1, points2)

                _set(cls, "forward", fp_fwd)
            dg = getattr(mod, "DGCNN_Propagation", None)
            if dg is not None:
                orig_df = dg.forward

                def dgcnn_fwd(self, coor, f, coor_q, f_q, _o=orig_df):
                    if pointnet2.dgcnn_fusable(self, coor, f, coor_q, f_q):
                        return pointnet2.dgcnn_propagation_forward(self, coor, f, coor_q, f_q)
                    return _o(self, coor, f, coor_q, f_q)

                _set(dg, "forward", dgcnn_fwd)
                orig_gf = dg.get_graph_feature

                def graph_fwd(self, coor_q, x_q, coor_k, x_k, _o=orig_gf):
                    if x_q.is_cuda:
                        return pointnet2.get_graph_feature(coor_q, x_q, coor_k, x_k, self.k)
                    return _o(self, coor_q, x_q, coor_k, x_k)

                _set(dg, "get_graph_feature", graph_fwd)
            ssg = getattr(mod, "PointNetSetAbstraction", None)
            if ssg is not None:
                orig_s = ssg.forward

                def ssg_fwd(self, xyz, points, _o=orig_s):
                    if _f32_cuda(xyz) and xyz.shape[1] == 3 and xyz.shape[0] <= 65535:
                        if not hasattr(self, "start_idx"):
                            self.start_idx = None
                        return pointnet2.PointNetSetAbstraction.forward(self, xyz, points)  # fused shared MLP in eval
                    return _o(self, xyz, points)

                _set(ssg, "forward", ssg_fwd)
            msg = getattr(mod, "PointNetSetAbstractionMsg", None)
            if msg is not None:
                orig_m = msg.forward

                def msg_fwd(self, xyz, points, _o=orig_m):
                    if _f32_cuda(xyz) and xyz.shape[1] == 3 and xyz.shape[0] <= 65535:
                        if not hasattr(self, "start_idx"):
                            self.start_idx = None
                        return pointnet2.PointNetSetAbstractionMsg.forward(self, xyz, points)
                    return _o(self, xyz, points)

                _set(msg, "forward", msg_fwd)
    if missing:
        import warnings
        warnings.warn("ppt_b200.patch: not found in the imported tree (left unpatched): " + ", ".join(missing))
    return [getattr(o, "__name__", repr(o)) + "." + n for o, n, _ in _installed]


def invalidate(model):
    """Drops every packed-weight cache hanging off `model`'s sub-modules.  The caches are keyed on (data_ptr,
    _version), which misses writes through `param.data.copy_()` -- the reference's own loading idiom
    (models/ULIP_models.py:507): call this after loading weights that way once a forward has already run."""
    for m in model.modules():
        for name in ("_ppt_key", "_ppt_blob", "_ppt_train_key", "_ppt_train_blob", "_ppt_front_key", "_ppt_front_blobs",
                     "_ppt_sa_cache", "_packed", "_packed_key", "_pos_packed", "_pos_key", "_ppt_fp_packed",
                     "_ppt_fp_folded"):
            if name in m.__dict__:
                object.__setattr__(m, name, {} if name == "_ppt_sa_cache" else None)


def unpatch_reference():
    while _installed:
        owner, name, original = _installed.pop()
        setattr(owner, name, original)
