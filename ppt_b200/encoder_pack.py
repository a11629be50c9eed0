"""Host-side preparation of the patch Encoder's weights for the tcgen05 kernels.

Done once per weight version (the Encoder is frozen in PPT, models/ULIP_models.py:505):

1. BatchNorm folding (eval mode, eps 1e-5):  W' = W * g/sqrt(v+eps),  b' = (b - m) * g/sqrt(v+eps) + beta
   for first_conv.{0,1} and second_conv.{0,1}  (models/pointbert/dvae.py:188-199).
2. second_conv.0 consumes cat([global.expand, feature]) (dvae.py:212): its weight splits into
   W3a (columns 0..255, applied to the per-group max) and W3b (columns 256..511, per point).
   first_conv.3 (128->256, no activation after it) and W3b are both linear, so they compose:
       W32 = W3b' @ W2   (512 x 128, computed in fp64)
   which halves the per-point work of that layer.  The per-group term becomes
       c = W3a' @ max_pts(W2 @ h1) + [ (W3a' + W3b') @ b2 + b3' ].
3. Biases of max-pooled layers move behind the max; reduce_dim's input bias folds into
   its own: tokens = Wr @ max_pts(W4 @ h3) + (Wr @ b4 + br)   (models/pointbert/point_encoder.py:133,239).
4. Every tensor-core weight is cut into 128-row x 64-column operand images in the K-major,
   128-byte-swizzled layout tcgen05.mma reads from shared memory (csrc/tc05.cuh: sw128_kmajor_off),
   so a pipeline stage is one contiguous 16 KB bulk copy.  Order: [unit][k-chunk][split][16 KB].
5. fp32-parity mode (ENC_FP16X3): every operand is split into fp16 hi + fp16 lo and each product is
   three MMAs (hi*hi, hi*lo, lo*hi).  fp16 carries 11 significant bits, so hi+lo carries ~22 -- but only
   inside fp16's narrow exponent range -- so weights and activations are pre-multiplied by powers of two
   (exact) that centre them in that range, and the accumulators are multiplied back in the epilogue.
   bf16 hi/lo (16 bits) measured 1.2e-5 on B200, fp16 hi/lo with scaling reaches the fp32 reference's
   own rounding noise.

Blob layout (bytes), mirrored by csrc/encoder.cu (EncoderBlob):
    [0, 8192)            fp32 section: W1' rows {w0,w1,w2,b} [128][4], bias_c [512], b4 [256], bias_tok [384],
                         scales [8] = 1/(weight scale * operand scale) of the five GEMMs, point-activation
                         scale, group-operand scale, 0   (all 1.0 outside the fp32-parity mode)
    then W2 (2 units x 2 chunks), W3A (4 x 4), W32 (4 x 2), W4 (2 x 8), WR (3 x 4) operand images,
    then W1T: the layer-1 image of layer1_image() (one 16 KB image in every mode),
    then W32F: W32 as plain fp32 [512][128] (train mode: second-moment statistics from the Gram matrix of h1).
"""
import torch

ENC_FP16, ENC_BF16, ENC_FP16X3 = 0, 1, 2
ACT_SCALE = 64.0   # fp32-parity mode: h1/h3 operands (values up to 1023 before fp16 saturates)
GRP_SCALE = 64.0   # fp32-parity mode: g/t operands
F32_SECTION_BYTES = 8192
IMAGE_BYTES = 16384
W32F_BYTES = 512 * 128 * 4
# (name, rows, cols)
SECTIONS = (("W2", 256, 128), ("W3A", 512, 256), ("W32", 512, 128), ("W4", 256, 512), ("WR", 384, 256))


def operand_dtype(mode):
    return torch.bfloat16 if mode == ENC_BF16 else torch.float16


def split_of(mode):
    return 2 if mode == ENC_FP16X3 else 1


def weight_scale(w):
    """Power of two that puts max|w| in [4096, 8192): hi keeps 11 bits, lo stays a normal fp16."""
    m = float(w.abs().max())
    if m == 0.0 or not torch.isfinite(torch.tensor(m)):
        return 1.0
    import math
    return 2.0 ** (12 - math.floor(math.log2(m)) )


def packed_bytes(mode):
    n = sum((r // 128) * (c // 64) for _, r, c in SECTIONS)
    # + the layer-1 image (W1T) + W32 in fp32 [512][128] (W32F: the train-mode statistics use the exact weights)
    return F32_SECTION_BYTES + n * split_of(mode) * IMAGE_BYTES + IMAGE_BYTES + W32F_BYTES


def layer1_image(w1, dtype):
    """Layer 1 (K = 3 plus bias) as ONE tensor-core K=16 slice.  Weights and coordinates are both
    split into hi + lo operand parts laid out along K, so a single MMA evaluates
        W_hi.x_hi + W_hi.x_lo + W_lo.x_hi + b_hi + b_lo      (fp32 accumulate)
    i.e. the fp32 product up to the dropped W_lo.x_lo term.  Row = channel, K slots:
        0-2 W_hi | 3-5 W_hi | 6-8 W_lo | 9 b_hi | 10 b_lo | 11-63 zero
    (the kernel writes the matching point rows: x_hi | x_lo | x_hi | 1 | 1 | 0)."""
    w1 = w1.to(torch.float32)
    hi = w1.to(dtype)
    lo = (w1 - hi.to(torch.float32)).to(dtype)
    k = torch.zeros(128, 64, dtype=torch.float32)
    k[:, 0:3] = hi[:, :3].float()
    k[:, 3:6] = hi[:, :3].float()
    k[:, 6:9] = lo[:, :3].float()
    k[:, 9] = hi[:, 3].float()
    k[:, 10] = lo[:, 3].float()
    return pack_kmajor(k, dtype, 1)  # every entry is exactly representable: no second rounding


def pack_kmajor(w, dtype, split=1):
    """w [R, K] fp32 (R % 128 == 0, K % 64 == 0) -> uint8 tensor of operand images
    [R/128][K/64][split][128 rows x 128 B], 16-byte pieces XOR-swizzled by (row & 7)."""
    R, K = w.shape
    assert R % 128 == 0 and K % 64 == 0
    w = w.detach().to(torch.float32).cpu()
    parts = [w.to(dtype)]
    if split == 2:
        parts.append((w - parts[0].to(torch.float32)).to(dtype))
    elif dtype == torch.float16:
        parts[0] = w.clamp(-65504.0, 65504.0).to(dtype)
    r = torch.arange(128).view(128, 1)
    q = torch.arange(8).view(1, 8)
    src_piece = (q ^ (r & 7)).view(1, 1, 128, 8, 1).expand(R // 128, K // 64, 128, 8, 8)
    imgs = []
    for p in parts:
        t = p.view(torch.int16).view(R // 128, 128, K // 64, 8, 8).permute(0, 2, 1, 3, 4)  # [u][kc][row][piece][8]
        imgs.append(torch.gather(t, 3, src_piece))  # physical piece q holds logical piece q ^ (row & 7)
    out = torch.stack(imgs, dim=2).contiguous()  # [u][kc][split][128][8][8] int16
    return out.view(torch.uint8).reshape(-1)


def fold(sd, eps=1e-5):
    """State dict (reference naming, + reduce_dim.{weight,bias}) -> folded fp64 matrices."""
    d = {k: v.detach().to(torch.float64).cpu() for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    s1 = d["first_conv.1.weight"] / torch.sqrt(d["first_conv.1.running_var"] + eps)
    w1 = d["first_conv.0.weight"].reshape(128, 3) * s1[:, None]
    b1 = (d["first_conv.0.bias"] - d["first_conv.1.running_mean"]) * s1 + d["first_conv.1.bias"]
    w2 = d["first_conv.3.weight"].reshape(256, 128)
    b2 = d["first_conv.3.bias"]
    s3 = d["second_conv.1.weight"] / torch.sqrt(d["second_conv.1.running_var"] + eps)
    w3 = d["second_conv.0.weight"].reshape(512, 512) * s3[:, None]
    b3 = (d["second_conv.0.bias"] - d["second_conv.1.running_mean"]) * s3 + d["second_conv.1.bias"]
    w3a, w3b = w3[:, :256], w3[:, 256:]
    w4 = d["second_conv.3.weight"].reshape(-1, 512)
    b4 = d["second_conv.3.bias"]
    wr, br = d["reduce_dim.weight"], d["reduce_dim.bias"]
    if w4.shape[0] != 256 or tuple(wr.shape) != (384, 256):
        raise ValueError("kernels are specialised for encoder_dims=256, trans_dim=384 "
                         "(models/pointbert/PointTransformer_8192point.yaml:17-24)")
    return {
        "W1": torch.cat([w1, b1[:, None]], dim=1),  # [128,4]
        "W2": w2, "W3A": w3a, "W32": w3b @ w2, "W4": w4, "WR": wr,
        "bias_c": (w3a + w3b) @ b2 + b3, "b4": b4, "bias_tok": wr @ b4 + br,
    }


def pack_encoder(sd, mode):
    """-> uint8 CPU tensor of packed_bytes(mode) bytes."""
    f = fold(sd)
    dtype, split = operand_dtype(mode), split_of(mode)
    if mode == ENC_FP16X3:
        ws = {name: weight_scale(f[name]) for name, _, _ in SECTIONS}
        act, grp = ACT_SCALE, GRP_SCALE
    else:
        ws = {name: 1.0 for name, _, _ in SECTIONS}
        act = grp = 1.0
    # inverse accumulator scales in launch order: stage1 (W2 x h1), linear c (W3A x g), stage2 (W32 x h1),
    # stage2 (W4 x h3), linear tokens (WR x t); then the two operand scales
    scales = torch.tensor([1.0 / (ws["W2"] * act), 1.0 / (ws["W3A"] * grp), 1.0 / (ws["W32"] * act),
                           1.0 / (ws["W4"] * act), 1.0 / (ws["WR"] * grp), act, grp, 0.0], dtype=torch.float64)
    f32 = torch.cat([f["W1"].reshape(-1), f["bias_c"], f["b4"], f["bias_tok"], scales]).to(torch.float32)
    assert f32.numel() * 4 <= F32_SECTION_BYTES
    head = torch.zeros(F32_SECTION_BYTES, dtype=torch.uint8)
    head[: f32.numel() * 4] = f32.view(torch.uint8)
    parts = [head] + [pack_kmajor((f[name] * ws[name]).to(torch.float32), dtype, split) for name, _, _ in SECTIONS]
    parts.append(layer1_image(f["W1"] * act, dtype))
    parts.append(f["W32"].to(torch.float32).contiguous().view(torch.uint8).reshape(-1))
    blob = torch.cat(parts)
    assert blob.numel() == packed_bytes(mode), (blob.numel(), packed_bytes(mode))
    return blob


# ---- pos_embed + cls rows (models/pointbert/point_encoder.py:135-142) ----------------------------------
POS_F32_SECTION_BYTES = 8192
POS_HIDDEN_SCALE = 64.0   # fp32-parity mode: GELU outputs as fp16 hi/lo operands


def pos_packed_bytes(mode):
    return POS_F32_SECTION_BYTES + 3 * 2 * split_of(mode) * IMAGE_BYTES


def pack_pos_embed(pos_embed_sd, cls_token, cls_pos, mode):
    """pos_embed = Sequential(Linear(3,128), GELU, Linear(128,384)) state dict (keys 0.weight, 0.bias, 2.weight,
    2.bias), cls_token / cls_pos [1,1,384] -> uint8 CPU blob for ppt_tokenizer_forward.

    fp32 section: W1 rows {w0,w1,w2,b1} [128][4] | b2 [384] | cls_token [384] | cls_pos [384] |
    scales {1/(weight scale * hidden scale), hidden scale, 0, 0}; then the 128 -> 384 weight as
    [3 units][2 chunks][split] K-major operand images (pack_kmajor)."""
    d = {k: v.detach().to(torch.float64).cpu() for k, v in pos_embed_sd.items()}
    w1, b1, w2, b2 = d["0.weight"], d["0.bias"], d["2.weight"], d["2.bias"]
    if tuple(w1.shape) != (128, 3) or tuple(w2.shape) != (384, 128):
        raise ValueError("kernels are specialised for pos_embed 3 -> 128 -> 384 (point_encoder.py:138-142)")
    dtype, split = operand_dtype(mode), split_of(mode)
    wscale = weight_scale(w2) if mode == ENC_FP16X3 else 1.0
    hscale = POS_HIDDEN_SCALE if mode == ENC_FP16X3 else 1.0
    scales = torch.tensor([1.0 / (wscale * hscale), hscale, 0.0, 0.0], dtype=torch.float64)
    f32 = torch.cat([torch.cat([w1, b1[:, None]], dim=1).reshape(-1), b2,
                     cls_token.detach().to(torch.float64).cpu().reshape(-1),
                     cls_pos.detach().to(torch.float64).cpu().reshape(-1), scales]).to(torch.float32)
    assert f32.numel() == 512 + 3 * 384 + 4
    head = torch.zeros(POS_F32_SECTION_BYTES, dtype=torch.uint8)
    head[: f32.numel() * 4] = f32.view(torch.uint8)
    blob = torch.cat([head, pack_kmajor((w2 * wscale).to(torch.float32), dtype, split)])
    assert blob.numel() == pos_packed_bytes(mode)
    return blob


# ---- train-mode BatchNorm (models/pointbert/dvae.py:190,196 under model.train()) -------------
def pack_encoder_train(sd, mode):
    """The blob ppt_encoder_forward_train works on: both BatchNorms packed as identities, i.e. the raw
    convolution weights.  The kernels fold the statistics of the current batch themselves (first_conv.1 into the
    W1' sections of this blob, which is why the device copy must be the caller's own mutable one; second_conv.1
    as per-channel scale / shift applied to the accumulators)."""
    ident = dict(sd)
    for conv, c in (("first_conv.1", 128), ("second_conv.1", 512)):
        ref = sd[conv + ".weight"]
        ident[conv + ".weight"] = torch.ones(c, dtype=ref.dtype)
        ident[conv + ".bias"] = torch.zeros(c, dtype=ref.dtype)
        ident[conv + ".running_mean"] = torch.zeros(c, dtype=ref.dtype)
        ident[conv + ".running_var"] = torch.full((c,), 1.0 - 1e-5, dtype=torch.float64)  # fold(): var + eps == 1
    return pack_encoder(ident, mode)


# ---- PointNet++ set-abstraction shared MLP (models/pointnet2/pointnet2_utils.py:166-172, 213-226) ---------------
def fold_conv_bn(conv_w, conv_b, bn_w, bn_b, bn_mean, bn_var, eps):
    """Conv (1x1) followed by eval-mode BatchNorm as one affine map: W' [Cout, Cin], b' [Cout] (fp64)."""
    w = conv_w.detach().to(torch.float64).cpu().reshape(conv_w.shape[0], -1)
    s = bn_w.detach().to(torch.float64).cpu() / torch.sqrt(bn_var.detach().to(torch.float64).cpu() + eps)
    b = conv_b.detach().to(torch.float64).cpu() if conv_b is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    return w * s[:, None], (b - bn_mean.detach().to(torch.float64).cpu()) * s + bn_b.detach().to(torch.float64).cpu()


def sa_mlp_packed_bytes(c0, c1, c2, c3):
    kc0, u1, u2, u3 = (c0 + 63) // 64, (c1 + 127) // 128, (c2 + 127) // 128, (c3 + 127) // 128
    bias = ((u1 + u2 + u3) * 128 * 4 + 1023) // 1024 * 1024
    return bias + (u1 * kc0 + u2 * 2 * u1 + u3 * 2 * u2) * IMAGE_BYTES


def pack_sa_mlp(convs, bns, xyz_first, mode):
    """Three Conv2d(1x1) + BatchNorm2d pairs (eval mode) -> uint8 CPU blob for ppt_sa_mlp_forward.
    xyz_first: the module concatenates [xyz, features] (PointNetSetAbstraction, sample_and_group) rather than
    [features, xyz] (PointNetSetAbstractionMsg); the kernel's order is [features, xyz], so layer 1's columns move."""
    if len(convs) != 3 or mode not in (ENC_FP16, ENC_BF16):
        raise ValueError("fused SA MLP: exactly three layers, fp16 or bf16 operands")
    dtype = operand_dtype(mode)
    folded = [fold_conv_bn(c.weight, c.bias, b.weight, b.bias, b.running_mean, b.running_var, b.eps)
              for c, b in zip(convs, bns)]
    w1 = folded[0][0]
    if xyz_first:
        w1 = torch.cat([w1[:, 3:], w1[:, :3]], dim=1)
    ws = [w1, folded[1][0], folded[2][0]]
    c0, c1, c2, c3 = w1.shape[1], w1.shape[0], ws[1].shape[0], ws[2].shape[0]
    if ws[1].shape[1] != c1 or ws[2].shape[1] != c2:
        raise ValueError("layer shapes do not chain")
    u = [(c + 127) // 128 for c in (c1, c2, c3)]
    kin = [(c0 + 63) // 64 * 64, 2 * u[0] * 64, 2 * u[1] * 64]
    bias = torch.zeros(sum(u) * 128, dtype=torch.float32)
    parts, off = [], 0
    for l in range(3):
        wp = torch.zeros(u[l] * 128, kin[l], dtype=torch.float32)
        wp[: ws[l].shape[0], : ws[l].shape[1]] = ws[l].to(torch.float32)
        parts.append(pack_kmajor(wp, dtype, 1))
        bias[off: off + ws[l].shape[0]] = folded[l][1].to(torch.float32)
        off += u[l] * 128
    head = torch.zeros((bias.numel() * 4 + 1023) // 1024 * 1024, dtype=torch.uint8)
    head[: bias.numel() * 4] = bias.view(torch.uint8)
    blob = torch.cat([head] + parts)
    assert blob.numel() == sa_mlp_packed_bytes(c0, c1, c2, c3), (blob.numel(), sa_mlp_packed_bytes(c0, c1, c2, c3))
    return blob, (c0, c1, c2, c3)


# ---- PointNetFeaturePropagation MLP (models/pointnet2/pointnet2_utils.py:273-279, 316-319) -------------------------
def fp_mlp_packed_bytes(c0, c1, c2):
    kc0, u1, u2 = (c0 + 63) // 64, (c1 + 127) // 128, (c2 + 127) // 128
    bias = ((u1 + u2) * 128 * 4 + 1023) // 1024 * 1024
    return bias + (u1 * kc0 + u2 * 2 * u1) * IMAGE_BYTES


def pack_fp_mlp(convs, bns, d1, mode):
    """Two Conv1d(1x1) + BatchNorm1d pairs (eval mode) -> uint8 CPU blob for ppt_fp_mlp_forward.  The module's input is
    cat([points1 (d1 channels), interpolated]); the kernel's operand order is [interpolated | points1], so layer 1's
    columns move (the interpolated rows are then read with aligned 16-byte loads)."""
    if len(convs) != 2 or mode not in (ENC_FP16, ENC_BF16):
        raise ValueError("fused FP MLP: exactly two layers, fp16 or bf16 operands")
    dtype = operand_dtype(mode)
    folded = [fold_conv_bn(c.weight, c.bias, b.weight, b.bias, b.running_mean, b.running_var, b.eps)
              for c, b in zip(convs, bns)]
    w1 = folded[0][0]
    w1 = torch.cat([w1[:, d1:], w1[:, :d1]], dim=1)
    w2 = folded[1][0]
    c0, c1, c2 = w1.shape[1], w1.shape[0], w2.shape[0]
    if w2.shape[1] != c1:
        raise ValueError("layer shapes do not chain")
    u1, u2 = (c1 + 127) // 128, (c2 + 127) // 128
    bias = torch.zeros((u1 + u2) * 128, dtype=torch.float32)
    bias[:c1] = folded[0][1].to(torch.float32)
    bias[u1 * 128: u1 * 128 + c2] = folded[1][1].to(torch.float32)
    parts = []
    for w, rows, k in ((w1, u1 * 128, (c0 + 63) // 64 * 64), (w2, u2 * 128, 2 * u1 * 64)):
        wp = torch.zeros(rows, k, dtype=torch.float32)
        wp[: w.shape[0], : w.shape[1]] = w.to(torch.float32)
        parts.append(pack_kmajor(wp, dtype, 1))
    head = torch.zeros((bias.numel() * 4 + 1023) // 1024 * 1024, dtype=torch.uint8)
    head[: bias.numel() * 4] = bias.view(torch.uint8)
    blob = torch.cat([head] + parts)
    assert blob.numel() == fp_mlp_packed_bytes(c0, c1, c2), (blob.numel(), fp_mlp_packed_bytes(c0, c1, c2))
    return blob, (c0, c1, c2)
