"""Weight re-layout for libb200diff.so (done once at load time, plain torch on any device).

conv_gemm weights are K-major [N, Kp] with k = ((tap * nsrc + src) * rup64(C_src)) + channel, zero padded so
that every 64-wide K chunk the kernel fetches belongs to exactly one (tap, source) pair
(include/b200_diffusion.h, b200_conv_gemm).
"""
import torch


def rup(v, m):
    return (v + m - 1) // m * m


def packed_k(ksize, c0, c1=0):
    return ksize * ksize * (rup(c0, 64) + (rup(c1, 64) if c1 > 0 else 0))


def pack_conv_weight(w, split=None):
    """nn.Conv2d weight [O, I, kh, kw] -> [O, Kp].  `split` = (c0, c1) when the conv input is the channel
    concatenation of two tensors (torch.cat([h, skip], dim=1), reference unet_2d_blocks.py:2444)."""
    O, I, kh, kw = w.shape
    assert kh == kw and kh in (1, 3)
    srcs = [I] if split is None else list(split)
    assert sum(srcs) == I
    out = w.new_zeros(O, kh * kw, sum(rup(c, 64) for c in srcs))
    wt = w.permute(0, 2, 3, 1).reshape(O, kh * kw, I)  # [O, tap, I]
    src_off, dst_off = 0, 0
    for c in srcs:
        out[:, :, dst_off:dst_off + c] = wt[:, :, src_off:src_off + c]
        src_off += c
        dst_off += rup(c, 64)
    return out.reshape(O, -1).contiguous()


def pack_linear_weight(w, split=None):
    """nn.Linear weight [N, K] -> [N, Kp] (1x1 'conv' over a row image)."""
    return pack_conv_weight(w[:, :, None, None], split)


def pack_geglu(w, b, tile_n):
    """GEGLU proj weight [2*inner, K] / bias [2*inner] (value rows first, gate rows second,
    reference activations.py:113-123 `hidden_states, gate = chunk(2)`) -> per-tile interleave
    [tile_n/2 value rows | tile_n/2 gate rows] so one CTA tile holds matching value/gate columns."""
    two_inner = w.shape[0]
    inner = two_inner // 2
    half = tile_n // 2
    assert inner % half == 0, (inner, tile_n)
    wv, wg = w[:inner], w[inner:]
    nt = inner // half
    wp = torch.stack([wv.reshape(nt, half, -1), wg.reshape(nt, half, -1)], dim=1).reshape(two_inner, -1)
    bp = None
    if b is not None:
        bv, bg = b[:inner], b[inner:]
        bp = torch.stack([bv.reshape(nt, half), bg.reshape(nt, half)], dim=1).reshape(two_inner).contiguous()
    return pack_linear_weight(wp), bp


def pack_upsample_conv(w):
    """nn.Conv2d 3x3 weight [O, I, 3, 3] applied AFTER a nearest-2x upsample -> four packed [O, 4 * rup64(I)] weights, one per output
    parity class (ph, pw) in the order 2 * ph + pw (b200_conv_gemm ksize = 2, up2x_parity = 1 + index).  Output row 2i + ph reads
    input rows {i + ph - 1, i + ph}: for ph = 0 filter row 0 lands on i - 1 and rows 1, 2 on i; for ph = 1 rows 0, 1 land on i and
    row 2 on i + 1 (columns alike).  Taps that share an input pixel are summed in fp32 and rounded once."""
    O, I, kh, kw = w.shape
    assert kh == 3 and kw == 3
    wf = w.to(torch.float32)
    sets = (((0,), (1, 2)), ((0, 1), (2,)))  # [parity][tap] -> filter rows / columns merged into that tap
    out = []
    for ph in range(2):
        for pw in range(2):
            m = wf.new_zeros(O, I, 2, 2)
            for a in range(2):
                for b in range(2):
                    for r in sets[ph][a]:
                        for s_ in sets[pw][b]:
                            m[:, :, a, b] += wf[:, :, r, s_]
            cp = rup(I, 64)
            packed = wf.new_zeros(O, 4, cp)
            packed[:, :, :I] = m.permute(0, 2, 3, 1).reshape(O, 4, I)
            out.append(packed.reshape(O, -1).to(w.dtype).contiguous())
    return out


def fold_layer_norm(w, gamma, beta, bias, dtype):
    """nn.LayerNorm(gamma, beta) followed by nn.Linear(w, bias), as the operands of b200_conv_gemm's folded form:
        LN(x) W^T + b = rstd * (x W'^T) + (b + W beta),   W' = W*gamma - rowmean(W*gamma)
    (subtracting the row means makes x W'^T = (x - mean(x)) (W gamma)^T, so the mean never has to be applied).
    Returns (W' rounded to `dtype` [N, K], bias b + W beta in `dtype` [N], the subtracted row means fp32 [N] - kept only so
    that unfold_layer_norm can invert the fold)."""
    wf = w.to(torch.float32)
    wg = wf if gamma is None else wf * gamma.to(torch.float32)[None, :]
    shift = wg.mean(dim=1)
    lb = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device) if beta is None else wf @ beta.to(torch.float32)
    if bias is not None:
        lb = lb + bias.to(torch.float32)
    return (wg - shift[:, None]).to(dtype), lb.to(dtype).contiguous(), shift.contiguous()


def unfold_layer_norm(w_folded, gamma, shift):
    """Inverse of the weight part of fold_layer_norm, evaluated in fp32 and rounded once: the fold rounds W' to 16 bit, so the
    weight comes back within one 16-bit ulp, not bit-exact (construct the model with fold_norms=False to keep exact copies of
    a checkpoint's weights)."""
    wg = w_folded.to(torch.float32) + shift.to(torch.float32)[:, None]
    if gamma is not None:
        wg = wg / gamma.to(torch.float32)[None, :]
    return wg.to(w_folded.dtype)


# ----------------------------------------------------------------------------------------------------------------------
# inverses (used by the shells' reference_state_dict(): packed buffers -> the reference's parameter tensors)
# ----------------------------------------------------------------------------------------------------------------------
def unpack_conv_weight(p, in_channels, ksize, split=None):
    """[O, Kp] -> nn.Conv2d weight [O, I, k, k]; exact inverse of pack_conv_weight (the zero padding is dropped)."""
    O = p.shape[0]
    srcs = [in_channels] if split is None else list(split)
    assert sum(srcs) == in_channels
    taps = ksize * ksize
    per_tap = sum(rup(c, 64) for c in srcs)
    assert p.shape[1] == taps * per_tap, (tuple(p.shape), taps, per_tap)
    pt = p.reshape(O, taps, per_tap)
    parts, off = [], 0
    for c in srcs:
        parts.append(pt[:, :, off:off + c])
        off += rup(c, 64)
    w = torch.cat(parts, dim=2)  # [O, tap, I]
    return w.reshape(O, ksize, ksize, in_channels).permute(0, 3, 1, 2).contiguous()


def unpack_linear_weight(p, in_features, split=None):
    """[N, Kp] -> nn.Linear weight [N, K]."""
    return unpack_conv_weight(p, in_features, 1, split)[:, :, 0, 0].contiguous()


def unpack_geglu(wp, bp, in_features, tile_n):
    """Inverse of pack_geglu: per-tile [value rows | gate rows] -> [all value rows ; all gate rows]."""
    w = unpack_linear_weight(wp, in_features)
    two_inner = w.shape[0]
    inner, half = two_inner // 2, tile_n // 2
    nt = inner // half
    wt = w.reshape(nt, 2, half, -1)
    wo = torch.cat([wt[:, 0].reshape(inner, -1), wt[:, 1].reshape(inner, -1)], 0).contiguous()
    bo = None
    if bp is not None:
        bt = bp.reshape(nt, 2, half)
        bo = torch.cat([bt[:, 0].reshape(inner), bt[:, 1].reshape(inner)], 0).contiguous()
    return wo, bo
