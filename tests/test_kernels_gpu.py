"""GPU parity of every sm_100a kernel against a plain fp32 PyTorch evaluation of the same op (inputs are the
bf16/fp16 values up-cast, TF32 off).  Cases live in tools/diag_gemm.py and tools/diag_ops.py so the same code is
the command-line diagnostic.  Tolerances (in those files): bf16 outputs within 1.5e-2*|ref| + 2e-2 for GEMM/conv
(K up to 11520 products of bf16 inputs, one final rounding), attention within 2e-2*|ref| + 6e-3, norms within
1e-2*|ref| + 1e-2; layout / timestep-embedding / scheduler-step kernels bit-exact against the eager formula."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tools import diag_gemm, diag_ops  # noqa: E402


@pytest.mark.parametrize("name", list(diag_gemm.CASES))
def test_conv_gemm(name):
    assert diag_gemm.run_case(name)


@pytest.mark.parametrize("name", list(diag_ops.CASES))
def test_ops(name):
    assert diag_ops.run_case(name)


def test_unfused_attention_d512():
    import torch.nn.functional as F
    from diffusers_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    S, D = 1024, 512
    qkv = (torch.randn(S, 3 * D, generator=g, device="cuda")).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    out = ops.attention_unfused(q, k, v, scale=D ** -0.5)
    ref = F.scaled_dot_product_attention(q.float()[None, None], k.float()[None, None], v.float()[None, None])[0, 0]
    assert (out.float() - ref).abs().max() < 6e-3


def test_group_norm_is_deterministic_and_counter_self_resets():
    from diffusers_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2 * 4096, 320, generator=g, device="cuda").bfloat16()
    outs = [ops.group_norm(x, batch=2, hw=4096, groups=32, eps=1e-5, silu=True) for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


def test_conv_gemm_rejects_bad_arguments():
    from diffusers_b200 import ops
    x = torch.zeros(128, 12, dtype=torch.bfloat16, device="cuda")  # channels not a multiple of 8
    w = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ops.B200Error):
        ops.linear(x, w, 64)
    x = torch.zeros(1, 15, 15, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 9 * 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ops.B200Error):  # stride 2 needs even H, W
        ops.conv_gemm(x.view(-1, 64), w, 64, batch=1, H=15, W=15, ksize=3, stride=2)


def test_attention_processor_plugin_matches_reference_processor_math():
    """B200AttnProcessor on a module shaped like the reference's `Attention` (to_q/to_k/to_v/to_out, heads) against
    AttnProcessor2_0's op sequence (models/attention_processor.py:2705-2789) in fp32."""
    import torch.nn as nn
    import torch.nn.functional as F
    from diffusers_b200.attention_processor import B200AttnProcessor, b200_attention_backend

    class FakeAttention(nn.Module):
        def __init__(self, dim, cross, heads):
            super().__init__()
            self.heads = heads
            self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim, bias=False), nn.Linear(cross, dim, bias=False), nn.Linear(cross, dim, bias=False)
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
            self.residual_connection, self.rescale_output_factor = False, 1.0

    torch.manual_seed(0)
    attn = FakeAttention(640, 2048, 10).cuda().bfloat16()
    x = torch.randn(2, 256, 640, device="cuda").bfloat16()
    ctx = torch.randn(2, 77, 2048, device="cuda").bfloat16()
    proc = B200AttnProcessor()
    for enc in (None, ctx):
        a = FakeAttention(640, 2048 if enc is not None else 640, 10).cuda().bfloat16() if enc is None else attn
        out = proc(a, x, encoder_hidden_states=enc)
        c = x if enc is None else enc
        lin = lambda m, t: F.linear(t.float(), m.weight.float(), None if m.bias is None else m.bias.float())  # noqa: E731
        q, k, v = lin(a.to_q, x), lin(a.to_k, c), lin(a.to_v, c)
        sp = lambda t: t.view(2, -1, 10, 64).transpose(1, 2)  # noqa: E731
        ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(2, 256, 640)
        ref = lin(a.to_out[0], ref)
        assert (out.float() - ref).abs().max() < 3e-2
    q = torch.randn(1, 512, 4, 128, device="cuda").bfloat16()
    o = b200_attention_backend(q, q, q)
    ref = F.scaled_dot_product_attention(*(t.float().permute(0, 2, 1, 3) for t in (q, q, q))).permute(0, 2, 1, 3)
    assert o.shape == q.shape and (o.float() - ref).abs().max() < 2e-2
    with pytest.raises(NotImplementedError):
        b200_attention_backend(q, q, q, is_causal=True)


@pytest.mark.parametrize("cm", [1, 2])
def test_conv_gemm_cluster_multicast_sizes(cm):
    """One CTA per tile (1) vs CTA pairs driving a cta_group::2 MMA (2) is normally picked by the cost model; force each
    through the B200_FORCE_CM test knob in a fresh process (the knob is read once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200_FORCE_CM=str(cm))
    cases = ["lin_sdxl", "lin_ragged", "lin_geglu", "lin_2src", "conv_32", "conv_s2", "conv_odd", "lin_bn160", "lin_n4"]
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "diag_gemm.py"), "--inproc", *cases], env=env,
                       capture_output=True, text=True, timeout=600)
    assert "SUMMARY" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    summary = p.stdout.split("SUMMARY", 1)[1]
    assert "FAIL" not in summary and "ERROR" not in summary, p.stdout[-3000:]



@pytest.mark.parametrize("rows,K,blocks,bcols,hd", [(576, 256, 8, 384, 128), (2304, 128, 2, 1536, 128), (200, 64, 3, 384, 64)])
def test_conv_gemm_column_blocks_to_separate_buffers(rows, K, blocks, bcols, hd):
    """b200_conv_gemm_args.y_peers (context parallelism: column block j of the output goes to rank j's buffer), here with local
    buffers: every block must equal the matching columns of the plain launch bit for bit, with and without the q/k RMSNorm + RoPE
    epilogue whose [q | k | v] pattern repeats per block."""
    from diffusers_b200 import ops, packing
    g = torch.Generator(device="cuda").manual_seed(7)
    N = blocks * bcols
    x = torch.randn(rows, K, generator=g, device="cuda").bfloat16()
    w = packing.pack_linear_weight(torch.randn(N, K, generator=g, device="cuda").float() * K ** -0.5).bfloat16().cuda()
    b = (torch.randn(N, generator=g, device="cuda") * 0.2).bfloat16()
    plain = ops.linear(x, w, N, bias=b)
    bufs = [torch.full((rows + 3, bcols + 64), 7.0, dtype=torch.bfloat16, device="cuda") for _ in range(blocks)]
    views = [t[3:, 64:] for t in bufs] if bcols % 8 == 0 else None  # offset views: separate bases, same row stride
    ops.linear(x, w, N, bias=b, out_blocks=views)
    torch.cuda.synchronize()
    for j, v in enumerate(views):
        assert torch.equal(v, plain[:, j * bcols:(j + 1) * bcols]), j
        assert float((bufs[j][:3] - 7).abs().max()) == 0 and float((bufs[j][:, :64] - 7).abs().max()) == 0  # nothing outside the view
    # q/k pattern per block: compare with the one-block-at-a-time launches
    qk = 2 * (bcols // 3)
    if qk % (2 * hd) == 0:
        ang = torch.rand(rows, hd // 2, generator=g, device="cuda") * 6.28
        cos, sin = torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous()
        cT, sT = ops.rope_tables_transposed(cos, sin)
        nw = (torch.randn(2, hd, generator=g, device="cuda") * 0.1 + 1).bfloat16()
        rope = ops.QkRope(nw, cT, sT, 0, qk, hd)
        ops.linear(x, w, N, bias=b, out_blocks=views, qk_rope=rope)
        for j, v in enumerate(views):
            one = ops.linear(x, w[j * bcols:(j + 1) * bcols], bcols, bias=b[j * bcols:(j + 1) * bcols], qk_rope=rope)
            torch.cuda.synchronize()
            assert torch.equal(v, one), j


@pytest.mark.parametrize("Sq,H,D,seg", [(4608, 3, 128, 576), (1152, 2, 64, 288), (2304, 2, 128, 1152)])
def test_attention_scatters_output_rows_to_segment_buffers(Sq, H, D, seg):
    """b200_attention_args.o_seg (context parallelism: output row r goes to the buffer of the rank that owns it), here with local
    buffers and segment sizes that are NOT multiples of the 128-row query tile (8 ranks: 576 rows): every segment equals the
    matching rows of the plain launch bit for bit, at this rank's head columns of the owners' wider [rows, all heads] buffers."""
    from diffusers_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(1, Sq, 3 * H * D, generator=g, device="cuda").bfloat16()
    q, k, v = qkv[:, :, :H * D], qkv[:, :, H * D:2 * H * D], qkv[:, :, 2 * H * D:]
    plain = ops.attention(q, k, v, heads=H, head_dim=D)
    nseg = Sq // seg
    wide = [torch.full((seg, 4 * H * D), 3.0, dtype=torch.bfloat16, device="cuda") for _ in range(nseg)]
    views = [w[:, H * D:2 * H * D] for w in wide]
    assert ops.attention(q, k, v, heads=H, head_dim=D, o_seg=views, o_seg_rows=seg) is None
    torch.cuda.synchronize()
    for s_, w in enumerate(wide):
        assert torch.equal(w[:, H * D:2 * H * D], plain[0, s_ * seg:(s_ + 1) * seg]), s_
        assert float((w[:, :H * D] - 3).abs().max()) == 0 and float((w[:, 2 * H * D:] - 3).abs().max()) == 0
