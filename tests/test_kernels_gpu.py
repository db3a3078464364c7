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
