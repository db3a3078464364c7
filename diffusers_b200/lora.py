"""LoRA adapters merged into the base weights BEFORE packing ("LoRA-merged weights").

The reference applies LoRA through peft layers and can fuse them (`fuse_lora`, loaders/lora_base.py:544: W <- W + scale * (alpha / r) * B A).
The kernels here run on packed weights, so the only form on the path is the fused one: merge into the reference-named state dict, then build
the shell (or `from_pretrained`) as usual - no extra launch, no second code path.

Accepted adapter keys (the diffusers serialisation of `save_lora_weights`, loaders/lora_pipeline.py): `<prefix>.<module>.lora_A.weight` /
`.lora_B.weight` (peft naming), the older `<module>.lora.down.weight` / `.lora.up.weight` and `<module>.lora_linear_layer.down/up.weight`,
with an optional `<module>.alpha` (or a `network_alphas` dict) - default alpha = rank.  Kohya-named files need the reference's converter first.
"""
import re

import torch

_PAIRS = (("lora_A.weight", "lora_B.weight"), ("lora.down.weight", "lora.up.weight"), ("lora_linear_layer.down.weight", "lora_linear_layer.up.weight"))


def _targets(lora_sd, prefix):
    """{module path (without prefix): (down key, up key)}"""
    out = {}
    pre = prefix + "." if prefix else ""
    for k in lora_sd:
        if pre and not k.startswith(pre):
            continue
        for down, up in _PAIRS:
            if k.endswith("." + down):
                mod = k[len(pre):-len(down) - 1]
                mod = re.sub(r"\.(default|default_0)$", "", mod)  # peft adapter-name infix "lora_A.default.weight" is handled below
                out[mod] = (k, k[:-len(down)] + up)
    # peft with an adapter name: "...lora_A.<adapter>.weight"
    for k in lora_sd:
        m = re.match(r"^(.*)\.lora_A\.([^.]+)\.weight$", k)
        if m and (not pre or k.startswith(pre)):
            mod = m.group(1)[len(pre):]
            out[mod] = (k, f"{m.group(1)}.lora_B.{m.group(2)}.weight")
    return out


def merge_lora(state_dict, lora_state_dict, prefix="unet", scale=1.0, network_alphas=None, strict=True):
    """-> a new state dict with W + scale * (alpha / rank) * (B @ A) for every module the adapter names (Linear [N, K]; Conv2d: A is the
    [r, C_in, kh, kw] "down" convolution and B the [C_out, r, 1, 1] "up" one).  fp32 arithmetic, result in the base weight's dtype.
    strict: an adapter module that the model does not have is an error (a silent skip would change the image)."""
    network_alphas = network_alphas or {}
    out = dict(state_dict)
    merged = []
    for mod, (kd, ku) in sorted(_targets(lora_state_dict, prefix).items()):
        wk = mod + ".weight"
        if wk not in state_dict:
            if strict:
                raise KeyError(f"LoRA targets '{mod}' ({kd}) but the model has no '{wk}'")
            continue
        if ku not in lora_state_dict:
            raise KeyError(f"LoRA has {kd} without its up-projection {ku}")
        A, B = lora_state_dict[kd].float(), lora_state_dict[ku].float()
        W = state_dict[wk]
        r = A.shape[0]
        if B.shape[1] != r or B.shape[0] != W.shape[0] or A.flatten(1).shape[1] != W.flatten(1).shape[1]:
            raise ValueError(f"LoRA shapes of '{mod}' do not fit the weight: A {tuple(A.shape)}, B {tuple(B.shape)}, W {tuple(W.shape)}")
        pre = (prefix + "." if prefix else "")
        alpha = None
        for cand in (pre + mod + ".alpha", mod + ".alpha"):
            if cand in lora_state_dict:
                alpha = float(lora_state_dict[cand])
            elif cand in network_alphas:
                alpha = float(network_alphas[cand])
        alpha = r if alpha is None else alpha
        delta = (B.flatten(1) @ A.flatten(1)).reshape(W.shape)
        out[wk] = (W.float() + (scale * alpha / r) * delta).to(W.dtype)
        merged.append(mod)
    if strict and not merged:
        raise KeyError(f"the adapter has no keys under prefix '{prefix}' in a supported naming (lora_A / lora_B, lora.down / lora.up)")
    return out
