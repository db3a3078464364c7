"""Reference-format checkpoints -> the B200 component shells.

The on-disk layout is the reference's own (`ModelMixin.save_pretrained` / `from_pretrained`,
models/modeling_utils.py:680-790, 886-1400; `ConfigMixin.save_config`, configuration_utils.py:146-186;
`DiffusionPipeline.save_pretrained`, pipelines/pipeline_utils.py:234-340):

    <root>/model_index.json                      {"_class_name": ..., "unet": ["diffusers", "UNet2DConditionModel"], ...}
    <root>/<component>/config.json               constructor arguments + "_class_name" + "_diffusers_version"
    <root>/<component>/diffusion_pytorch_model[.<variant>].safetensors
    <root>/<component>/diffusion_pytorch_model[.<variant>].safetensors.index.json   (sharded: {"weight_map": {key: shard file}})
    <root>/scheduler/scheduler_config.json

Tensors keep the reference's parameter names; the shells re-pack them at construction (packing.py), so loading is:
read config -> read every tensor the shell's parameter spec names -> constructor.  Nothing here touches the network:
`path` is a local directory (the reference's hub download is out of scope).
"""
import json
import os

import torch

CONFIG_NAME = "config.json"                                   # utils/constants.py:29
WEIGHTS_NAME = "diffusion_pytorch_model.bin"                  # :30
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"  # :33
SCHEDULER_CONFIG_NAME = "scheduler_config.json"               # schedulers/scheduling_utils.py:27
MODEL_INDEX_NAME = "model_index.json"                         # pipelines/pipeline_utils.py:216

# AutoencoderKL checkpoints older than the Attention refactor name the mid-block attention weights like this
# (models/modeling_utils.py `_convert_deprecated_attention_blocks`)
_DEPRECATED_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def add_variant(weights_name, variant=None):
    """utils/hub_utils.py:212 - 'diffusion_pytorch_model.safetensors' + 'fp16' -> 'diffusion_pytorch_model.fp16.safetensors'."""
    if variant is None:
        return weights_name
    parts = weights_name.split(".")
    return ".".join(parts[:-1] + [variant] + parts[-1:])


def load_config(path, config_name=CONFIG_NAME):
    """Constructor arguments from a reference config file (private keys such as `_class_name` are returned too)."""
    f = path if os.path.isfile(path) else os.path.join(path, config_name)
    if not os.path.isfile(f):
        raise EnvironmentError(f"no {config_name} under {path}")
    with open(f) as fh:
        return json.load(fh)


def public_config(cfg):
    """Drop the bookkeeping keys (`_class_name`, `_diffusers_version`, `_name_or_path`, ...)."""
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if not k.startswith("_")}


def _safetensors_load(file, wanted=None):
    from safetensors import safe_open
    out = {}
    with safe_open(file, framework="pt", device="cpu") as f:
        for k in f.keys():
            if wanted is None or k in wanted:
                out[k] = f.get_tensor(k)
    return out


def load_state_dict(path, variant=None, wanted=None, safetensors_name=SAFETENSORS_WEIGHTS_NAME, pickle_name=WEIGHTS_NAME):
    """All tensors (or only the names in `wanted`) of one component directory: single safetensors file, sharded
    safetensors with an index, or the legacy pickle (`weights_only`).  transformers components (text encoders) use
    safetensors_name='model.safetensors', pickle_name='pytorch_model.bin'."""
    SAFETENSORS_WEIGHTS_NAME, WEIGHTS_NAME = safetensors_name, pickle_name  # noqa: N806  (shadow the module defaults below)
    st = os.path.join(path, add_variant(SAFETENSORS_WEIGHTS_NAME, variant))
    idx = os.path.join(path, add_variant(SAFETENSORS_WEIGHTS_NAME + ".index.json", variant))
    if not os.path.isfile(idx):
        # the reference writes the index as '<name>.safetensors.index[.variant].json' for variants
        alt = os.path.join(path, add_variant(SAFETENSORS_WEIGHTS_NAME, None) + ".index" + (f".{variant}" if variant else "") + ".json")
        idx = alt if os.path.isfile(alt) else idx
    if os.path.isfile(st):
        return _safetensors_load(st, wanted)
    if os.path.isfile(idx):
        with open(idx) as fh:
            wm = json.load(fh)["weight_map"]
        out = {}
        for shard in sorted(set(wm.values())):
            names = {k for k, v in wm.items() if v == shard and (wanted is None or k in wanted)}
            if names:
                out.update(_safetensors_load(os.path.join(path, shard), names))
        return out
    pt = os.path.join(path, add_variant(WEIGHTS_NAME, variant))
    if os.path.isfile(pt):
        sd = torch.load(pt, map_location="cpu", weights_only=True)
        return sd if wanted is None else {k: v for k, v in sd.items() if k in wanted}
    raise EnvironmentError(f"no {add_variant(SAFETENSORS_WEIGHTS_NAME, variant)} (or sharded index, or {WEIGHTS_NAME}) under {path}")


def convert_deprecated_attention_keys(sd):
    """`...attentions.N.query.weight` -> `...attentions.N.to_q.weight` (old AutoencoderKL checkpoints)."""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts:
            i = parts.index("attentions")
            if i + 2 < len(parts) and parts[i + 2] in _DEPRECATED_ATTN:
                parts[i + 2:i + 3] = _DEPRECATED_ATTN[parts[i + 2]].split(".")
                k = ".".join(parts)
        out[k] = v
    return out


class FromPretrainedMixin:
    """`Model.from_pretrained(path, subfolder=None, variant=None, torch_dtype=bf16, device="cuda")` for the shells.

    Class attributes a shell sets: `_ref_class_names` (accepted `_class_name` values), `_param_spec(cfg) -> {name: shape}`
    (the reference parameter names the shell consumes) and optionally `_fix_keys(state_dict)`.
    """
    _ref_class_names = ()

    @classmethod
    def _param_spec(cls, cfg):
        raise NotImplementedError

    @staticmethod
    def _fix_keys(sd):
        return sd

    @classmethod
    def from_pretrained(cls, path, subfolder=None, variant=None, torch_dtype=torch.bfloat16, device="cuda", fold_norms=None, **unsupported):
        """fold_norms (UNet2DConditionModel only): None = the class default (LayerNorms folded into the GEMMs, weights re-rounded
        once), False = keep bit-exact copies of the checkpoint's weights."""
        for k, v in unsupported.items():
            if v not in (None, False):
                raise NotImplementedError(f"from_pretrained option {k}={v!r} is outside the hot path (local directories only)")
        root = os.path.join(path, subfolder) if subfolder else path
        raw = load_config(root)
        name = raw.get("_class_name")
        if name is not None and cls._ref_class_names and name not in cls._ref_class_names:
            raise ValueError(f"{root} holds a {name}, not one of {cls._ref_class_names}")
        cfg = public_config(raw)
        sd = cls._fix_keys(load_state_dict(root, variant=variant))
        spec = cls._param_spec_for(cfg)
        missing = [k for k in spec if k not in sd]
        if missing:
            raise ValueError(f"{root}: checkpoint lacks {len(missing)} tensors the model needs, e.g. {missing[:3]}")
        extra = {} if fold_norms is None else dict(fold_norms=fold_norms)
        keep = {k: sd[k] for k in spec}
        opt = cls._optional_param_spec(cfg)
        if opt and all(k in sd for k in opt):  # e.g. the encoder half of an AutoencoderKL checkpoint
            keep.update({k: sd[k] for k in opt})
        return cls(cfg, keep, dtype=torch_dtype, device=device, **extra)

    @classmethod
    def _optional_param_spec(cls, cfg):
        return {}

    @classmethod
    def _param_spec_for(cls, cfg):
        return cls._param_spec(cfg)

    def save_pretrained(self, save_directory, safe_serialization=True, variant=None):
        """The reference's `ModelMixin.save_pretrained` layout (models/modeling_utils.py:680-790): config.json with
        `_class_name` + one `diffusion_pytorch_model[.variant].safetensors` holding `reference_state_dict()`, so the
        unmodified reference (or `from_pretrained` here) loads it back."""
        if not safe_serialization:
            raise NotImplementedError("only safetensors serialization")
        if not hasattr(self, "reference_state_dict"):
            raise NotImplementedError(f"{type(self).__name__} cannot rebuild the reference's parameters from its packed buffers")
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        cfg = {"_class_name": self._ref_class_names[0], "_diffusers_version": "0.40.0.dev0"}
        cfg.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in dict(self.config).items()})
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        save_file(self.reference_state_dict(), os.path.join(save_directory, add_variant(SAFETENSORS_WEIGHTS_NAME, variant)),
                  metadata={"format": "pt"})


def scheduler_kwargs(cls, cfg):
    """Constructor arguments out of a reference scheduler config.  Options the class does not name reach its `**kwargs`
    (where unsupported non-default values raise NotImplementedError) - never dropped silently."""
    import inspect
    params = inspect.signature(cls.__init__).parameters
    accepted = set(params) - {"self"}
    has_var_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
    cfg = public_config(dict(cfg))
    dropped = [k for k in cfg if k not in accepted and not has_var_kw]
    if dropped:
        raise NotImplementedError(f"{cls.__name__} does not take {dropped}")
    return {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}


def scheduler_from_pretrained(cls, path, subfolder=None):
    """`SchedulerMixin.from_pretrained` (schedulers/scheduling_utils.py:100-160): constructor arguments from
    scheduler_config.json."""
    root = os.path.join(path, subfolder) if subfolder else path
    raw = load_config(root, SCHEDULER_CONFIG_NAME)
    name = raw.get("_class_name")
    if name is not None and name != cls.__name__:
        raise NotImplementedError(f"{root} configures a {name}; this path implements {cls.__name__}")
    return cls(**scheduler_kwargs(cls, raw))


def load_model_index(path):
    """model_index.json -> {component: (library, class name)} (pipelines/pipeline_utils.py:224-232 register_modules)."""
    idx = load_config(path, MODEL_INDEX_NAME)
    comps = {k: tuple(v) for k, v in idx.items()
             if not k.startswith("_") and isinstance(v, list) and len(v) == 2 and v[1] is not None}  # [null, null] = empty slot
    return idx.get("_class_name"), comps


def load_pipeline_components(path, expected_class, wanted, torch_dtype=torch.bfloat16, device="cuda", variant=None):
    """Instantiate the components named in `wanted` = {slot: shell class or scheduler class} from a reference pipeline
    directory.  Slots of the reference pipeline that are not on the path (text encoders, tokenizers, ...) are ignored:
    the shells take precomputed embeddings."""
    cls_name, comps = load_model_index(path)
    if cls_name is not None and cls_name != expected_class:
        raise ValueError(f"{path} is a {cls_name} checkpoint, not {expected_class}")
    out = {}
    for slot, shell in wanted.items():
        if slot not in comps:
            raise ValueError(f"{path}/model_index.json has no '{slot}' component")
        ref_cls = comps[slot][1]
        if not isinstance(shell, (tuple, list)) and hasattr(shell, "_ref_class_names"):
            if ref_cls not in shell._ref_class_names:
                raise NotImplementedError(f"{slot}: {ref_cls} is not on the accelerated path ({shell._ref_class_names})")
            out[slot] = shell.from_pretrained(path, subfolder=slot, variant=variant, torch_dtype=torch_dtype, device=device)
        else:
            choices = shell if isinstance(shell, (tuple, list)) else (shell,)  # scheduler slot: any of the implemented steppers
            match = [c for c in choices if c.__name__ == ref_cls]
            if not match:
                raise NotImplementedError(f"{slot}: {ref_cls} is not on the accelerated path ({[c.__name__ for c in choices]})")
            out[slot] = scheduler_from_pretrained(match[0], path, subfolder=slot)
    return out
