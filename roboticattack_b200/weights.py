"""Parameter inventory of the hot path (HF checkpoint names) and synthetic initialisation.

Names follow the ``openvla/openvla-7b`` checkpoint as loaded by ``OpenVLAForActionPrediction``
(``prismatic/extern/hf/modeling_prismatic.py``): ``vision_backbone.featurizer.*`` (DINOv2),
``vision_backbone.fused_featurizer.*`` (SigLIP), ``projector.fc{1,2,3}``, ``language_model.model.*``,
``language_model.lm_head``.
"""
from __future__ import annotations

import torch

from .config import OpenVLAConfig

DINO = "vision_backbone.featurizer."
SIGLIP = "vision_backbone.fused_featurizer."
PROJ = "projector."
LM = "language_model."

def param_shapes(cfg: OpenVLAConfig, all_blocks: bool = False) -> dict:
    """name -> shape for every parameter the hot path touches (unused checkpoint entries such as the towers'
    final ``norm``, SigLIP ``attn_pool`` and the last block of each tower are omitted; ``all_blocks`` adds the last
    block of each tower, which timm executes although its output is discarded)."""
    shapes = {}
    for prefix, v in ((DINO, cfg.dino), (SIGLIP, cfg.siglip)):
        shapes[prefix + "patch_embed.proj.weight"] = (v.dim, 3, v.patch, v.patch)
        shapes[prefix + "patch_embed.proj.bias"] = (v.dim,)
        shapes[prefix + "pos_embed"] = (1, v.num_patches, v.dim)
        if v.num_prefix:
            shapes[prefix + "cls_token"] = (1, 1, v.dim)
            shapes[prefix + "reg_token"] = (1, v.num_prefix - 1, v.dim)
        for i in range(v.depth if all_blocks else v.blocks_used):
            p = f"{prefix}blocks.{i}."
            shapes[p + "norm1.weight"] = (v.dim,)
            shapes[p + "norm1.bias"] = (v.dim,)
            shapes[p + "attn.qkv.weight"] = (3 * v.dim, v.dim)
            shapes[p + "attn.qkv.bias"] = (3 * v.dim,)
            shapes[p + "attn.proj.weight"] = (v.dim, v.dim)
            shapes[p + "attn.proj.bias"] = (v.dim,)
            shapes[p + "norm2.weight"] = (v.dim,)
            shapes[p + "norm2.bias"] = (v.dim,)
            shapes[p + "mlp.fc1.weight"] = (v.mlp_hidden, v.dim)
            shapes[p + "mlp.fc1.bias"] = (v.mlp_hidden,)
            shapes[p + "mlp.fc2.weight"] = (v.dim, v.mlp_hidden)
            shapes[p + "mlp.fc2.bias"] = (v.dim,)
            if v.layerscale:
                shapes[p + "ls1.scale_factor"] = (v.dim,)
                shapes[p + "ls2.scale_factor"] = (v.dim,)
    vd, ph, hd = cfg.vision_dim, cfg.proj_hidden, cfg.llm.hidden
    shapes[PROJ + "fc1.weight"] = (ph, vd)
    shapes[PROJ + "fc1.bias"] = (ph,)
    shapes[PROJ + "fc2.weight"] = (hd, ph)
    shapes[PROJ + "fc2.bias"] = (hd,)
    shapes[PROJ + "fc3.weight"] = (hd, hd)
    shapes[PROJ + "fc3.bias"] = (hd,)
    l = cfg.llm
    shapes[LM + "model.embed_tokens.weight"] = (l.vocab, l.hidden)
    for i in range(l.layers):
        p = f"{LM}model.layers.{i}."
        shapes[p + "input_layernorm.weight"] = (l.hidden,)
        for n in ("q", "k", "v", "o"):
            shapes[p + f"self_attn.{n}_proj.weight"] = (l.hidden, l.hidden)
        shapes[p + "post_attention_layernorm.weight"] = (l.hidden,)
        shapes[p + "mlp.gate_proj.weight"] = (l.ffn, l.hidden)
        shapes[p + "mlp.up_proj.weight"] = (l.ffn, l.hidden)
        shapes[p + "mlp.down_proj.weight"] = (l.hidden, l.ffn)
    shapes[LM + "model.norm.weight"] = (l.hidden,)
    shapes[LM + "lm_head.weight"] = (l.vocab, l.hidden)
    return shapes



def random_tensor(name, shape, g, device, init="reference"):
    """One synthetic parameter in fp32 (see ``random_state_dict``)."""
    leaf = name.rsplit(".", 1)[-1]
    is_norm = ("norm" in name.rsplit(".", 2)[-2]) if name.count(".") >= 2 else False
    if init == "reference":
        if "scale_factor" in name:
            return torch.full(shape, 1e-5, device=device)
        if is_norm and leaf == "weight":
            return torch.ones(shape, device=device)
        if leaf == "bias":
            return torch.zeros(shape, device=device)
        return torch.randn(shape, generator=g, device=device) * 0.02
    if "scale_factor" in name:
        return torch.randn(shape, generator=g, device=device) * 0.3
    if is_norm and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g, device=device)
    if name.endswith("embed_tokens.weight") or leaf in ("pos_embed", "cls_token", "reg_token"):
        return torch.randn(shape, generator=g, device=device) * 0.5
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g, device=device) * (1.0 / fan_in ** 0.5)


def random_state_dict(cfg: OpenVLAConfig, seed: int = 0, device="cpu", dtype=torch.bfloat16, init: str = "reference"):
    """Random-init weights of the right shapes (no checkpoint is available offline).

    ``init="reference"``: ``_init_weights`` of the HF port (modeling_prismatic.py:185-205): N(0, 0.02) for
    Linear / Conv / Embedding weights, zero biases, unit norm weights, LayerScale 1e-5.
    ``init="test"``: fan-in scaled weights, non-zero biases / norm offsets and O(0.3) LayerScale so that every
    branch carries signal and gradient in the parity tests.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    return {name: random_tensor(name, shape, g, device, init).to(dtype) for name, shape in param_shapes(cfg).items()}


def resolve_vla(src):
    """The reference's ``vla_path`` (UADA_ddp.py:37-50, *_wrapper.py) -> (state dict, engine configuration or None).  A file is a
    ``torch.save``d state dict; anything else (a hub id such as ``openvla/openvla-7b`` or a checkpoint directory) goes through
    ``AutoModelForVision2Seq.from_pretrained`` exactly as the reference loads it (bf16, low_cpu_mem_usage, trust_remote_code),
    which needs the reference's own environment (transformers 4.40 + the prismatic HF classes)."""
    import os
    if isinstance(src, (str, os.PathLike)) and os.path.isfile(src):
        return torch.load(src, map_location="cpu", weights_only=True, mmap=True), None
    try:
        from transformers import AutoModelForVision2Seq
    except ImportError as e:
        raise RuntimeError(f"cannot load '{src}': this transformers has no AutoModelForVision2Seq (the reference pins 4.40.1); "
                           "pass a torch.save'd state dict, a state dict or a loaded model instead") from e
    try:   # the registrations the reference performs before from_pretrained (UADA_ddp.py:38-40)
        from transformers import AutoConfig, AutoProcessor
        from prismatic.extern.hf.configuration_prismatic import OpenVLAConfig as HFOpenVLAConfig
        from prismatic.extern.hf.modeling_prismatic import OpenVLAForActionPrediction
        from prismatic.extern.hf.processing_prismatic import PrismaticProcessor
        AutoConfig.register("openvla", HFOpenVLAConfig)
        AutoProcessor.register(HFOpenVLAConfig, PrismaticProcessor)
        AutoModelForVision2Seq.register(HFOpenVLAConfig, OpenVLAForActionPrediction)
    except Exception:   # noqa: BLE001 -- already registered, or remote code is used
        pass
    vla = AutoModelForVision2Seq.from_pretrained(src, torch_dtype=torch.bfloat16, low_cpu_mem_usage=True, trust_remote_code=True)
    cfg = getattr(vla, "cfg", None)
    if cfg is None and hasattr(getattr(vla, "config", None), "text_config"):
        from .config import config_from_hf
        cfg = config_from_hf(vla.config)
    return vla.state_dict(), cfg
