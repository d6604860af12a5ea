"""Golden vectors of the multimodal forward, produced by the REFERENCE's own ``OpenVLAForActionPrediction.forward``
(prismatic/extern/hf/modeling_prismatic.py:362-415 with ``PrismaticVisionBackbone`` :63-123, ``PrismaticProjector`` :127-158 and the
LayerScale patch :52-59), imported from /root/reference and executed on the CPU with a toy configuration.

The reference module needs ``timm`` (absent here) at import time, so a stand-in ``timm`` module is registered whose
``create_model`` returns a ViT with timm's module / parameter names (the arithmetic of its blocks is cross-checked against
transformers' DINOv2-with-registers and SigLIP models in tests/test_oracle_golden.py); everything the golden pins is the
reference's own code: the 3+3 channel split and the order of the two towers, ``get_intermediate_layers(n={depth-2})`` through
``unpack_tuple``, the ``gamma -> scale_factor`` LayerScale patch, the projector, the [BOS | patches | text] splice of embeddings,
attention mask and labels, the call into HF Llama and the shifted cross-entropy.  One environment adaptation: transformers 5.x
calls ``tie_weights(recompute_mapping=...)``, which the 4.40-era override does not accept; it is wrapped to drop the argument.

Run in the authoring container:  python tests/golden/make_golden_glue.py
Writes tests/golden/reference_golden_glue.npz (weights, inputs, outputs; nothing from /root/reference is copied)."""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SPECS = {"vit_large_patch14_reg4_dinov2.lvd142m": (32, 3, 2, 128, 5, True),   # dim, depth, heads, mlp, prefix tokens, LayerScale
         "vit_so400m_patch14_siglip_224": (40, 4, 2, 136, 0, False)}
IMG, VOCAB = 28, 384


class LayerScale(nn.Module):   # timm.models.vision_transformer.LayerScale: the class the reference patches
    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.qkv, self.proj = nn.Linear(d, 3 * d), nn.Linear(d, d)


class _Mlp(nn.Module):
    def __init__(self, d, m):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, m), nn.Linear(m, d)


class _Block(nn.Module):
    def __init__(self, d, m, ls):
        super().__init__()
        self.norm1, self.attn, self.norm2, self.mlp = nn.LayerNorm(d, eps=1e-6), _Attn(d), nn.LayerNorm(d, eps=1e-6), _Mlp(d, m)
        self.ls1 = LayerScale(d) if ls else nn.Identity()
        self.ls2 = LayerScale(d) if ls else nn.Identity()


class _PatchEmbed(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.proj = nn.Conv2d(3, d, 14, 14)


class StandInViT(nn.Module):
    """timm ``VisionTransformer`` surface used by the reference: ``blocks``, ``embed_dim``, ``get_intermediate_layers``."""

    def __init__(self, name, img_size):
        super().__init__()
        d, depth, heads, m, npre, ls = SPECS[name]
        self.embed_dim, self.heads, self.npre = d, heads, npre
        self.patch_embed = _PatchEmbed(d)
        self.pos_embed = nn.Parameter(torch.zeros(1, (img_size // 14) ** 2, d))
        if npre:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, d))
            self.reg_token = nn.Parameter(torch.zeros(1, npre - 1, d))
        self.blocks = nn.ModuleList([_Block(d, m, ls) for _ in range(depth)])

    def get_intermediate_layers(self, x, n):
        x = self.patch_embed.proj(x).flatten(2).transpose(1, 2) + self.pos_embed
        if self.npre:
            x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), self.reg_token.expand(x.shape[0], -1, -1), x], 1)
        outs = []
        for i, b in enumerate(self.blocks):
            B, N, C = x.shape
            qkv = b.attn.qkv(b.norm1(x)).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(*qkv.unbind(0)).transpose(1, 2).reshape(B, N, C)
            x = x + b.ls1(b.attn.proj(a))
            x = x + b.ls2(b.mlp.fc2(F.gelu(b.mlp.fc1(b.norm2(x)))))
            if i in n:
                outs.append(x[:, self.npre:])
        return tuple(outs)


def load_reference_model_module():
    timm = types.ModuleType("timm")
    timm.__version__ = "0.9.10"
    timm.create_model = lambda name, pretrained, num_classes, img_size, act_layer: StandInViT(name, img_size)
    tm, tv = types.ModuleType("timm.models"), types.ModuleType("timm.models.vision_transformer")
    tv.LayerScale = LayerScale
    timm.__path__, tm.__path__ = [], []
    for m in (timm, tm, tv):
        m.__spec__ = importlib.machinery.ModuleSpec(m.__name__, None)
        sys.modules[m.__name__] = m
    for pk in ("prismatic", "prismatic.extern", "prismatic.extern.hf"):
        m = types.ModuleType(pk)
        m.__path__ = []
        sys.modules[pk] = m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    cp = load("prismatic.extern.hf.configuration_prismatic", "prismatic/extern/hf/configuration_prismatic.py")
    mp = load("prismatic.extern.hf.modeling_prismatic", "prismatic/extern/hf/modeling_prismatic.py")
    tie = mp.PrismaticForConditionalGeneration.tie_weights
    mp.PrismaticForConditionalGeneration.tie_weights = lambda self, *a, **k: tie(self)
    return cp, mp


def main():
    cp, mp = load_reference_model_module()
    text = dict(hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2, vocab_size=VOCAB,
                max_position_embeddings=512, rms_norm_eps=1e-6, pad_token_id=VOCAB - 1)
    cfg = cp.OpenVLAConfig(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure", image_sizes=[IMG, IMG],
                           text_config=text, attn_implementation="eager")
    torch.manual_seed(0)
    model = mp.OpenVLAForActionPrediction(cfg).eval()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n and n.endswith("weight"):
                v = 1 + 0.2 * torch.randn(p.shape, generator=g)
            elif "scale_factor" in n:
                v = 0.5 * torch.randn(p.shape, generator=g)
            else:
                v = 0.08 * torch.randn(p.shape, generator=g)
            p.copy_(v.half().float())   # fp16-representable values: the fixture stores the weights as float16, exactly
    B, T = 3, 12
    ids = torch.randint(3, VOCAB - 2, (B, T), generator=g)
    ids[:, 0] = 1
    mask = torch.ones(B, T, dtype=torch.bool)
    mask[1, -2:] = False
    mask[2, -5:] = False
    ids[~mask] = VOCAB - 1
    labels = torch.full((B, T), -100)
    labels[0, -8:] = ids[0, -8:]
    labels[1, -10:-2] = ids[1, -10:-2]
    labels[2, 3:7] = ids[2, 3:7]
    px = torch.randn(B, 6, IMG, IMG, generator=g)
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, pixel_values=px, labels=labels)
        feats = model.vision_backbone(px)
        proj = model.projector(feats)
    arrays = {"w:" + k: v.detach().numpy().astype(np.float16) for k, v in model.state_dict().items()}
    assert all(np.array_equal(arrays["w:" + k].astype(np.float32), v.detach().numpy()) for k, v in model.state_dict().items())
    arrays.update(input_ids=ids.numpy(), attention_mask=mask.numpy(), labels=labels.numpy(), pixel_values=px.numpy(),
                  loss=np.float64(out.loss.item()), logits=out.logits.numpy(), vision_features=feats.numpy(), projected=proj.numpy())
    # ---- predict_action (modeling_prismatic.py:506-536): empty-token insertion, de-tokenisation with the full-size vocabulary
    # (32064 - 64), un-normalisation with q01 / q99 / mask.  ``generate`` (a GenerationMixin method the 4.40-era class no
    # longer inherits under transformers 5.x) is replaced by a stand-in that records its input and appends fixed token ids.
    stats = {"toy": {"action": {"q01": [-0.5, -0.2, -1.0, -0.3, -0.1, -0.7, 0.0], "q99": [0.6, 0.9, 1.0, 0.2, 0.4, 0.3, 1.0],
                                "mask": [True, True, True, True, True, True, False]}}}
    text2 = dict(text, vocab_size=32064, pad_token_id=32000)
    cfg2 = cp.OpenVLAConfig(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure", image_sizes=[IMG, IMG],
                            text_config=text2, attn_implementation="eager", norm_stats=stats)
    m2 = mp.OpenVLAForActionPrediction(cfg2).eval()
    cases = np.array([[31872, 31744, 31999, 31745, 31900, 31800, 31744], [31998, 31873, 31871, 31760, 31999, 31744, 31872],
                      [32000, 31743, 31500, 31872, 31872, 31872, 31999]], dtype=np.int64)   # the last row leaves the action range (clip)
    prompts = [torch.tensor([[1, 500, 600, 29871]]), torch.tensor([[1, 500, 600, 700]]), torch.tensor([[1, 29871]])]
    seen, acts = [], []
    for toks, prompt in zip(cases, prompts):
        def fake_generate(input_ids, max_new_tokens, **kw):
            seen.append(input_ids.numpy().copy())
            assert max_new_tokens == 7
            return torch.cat([input_ids, torch.from_numpy(toks)[None]], dim=1)
        m2.generate = fake_generate
        acts.append(m2.predict_action(prompt, unnorm_key="toy"))
    arrays.update(pa_tokens=cases, pa_actions=np.stack(acts), pa_q01=np.array(stats["toy"]["action"]["q01"]),
                  pa_q99=np.array(stats["toy"]["action"]["q99"]), pa_mask=np.array(stats["toy"]["action"]["mask"]))
    for i, (pr, sn) in enumerate(zip(prompts, seen)):
        arrays[f"pa_prompt{i}"] = pr.numpy()
        arrays[f"pa_generate_input{i}"] = sn
    np.savez_compressed(os.path.join(HERE, "reference_golden_glue.npz"), **arrays)
    print("wrote", len(arrays), "arrays; loss", out.loss.item(), "logits", tuple(out.logits.shape))


if __name__ == "__main__":
    main()
