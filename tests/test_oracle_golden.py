"""Pin the oracle against outputs of the reference's own functions (tests/golden/make_golden.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from oracle import llama as ollama
from oracle import losses as ol
from oracle import optim as oo
from roboticattack_b200.config import NORM_MEAN, NORM_STD, tiny
from roboticattack_b200.weights import LM, random_state_dict

MODES = {"warp": (ofe.MODE_WARP, True), "paste20": (ofe.MODE_PASTE20, False), "fix": (ofe.MODE_FIX, False),
         "rpaste": (ofe.MODE_FIX, False)}


@pytest.mark.parametrize("tag", ["s64", "s224"])
@pytest.mark.parametrize("mode", list(MODES))
def test_frontend_matches_reference(golden, tag, mode):
    obs = torch.from_numpy(golden[f"fe_{tag}_obs"])
    patch = torch.from_numpy(golden[f"fe_{tag}_patch"]).clone().requires_grad_(True)
    B, S = obs.shape[0], obs.shape[1]
    m, geometry = MODES[mode]
    random.seed(42)
    np.random.seed(42)
    xy, theta = ofe.draw_placements(B, (S, S), patch.shape[1:], geometry)
    y = ofe.apply_patch_batch(obs, patch, xy, theta, m, NORM_MEAN, NORM_STD)
    gw = torch.from_numpy(np.random.default_rng(7).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * gw).sum().backward()
    if tag == "s64":
        np.testing.assert_allclose(y.detach().numpy(), golden[f"fe_{tag}_{mode}_out"], rtol=0, atol=1e-6)
    else:
        np.testing.assert_allclose(y.detach()[:, :, ::5, ::5].numpy(), golden[f"fe_{tag}_{mode}_out_strided"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(y.detach().double().sum(dim=(2, 3)).numpy(), golden[f"fe_{tag}_{mode}_out_sum"], rtol=1e-9)
    np.testing.assert_allclose(patch.grad.numpy(), golden[f"fe_{tag}_{mode}_grad"], rtol=1e-5, atol=1e-5)


def test_frontend_no_patch(golden):
    obs = torch.from_numpy(golden["fe_s64_obs"])
    y = ofe.apply_patch_batch(obs, torch.zeros(3, 4, 4), None, None, ofe.MODE_NONE, NORM_MEAN, NORM_STD)
    np.testing.assert_allclose(y.double().sum(dim=(2, 3)).numpy(), golden["fe_s64_none_out_sum"], rtol=1e-9)


def test_action_tokenizer(golden):
    np.testing.assert_array_equal(ol.decode_token_ids_to_actions(golden["tok_ids"]), golden["tok_decode"])
    np.testing.assert_array_equal(ol.encode_actions_to_token_ids(golden["tok_actions"]), golden["tok_encode_ids"])
    from roboticattack_b200.config import ACTION_TOKEN_BEGIN_IDX, ZERO_ACTION_TOKEN
    assert ACTION_TOKEN_BEGIN_IDX == int(golden["tok_begin_idx"])
    assert ol.encode_actions_to_token_ids(np.zeros(1))[0] == ZERO_ACTION_TOKEN


def _full_logits(golden):
    z = torch.from_numpy(golden["loss_zslice"])
    B, Tm1, _ = z.shape
    L = 256 + Tm1 + 1
    logits = torch.zeros(B, L, 32064)
    logits[:, 256:L - 1, 31744:32000] = z
    return logits, L


@pytest.mark.parametrize("mi", range(4))
def test_uada_heads(golden, mi):
    labels = torch.from_numpy(golden["loss_labels"]).clone()
    maskidx = golden[f"uada_mask{mi}_idx"].tolist()
    ml = ol.mask_labels_uada(labels, maskidx)
    np.testing.assert_array_equal(ml.numpy(), golden[f"uada_mask{mi}_labels"])
    logits, L = _full_logits(golden)
    lg = logits.clone().requires_grad_(True)
    loss, uad = ol.weighted_loss_uada(lg, ml, 5)
    loss.backward()
    np.testing.assert_allclose(loss.item(), golden[f"uada_mask{mi}_loss"], rtol=1e-6)
    np.testing.assert_allclose(float(uad), golden[f"uada_mask{mi}_uad"], rtol=1e-6)
    np.testing.assert_allclose(lg.grad[:, 256:L - 1, 31744:32000].numpy(), golden[f"uada_mask{mi}_dz"], rtol=1e-5, atol=1e-8)
    for w in (1, 5):
        l2, u2 = ol.weighted_loss_uada(logits, ml, w)
        np.testing.assert_allclose(l2.item(), golden[f"ddp_mask{mi}_w{w}_loss"], rtol=1e-6)
        np.testing.assert_allclose(float(u2), golden[f"ddp_mask{mi}_w{w}_uad"], rtol=1e-6)
    # relative distance per (sample, dof)
    preds = logits[:, 256:-1].argmax(dim=2)
    gt = ml[:, 1:]
    m = gt > 31743
    cp = torch.tensor(ol.decode_token_ids_to_actions(preds[m].numpy())).view(-1, len(maskidx))
    cg = torch.tensor(ol.decode_token_ids_to_actions(gt[m].numpy())).view(-1, len(maskidx))
    rd = ol.relative_distance(cp, cg).t().numpy()
    np.testing.assert_allclose(rd, golden[f"uada_mask{mi}_rd"], rtol=1e-9)


def test_upa_heads(golden):
    labels = torch.from_numpy(golden["loss_labels"]).clone()
    logits, L = _full_logits(golden)
    lg = logits.clone().requires_grad_(True)
    total, ang, dist = ol.weighted_loss_upa(lg, labels, 0.8, 0.2, 256)
    total.backward()
    np.testing.assert_allclose(total.item(), golden["upa_loss"], rtol=1e-6)
    np.testing.assert_allclose(ang.item(), golden["upa_angle"], rtol=1e-6)
    np.testing.assert_allclose(dist.item(), golden["upa_dist"], rtol=1e-6)
    np.testing.assert_allclose(lg.grad[:, 256:L - 1, 31744:32000].numpy(), golden["upa_dz"], rtol=1e-5, atol=1e-9)
    for mi in range(3):
        ml = ol.mask_labels_upa(torch.from_numpy(golden["loss_labels"]).clone(), golden[f"upa_mask{mi}_idx"].tolist())
        np.testing.assert_array_equal(ml.numpy(), golden[f"upa_mask{mi}_labels"])


def test_cosine_schedule(golden):
    lrs = np.array([2e-3 * oo.cosine_with_warmup_lambda(s, 20, 2000) for s in range(2000)])
    np.testing.assert_allclose(lrs, golden["sched_lrs"], rtol=1e-12, atol=1e-18)
    assert lrs[0] == 0.0   # lr is 0 for the whole of outer iteration 0


def test_llama_matches_hf(golden):
    cfg = tiny()
    sd = random_state_dict(cfg, seed=3, dtype=torch.float32, init="test")
    emb = torch.from_numpy(golden["llama_emb"])
    mask = torch.from_numpy(golden["llama_mask"])
    lab = torch.from_numpy(golden["llama_labels"])
    loss, logits = ollama.llama_forward(sd, LM, cfg.llm, emb, mask, lab)
    np.testing.assert_allclose(loss.item(), golden["llama_loss"], rtol=2e-5)
    sup = lab[:, 1:] != -100
    got = logits[:, :-1][sup][:, 31744:32000].numpy()
    np.testing.assert_allclose(got, golden["llama_sup_logits_action"], rtol=1e-3, atol=2e-4)


def test_adamw_rule_matches_definition():
    """eps is added before the bias correction (transformers.AdamW), unlike torch.optim.AdamW."""
    torch.manual_seed(0)
    p = torch.rand(3, 5, 5)
    opt = oo.HFAdamW(p.shape, lr=2e-3)
    p0 = p.clone()
    g = torch.randn_like(p)
    opt.step(p, g)
    # first step: m = 0.1 g, v = 0.001 g^2, step = lr*sqrt(0.001)/0.1
    expect = p0 - 2e-3 * (0.001 ** 0.5) / 0.1 * (0.1 * g) / ((0.001 * g * g).sqrt() + 1e-6)
    torch.testing.assert_close(p, expect, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------ eval-time paste
SIM_CASES = ["s64_plain", "s64_geo", "s224_plain", "s224_geo", "s224_geo_default"]


@pytest.mark.parametrize("tag", SIM_CASES)
def test_simulation_paste_matches_reference(tag):
    """oracle.frontend.simulation_paste == the reference's own simulation_random_patch (tests/golden/make_golden_sim.py)."""
    import os
    from oracle import frontend as ofe
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_sim.npz"))
    geo, angle, shx, shy, x, y = g[f"{tag}_args"]
    out = ofe.simulation_paste(g[f"{tag}_img"], torch.from_numpy(g[f"{tag}_patch"]), bool(geo), angle, shx, shy, (int(x), int(y)))
    assert out.dtype == np.uint8 and out.shape == g[f"{tag}_out"].shape
    assert np.array_equal(out, g[f"{tag}_out"])
    assert (out != g[f"{tag}_img"]).any(), "the patch must be visible"


def _rand_vit_sd(c, prefix, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.2   # noqa: E731
    d, m = c.dim, c.mlp_hidden
    sd = {prefix + "patch_embed.proj.weight": r(d, 3, c.patch, c.patch), prefix + "patch_embed.proj.bias": r(d),
          prefix + "pos_embed": r(1, c.num_patches, d)}
    if c.num_prefix:
        sd[prefix + "cls_token"] = r(1, 1, d)
        sd[prefix + "reg_token"] = r(1, c.num_prefix - 1, d)
    for i in range(c.depth):
        p = f"{prefix}blocks.{i}."
        sd.update({p + "norm1.weight": 1 + r(d), p + "norm1.bias": r(d), p + "attn.qkv.weight": r(3 * d, d), p + "attn.qkv.bias": r(3 * d),
                   p + "attn.proj.weight": r(d, d), p + "attn.proj.bias": r(d), p + "norm2.weight": 1 + r(d), p + "norm2.bias": r(d),
                   p + "mlp.fc1.weight": r(m, d), p + "mlp.fc1.bias": r(m), p + "mlp.fc2.weight": r(d, m), p + "mlp.fc2.bias": r(d)})
        if c.layerscale:
            sd.update({p + "ls1.scale_factor": r(d), p + "ls2.scale_factor": r(d)})
    return sd


def test_vit_restatement_matches_hf_dinov2_with_registers():
    """timm is absent from the container, so ``oracle/vit.py`` restates its ``VisionTransformer`` from the published algorithm.
    Cross-check of the DINOv2-reg4 wiring (pre-LN blocks, fused q|k|v, LayerScale, eps 1e-6, [cls, reg x4, patches], cls
    without position embedding, second-to-last block output) against an INDEPENDENT implementation of the same model
    family: ``transformers.Dinov2WithRegistersModel`` with the same random weights."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    from oracle import vit as ovit
    from roboticattack_b200.config import ViTConfig
    c = ViTConfig(dim=64, depth=4, heads=2, mlp_hidden=256, num_prefix=5, layerscale=True, img=28, patch=14)
    sd = _rand_vit_sd(c, "f.", 11)
    hf = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(hidden_size=64, num_hidden_layers=4, num_attention_heads=2, mlp_ratio=4,
                                                            image_size=28, patch_size=14, num_register_tokens=4, layer_norm_eps=1e-6,
                                                            hidden_act="gelu", attn_implementation="eager")).eval()
    t = hf.state_dict()
    t["embeddings.patch_embeddings.projection.weight"] = sd["f.patch_embed.proj.weight"]
    t["embeddings.patch_embeddings.projection.bias"] = sd["f.patch_embed.proj.bias"]
    t["embeddings.cls_token"] = sd["f.cls_token"]
    t["embeddings.register_tokens"] = sd["f.reg_token"]
    t["embeddings.position_embeddings"] = torch.cat([torch.zeros(1, 1, 64), sd["f.pos_embed"]], dim=1)   # no_embed_class=True
    for i in range(c.depth):
        p, h = f"f.blocks.{i}.", f"encoder.layer.{i}."
        qw, kw, vw = sd[p + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = sd[p + "attn.qkv.bias"].chunk(3, 0)
        t.update({h + "norm1.weight": sd[p + "norm1.weight"], h + "norm1.bias": sd[p + "norm1.bias"],
                  h + "attention.attention.query.weight": qw, h + "attention.attention.query.bias": qb,
                  h + "attention.attention.key.weight": kw, h + "attention.attention.key.bias": kb,
                  h + "attention.attention.value.weight": vw, h + "attention.attention.value.bias": vb,
                  h + "attention.output.dense.weight": sd[p + "attn.proj.weight"], h + "attention.output.dense.bias": sd[p + "attn.proj.bias"],
                  h + "layer_scale1.lambda1": sd[p + "ls1.scale_factor"], h + "layer_scale2.lambda1": sd[p + "ls2.scale_factor"],
                  h + "norm2.weight": sd[p + "norm2.weight"], h + "norm2.bias": sd[p + "norm2.bias"],
                  h + "mlp.fc1.weight": sd[p + "mlp.fc1.weight"], h + "mlp.fc1.bias": sd[p + "mlp.fc1.bias"],
                  h + "mlp.fc2.weight": sd[p + "mlp.fc2.weight"], h + "mlp.fc2.bias": sd[p + "mlp.fc2.bias"]})
    missing = hf.load_state_dict(t, strict=True)
    img = torch.randn(2, 3, 28, 28, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = hf(pixel_values=img, output_hidden_states=True).hidden_states[c.depth - 1][:, c.num_prefix:]
        got = ovit.vit_forward(sd, "f.", c, img)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_vit_restatement_matches_hf_siglip():
    """Same cross-check for the SigLIP tower (no cls token, no LayerScale, learned position embedding on every token):
    ``transformers.SiglipVisionModel`` with ``hidden_act="gelu"`` -- timm 0.9.10's model definition passes no activation, i.e.
    the erf GELU that ``oracle/vit.py`` (and the engine) use."""
    from transformers import SiglipVisionConfig, SiglipVisionModel
    from oracle import vit as ovit
    from roboticattack_b200.config import ViTConfig
    c = ViTConfig(dim=72, depth=4, heads=2, mlp_hidden=264, num_prefix=0, layerscale=False, img=28, patch=14)
    sd = _rand_vit_sd(c, "s.", 12)
    hf = SiglipVisionModel(SiglipVisionConfig(hidden_size=72, intermediate_size=264, num_hidden_layers=4, num_attention_heads=2,
                                              image_size=28, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu",
                                              attn_implementation="eager")).eval()
    t = hf.state_dict()
    t["vision_model.embeddings.patch_embedding.weight"] = sd["s.patch_embed.proj.weight"]
    t["vision_model.embeddings.patch_embedding.bias"] = sd["s.patch_embed.proj.bias"]
    t["vision_model.embeddings.position_embedding.weight"] = sd["s.pos_embed"][0]
    for i in range(c.depth):
        p, h = f"s.blocks.{i}.", f"vision_model.encoder.layers.{i}."
        qw, kw, vw = sd[p + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = sd[p + "attn.qkv.bias"].chunk(3, 0)
        t.update({h + "layer_norm1.weight": sd[p + "norm1.weight"], h + "layer_norm1.bias": sd[p + "norm1.bias"],
                  h + "self_attn.q_proj.weight": qw, h + "self_attn.q_proj.bias": qb, h + "self_attn.k_proj.weight": kw,
                  h + "self_attn.k_proj.bias": kb, h + "self_attn.v_proj.weight": vw, h + "self_attn.v_proj.bias": vb,
                  h + "self_attn.out_proj.weight": sd[p + "attn.proj.weight"], h + "self_attn.out_proj.bias": sd[p + "attn.proj.bias"],
                  h + "layer_norm2.weight": sd[p + "norm2.weight"], h + "layer_norm2.bias": sd[p + "norm2.bias"],
                  h + "mlp.fc1.weight": sd[p + "mlp.fc1.weight"], h + "mlp.fc1.bias": sd[p + "mlp.fc1.bias"],
                  h + "mlp.fc2.weight": sd[p + "mlp.fc2.weight"], h + "mlp.fc2.bias": sd[p + "mlp.fc2.bias"]})
    hf.load_state_dict(t, strict=True)
    img = torch.randn(2, 3, 28, 28, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = hf(pixel_values=img, output_hidden_states=True).hidden_states[c.depth - 1]
        got = ovit.vit_forward(sd, "s.", c, img)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_multimodal_forward_matches_reference_model_class():
    """``oracle.model.forward`` against the REFERENCE's own ``OpenVLAForActionPrediction.forward`` executed on the CPU
    (tests/golden/make_golden_glue.py -> reference_golden_glue.npz): channel split and tower order, second-to-last-block
    features, the LayerScale ``scale_factor`` patch, projector, [BOS | patches | text] splice of embeddings / mask / labels,
    HF Llama with padding, fp32 logits and the shifted cross-entropy."""
    import os
    from oracle import model as om
    from roboticattack_b200.config import LlamaConfig, OpenVLAConfig, ViTConfig
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_glue.npz"))
    sd = {k[2:]: torch.from_numpy(gold[k].astype(np.float32)) for k in gold.files if k.startswith("w:")}
    cfg = OpenVLAConfig(dino=ViTConfig(dim=32, depth=3, heads=2, mlp_hidden=128, num_prefix=5, layerscale=True, img=28),
                        siglip=ViTConfig(dim=40, depth=4, heads=2, mlp_hidden=136, num_prefix=0, layerscale=False, img=28),
                        llm=LlamaConfig(hidden=64, layers=2, heads=2, ffn=176, vocab=384), name="glue")
    ids, mask = torch.from_numpy(gold["input_ids"]), torch.from_numpy(gold["attention_mask"])
    labels, px = torch.from_numpy(gold["labels"]), torch.from_numpy(gold["pixel_values"])
    feats = om.vision_backbone(sd, cfg, px)
    torch.testing.assert_close(feats, torch.from_numpy(gold["vision_features"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(om.projector(sd, feats), torch.from_numpy(gold["projected"]), rtol=1e-5, atol=1e-5)
    out = om.forward(sd, cfg, ids, mask, px, labels)
    np.testing.assert_allclose(out.loss.item(), float(gold["loss"]), rtol=1e-6)
    # rows at padded text positions depend on how the attention implementation treats fully padded keys; every position
    # that can carry a label or be attended to is compared
    P = cfg.num_patches
    mm_mask = torch.cat([mask[:, :1], torch.ones(mask.shape[0], P, dtype=torch.bool), mask[:, 1:]], dim=1)
    got, ref = out.logits.float(), torch.from_numpy(gold["logits"])
    torch.testing.assert_close(got[mm_mask], ref[mm_mask], rtol=1e-4, atol=1e-4)


def test_checkpoint_names_match_the_reference_model_state_dict():
    """The engine loads weights by HF checkpoint name (``vla_engine_load_weight``).  ``param_shapes`` against the state dict of
    the reference's own ``OpenVLAForActionPrediction`` (fixture of make_golden_glue.py): every name and shape the engine asks
    for exists there, and the only reference tensors it ignores are the last block of each tower, whose output the reference
    computes and discards (``get_intermediate_layers(n={depth-2})``)."""
    import os
    from roboticattack_b200.config import LlamaConfig, OpenVLAConfig, ViTConfig
    from roboticattack_b200.weights import param_shapes
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_glue.npz"))
    ref = {k[2:]: tuple(gold[k].shape) for k in gold.files if k.startswith("w:")}
    cfg = OpenVLAConfig(dino=ViTConfig(dim=32, depth=3, heads=2, mlp_hidden=128, num_prefix=5, layerscale=True, img=28),
                        siglip=ViTConfig(dim=40, depth=4, heads=2, mlp_hidden=136, num_prefix=0, layerscale=False, img=28),
                        llm=LlamaConfig(hidden=64, layers=2, heads=2, ffn=176, vocab=384), name="glue")
    ours = {k: tuple(v) for k, v in dict(param_shapes(cfg)).items()}
    assert set(ours) <= set(ref), sorted(set(ours) - set(ref))
    assert all(ours[k] == ref[k] for k in ours), [(k, ours[k], ref[k]) for k in ours if ours[k] != ref[k]]
    ignored = set(ref) - set(ours)
    assert ignored and all(k.startswith(("vision_backbone.featurizer.blocks.2.", "vision_backbone.fused_featurizer.blocks.3.")) for k in ignored)
