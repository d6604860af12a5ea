"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden.py

What is executed from /root/reference (never copied into this repo):
  * VLAAttacker/white_patch/appply_random_transform.py -- loaded as text, the 3-space indent of line 43 fixed in
    memory (the file is an IndentationError as shipped), then exec'd: RandomPatchTransform
    .apply_random_patch_batch / .paste_patch_fix / .random_paste_patch / .im_process, plus autograd to the patch.
  * VLAAttacker/white_patch/{UADA,UADA_ddp,UPA}.py -- imported with their unavailable plotting / logging /
    dataset imports stubbed: mask_labels, weighted_loss, cal_UAD, calculate_relative_distance.
  * prismatic/vla/action_tokenizer.py -- ActionTokenizer encode / decode tables.
Third-party pins (installed versions, not the reference's pinned ones): transformers LlamaForCausalLM (tiny
config, fp32) and get_cosine_schedule_with_warmup.
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def load_reference_frontend():
    path = os.path.join(REF, "VLAAttacker/white_patch/appply_random_transform.py")
    src = open(path).read()
    src = src.replace("\n   def simulation_random_patch", "\n    def simulation_random_patch")
    mod = types.ModuleType("appply_random_transform")
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_file_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Stub(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Stub(self.__name__ + "." + item)

    def __call__(self, *a, **k):
        return None


def stub_modules(names):
    for n in names:
        parts = n.split(".")
        for i in range(1, len(parts) + 1):
            key = ".".join(parts[:i])
            if key not in sys.modules:
                sys.modules[key] = _Stub(key)


def load_attack_module(fname, frontend_mod, action_tok_mod):
    stub_modules(["seaborn", "matplotlib", "matplotlib.pyplot", "wandb", "prismatic", "prismatic.vla",
                  "prismatic.models", "prismatic.models.backbones", "prismatic.models.backbones.llm",
                  "prismatic.models.backbones.llm.prompting", "prismatic.extern", "prismatic.extern.hf",
                  "prismatic.extern.hf.configuration_prismatic", "prismatic.extern.hf.modeling_prismatic",
                  "prismatic.extern.hf.processing_prismatic", "prismatic.util", "prismatic.util.data_utils",
                  "white_patch", "white_patch.openvla_dataloader", "openvla_dataloader", "tensorflow", "dlimp",
                  "draccus", "accelerate", "timm"])
    sys.modules["prismatic.vla.action_tokenizer"] = action_tok_mod
    sys.modules["white_patch.appply_random_transform"] = frontend_mod
    sys.modules["appply_random_transform"] = frontend_mod
    import transformers
    # names removed from the installed transformers 5.5 that the modules import at top level
    for missing in ("AdamW", "AutoModelForVision2Seq"):
        if not hasattr(transformers, missing):
            setattr(transformers, missing, object)
    return load_file_module("ref_" + fname.replace(".py", ""), "VLAAttacker/white_patch/" + fname)


class FakeTokenizer:
    vocab_size = 32000


def synth_labels(B, T, rng):
    """Collator layout (prismatic/vla/datasets/datasets.py:56-69): -100 except the last 8 = 7 action ids + EOS."""
    labels = torch.full((B, T), -100, dtype=torch.int64)
    labels[:, -8:-1] = torch.from_numpy(rng.integers(31744, 32000, size=(B, 7)))
    labels[:, -1] = 2
    return labels


def main():
    from PIL import Image
    out = {}
    fe = load_reference_frontend()
    atok = load_file_module("ref_action_tokenizer", "prismatic/vla/action_tokenizer.py")

    # ---------------- front end ----------------
    mean = [torch.tensor([0.484375, 0.455078125, 0.40625]), torch.tensor([0.5, 0.5, 0.5])]
    std = [torch.tensor([0.228515625, 0.2236328125, 0.224609375]), torch.tensor([0.5, 0.5, 0.5])]
    for tag, (S, p, B) in {"s64": (64, 16, 3), "s224": (224, 50, 2)}.items():
        rng = np.random.default_rng(1234)
        obs = rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8)
        images = [Image.fromarray(o) for o in obs]
        torch.manual_seed(42)
        patch = torch.rand(3, p, p)
        t = fe.RandomPatchTransform("cpu", False)
        out[f"fe_{tag}_obs"] = obs
        out[f"fe_{tag}_patch"] = patch.numpy()
        for mode, fn in (("warp", lambda pt: t.apply_random_patch_batch(images, pt, mean, std, True)),
                         ("paste20", lambda pt: t.apply_random_patch_batch(images, pt, mean, std, False)),
                         ("fix", lambda pt: t.paste_patch_fix(images, pt, mean, std)),
                         ("rpaste", lambda pt: t.random_paste_patch(images, pt, mean, std))):
            random.seed(42)
            np.random.seed(42)
            pt = patch.clone().requires_grad_(True)
            y = fn(pt)
            gw = torch.from_numpy(np.random.default_rng(7).standard_normal(y.shape).astype(np.float32))
            (y * gw).sum().backward()
            if tag == "s64":
                out[f"fe_{tag}_{mode}_out"] = y.detach().numpy()
            else:   # full 224 output is 2.4 MB: keep a strided sample + per-channel sums
                out[f"fe_{tag}_{mode}_out_strided"] = y.detach()[:, :, ::5, ::5].numpy()
                out[f"fe_{tag}_{mode}_out_sum"] = y.detach().double().sum(dim=(2, 3)).numpy()
            out[f"fe_{tag}_{mode}_grad"] = pt.grad.numpy()
        y = t.im_process(images, mean, std)
        out[f"fe_{tag}_none_out_sum"] = y.double().sum(dim=(2, 3)).numpy()

    # ---------------- action tokenizer ----------------
    tok = atok.ActionTokenizer(FakeTokenizer())
    ids = np.arange(31744, 32000)
    out["tok_ids"] = ids
    out["tok_decode"] = tok.decode_token_ids_to_actions(ids)
    acts = np.linspace(-1.2, 1.2, 97)
    out["tok_actions"] = acts
    out["tok_encode_ids"] = 32000 - np.digitize(np.clip(acts, -1.0, 1.0), tok.bins)   # body of __call__ before decode
    out["tok_begin_idx"] = np.array(tok.action_token_begin_idx)

    # ---------------- loss heads ----------------
    uada = load_attack_module("UADA.py", fe, atok)
    ddp = load_attack_module("UADA_ddp.py", fe, atok)
    upa = load_attack_module("UPA.py", fe, atok)
    rng = np.random.default_rng(99)
    B, T, P = 4, 14, 256
    L = T + P
    labels = synth_labels(B, T, rng)
    zslice = torch.from_numpy((rng.standard_normal((B, T - 1, 256)) * 3).astype(np.float32))
    logits = torch.zeros(B, L, 32064)
    logits[:, P:L - 1, 31744:32000] = zslice
    out["loss_labels"] = labels.numpy()
    out["loss_zslice"] = zslice.numpy()

    a = uada.OpenVLAAttacker.__new__(uada.OpenVLAAttacker)
    a.action_tokenizer = tok
    for mi, maskidx in enumerate(([0], [0, 1, 2], [6], [0, 1, 2, 3, 4, 5, 6])):
        ml = a.mask_labels(labels.clone(), maskidx)
        out[f"uada_mask{mi}_idx"] = np.array(maskidx)
        out[f"uada_mask{mi}_labels"] = ml.numpy()
        lg = logits.clone().requires_grad_(True)
        loss, uad = a.weighted_loss(lg, ml, maskidx)
        loss.backward()
        out[f"uada_mask{mi}_loss"] = loss.detach().numpy()
        out[f"uada_mask{mi}_uad"] = np.array(float(uad))
        out[f"uada_mask{mi}_dz"] = lg.grad[:, P:L - 1, 31744:32000].numpy()
        d = ddp.OpenVLAAttacker.__new__(ddp.OpenVLAAttacker)
        d.action_tokenizer = tok
        ml2 = d.mask_labels(labels.clone(), maskidx)
        assert torch.equal(ml, ml2)
        for w in (1, 5):
            loss2, uad2 = d.weighted_loss(logits, ml2, "cpu", w)
            out[f"ddp_mask{mi}_w{w}_loss"] = loss2.numpy()
            out[f"ddp_mask{mi}_w{w}_uad"] = np.array(float(uad2))
        # relative distance (UADA.py:165-178 + :355-369)
        action_logits = logits[:, P:-1]
        preds = action_logits.argmax(dim=2)
        gt = ml[:, 1:]
        m = gt > tok.action_token_begin_idx
        cp = torch.tensor(tok.decode_token_ids_to_actions(preds[m].numpy()))
        cg = torch.tensor(tok.decode_token_ids_to_actions(gt[m].numpy()))
        rd = a.calculate_relative_distance(cp, cg, maskidx, {f"{i}": [] for i in maskidx})
        out[f"uada_mask{mi}_rd"] = np.array([rd[str(i)] for i in maskidx])

    u = upa.OpenVLAAttacker.__new__(upa.OpenVLAAttacker)
    u.action_tokenizer = tok
    u.alpha, u.belta = 0.8, 0.2
    u.vla = types.SimpleNamespace(vision_backbone=types.SimpleNamespace(
        featurizer=types.SimpleNamespace(patch_embed=types.SimpleNamespace(num_patches=P))))
    lg = logits.clone().requires_grad_(True)
    total, ang, dist = u.weighted_loss(lg, labels.clone())
    total.backward()
    out["upa_loss"] = total.detach().numpy()
    out["upa_angle"] = np.array(ang)
    out["upa_dist"] = np.array(dist)
    out["upa_dz"] = lg.grad[:, P:L - 1, 31744:32000].numpy()
    for mi, maskidx in enumerate(([0], [2, 4], [0, 1, 2, 3, 4, 5, 6])):
        out[f"upa_mask{mi}_idx"] = np.array(maskidx)
        out[f"upa_mask{mi}_labels"] = u.mask_labels(labels.clone(), maskidx).numpy()

    # ---------------- cosine schedule (installed transformers) ----------------
    import transformers
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=2e-3)
    sch = transformers.get_cosine_schedule_with_warmup(opt, num_warmup_steps=20, num_training_steps=2000,
                                                       num_cycles=0.5, last_epoch=-1)
    lrs = []
    for _ in range(2000):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    out["sched_lrs"] = np.array(lrs)

    # ---------------- HF Llama (installed transformers, tiny config, fp32) ----------------
    from transformers import LlamaConfig as HFLlamaConfig, LlamaForCausalLM
    from roboticattack_b200.config import tiny
    from roboticattack_b200.weights import LM, random_state_dict
    cfg = tiny()
    sd = random_state_dict(cfg, seed=3, dtype=torch.float32, init="test")
    hf_cfg = HFLlamaConfig(vocab_size=cfg.llm.vocab, hidden_size=cfg.llm.hidden, intermediate_size=cfg.llm.ffn,
                           num_hidden_layers=cfg.llm.layers, num_attention_heads=cfg.llm.heads,
                           num_key_value_heads=cfg.llm.heads, rms_norm_eps=cfg.llm.rms_eps, rope_theta=cfg.llm.rope_theta,
                           pad_token_id=32000, attention_bias=False, tie_word_embeddings=False)
    hf = LlamaForCausalLM(hf_cfg).eval()
    missing = hf.load_state_dict({k[len(LM):]: v for k, v in sd.items() if k.startswith(LM)}, strict=False)
    assert not missing.unexpected_keys, missing
    assert all("rotary" in k for k in missing.missing_keys), missing
    g = torch.Generator().manual_seed(11)
    Bl, Ll = 3, 40
    emb = torch.randn(Bl, Ll, cfg.llm.hidden, generator=g) * 0.5
    mask = torch.ones(Bl, Ll, dtype=torch.bool)
    mask[1, 33:] = False
    mask[2, 37:] = False
    lab = torch.full((Bl, Ll), -100, dtype=torch.int64)
    for b, n in enumerate((40, 33, 37)):
        lab[b, n - 8:n - 1] = torch.randint(31744, 32000, (7,), generator=g)
        lab[b, n - 1] = 2
    with torch.no_grad():
        o = hf(inputs_embeds=emb, attention_mask=mask, labels=lab)
    out["llama_emb"] = emb.numpy()
    out["llama_mask"] = mask.numpy()
    out["llama_labels"] = lab.numpy()
    out["llama_loss"] = o.loss.numpy()
    sup = lab[:, 1:] != -100
    out["llama_sup_logits_action"] = o.logits[:, :-1][sup][:, 31744:32000].float().numpy()

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    total = sum(v.nbytes for v in out.values())
    print(f"wrote {len(out)} arrays, {total/1e6:.2f} MB raw ->", os.path.join(HERE, "reference_golden.npz"),
          os.path.getsize(os.path.join(HERE, "reference_golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
