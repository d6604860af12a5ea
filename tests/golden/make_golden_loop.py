"""Golden trajectory of the attack loop, produced by the REFERENCE's own ``OpenVLAAttacker.patchattack_unconstrained``
(VLAAttacker/white_patch/UADA.py:93-292) driving the reference's own model class and front end on the CPU:

  reference UADA.py loop -> reference ``RandomPatchTransform.apply_random_patch_batch`` -> reference
  ``OpenVLAForActionPrediction.forward`` (toy sizes, bf16, stand-in timm ViT as in make_golden_glue.py) -> reference
  ``weighted_loss`` + 1/CE -> backward -> optimizer step -> clamp, with the validation pass of outer iteration 0 in between.

Environment adaptations (none touches the arithmetic being pinned): ``transformers.AdamW`` no longer exists in transformers
5.x, so the name is bound to a torch Optimizer restating its update rule (the one part of the step this golden does NOT pin);
plotting / wandb / dataset modules are stubbed as in make_golden.py.

Run in the authoring container:  python tests/golden/make_golden_loop.py
Writes tests/golden/reference_golden_loop.npz (weights as bf16 bit patterns, batches, per-step losses, patches)."""
import argparse
import math
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import make_golden_glue as glue  # noqa: E402

VOCAB, HIDDEN, IMG, P_HW = 32064, 32, 28, 8
B, T, N_TRAIN, N_VAL, NUM_ITER, INNER, LR = 2, 14, 3, 2, 3, 2, 2e-3
MASKIDX = [0, 1, 2]


def big_matrix(rows, cols, a, b):
    """Embedding / lm_head values by formula (exact in bf16), so that the fixture does not have to store 2 x 32064 x 32 numbers."""
    i = torch.arange(rows)[:, None]
    j = torch.arange(cols)[None, :]
    return (((i * a + j * b) % 257) - 128).float() / 4096.0


class HFAdamW(torch.optim.Optimizer):
    """transformers.AdamW(lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, correct_bias=True) -- restated (removed upstream)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        STEPS.append(self)

    @torch.no_grad()
    def step(self):
        for gr in self.param_groups:
            for p in gr["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"], st["m"], st["v"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                st["step"] += 1
                b1, b2 = gr["betas"]
                st["m"].mul_(b1).add_(p.grad, alpha=1 - b1)
                st["v"].mul_(b2).addcmul_(p.grad, p.grad, value=1 - b2)
                step_size = gr["lr"] * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
                p.addcdiv_(st["m"], st["v"].sqrt().add_(gr["eps"]), value=-step_size)
                GRADS.append(p.grad.detach().clone())
                PATCHES.append(p.detach().clone())


STEPS, GRADS, PATCHES, FINAL = [], [], [], []


def batches(n, seed):
    from PIL import Image
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        obs = rng.integers(0, 256, size=(B, IMG, IMG, 3), dtype=np.uint8)
        ids = torch.from_numpy(rng.integers(3, 31000, size=(B, T)))
        ids[:, 0] = 1
        labels = torch.full((B, T), -100, dtype=torch.int64)
        mask = torch.ones(B, T, dtype=torch.bool)
        for b in range(B):
            n_tok = T - (b + k) % 3          # ragged: right padding with 32000 / -100 / False
            ids[b, n_tok - 8:n_tok - 1] = torch.from_numpy(rng.integers(31744, 32000, size=7))
            ids[b, n_tok - 1] = 2
            labels[b, n_tok - 8:n_tok] = ids[b, n_tok - 8:n_tok]
            ids[b, n_tok:] = 32000
            mask[b, n_tok:] = False
        out.append({"pixel_values": [Image.fromarray(o) for o in obs], "obs": obs, "input_ids": ids, "attention_mask": mask, "labels": labels})
    return out


def main(kind="UADA", optimizer="adamW"):
    del STEPS[:], GRADS[:], PATCHES[:]
    cp, mp = glue.load_reference_model_module()
    glue.SPECS.update({"vit_large_patch14_reg4_dinov2.lvd142m": (32, 3, 2, 64, 5, True), "vit_so400m_patch14_siglip_224": (40, 3, 2, 72, 0, False)})
    text = dict(hidden_size=HIDDEN, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2, vocab_size=VOCAB,
                max_position_embeddings=512, rms_norm_eps=1e-6, pad_token_id=32000)
    cfg = cp.OpenVLAConfig(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure", image_sizes=[IMG, IMG],
                           text_config=text, attn_implementation="eager")
    torch.manual_seed(0)
    model = mp.OpenVLAForActionPrediction(cfg).eval()
    for vit in (model.vision_backbone.featurizer, model.vision_backbone.fused_featurizer):
        vit.patch_embed.num_patches = (IMG // 14) ** 2          # attribute of timm's PatchEmbed the attack loop reads (UADA.py:166)
    g = torch.Generator().manual_seed(11)
    weights = {}
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("embed_tokens.weight"):
                v = big_matrix(VOCAB, HIDDEN, 131, 71)
            elif n.endswith("lm_head.weight"):
                v = big_matrix(VOCAB, HIDDEN, 89, 153)
            elif "norm" in n and n.endswith("weight"):
                v = 1 + 0.2 * torch.randn(p.shape, generator=g)
            elif "scale_factor" in n:
                v = 0.5 * torch.randn(p.shape, generator=g)
            else:
                v = 0.15 * torch.randn(p.shape, generator=g)
            p.copy_(v.bfloat16().float())
            if not (n.endswith("embed_tokens.weight") or n.endswith("lm_head.weight")):
                weights["w:" + n] = p.detach().bfloat16().view(torch.int16).numpy()
    model = model.to(torch.bfloat16)

    fe = mg.load_reference_frontend()
    tok = mg.load_file_module("ref_action_tokenizer", "prismatic/vla/action_tokenizer.py")
    import transformers
    transformers.AdamW = HFAdamW
    uada = mg.load_attack_module(f"{kind}.py", fe, tok)
    uada.transformers.AdamW = HFAdamW

    class RoundTripTokenizer:
        """Stand-in for the Llama tokenizer (not available offline): ``decode`` then ``__call__`` returns [BOS, 29871 ('')] + the ids,
        which is what the real tokenizer does with a string of action tokens (TMA.py:93 drops the first two)."""
        vocab_size = 32000

        def decode(self, ids):
            return " ".join(str(int(i)) for i in ids)

        def __call__(self, text):
            return argparse.Namespace(input_ids=[1, 29871] + [int(t) for t in text.split()])

    class Proc:
        tokenizer = RoundTripTokenizer()

        class image_processor:
            apply_transform = None

    train, val = batches(N_TRAIN, 100), batches(N_VAL, 200)
    with tempfile.TemporaryDirectory() as d:
        extra = dict(alpha=0.8, belta=0.2) if kind == "UPA" else {}
        att = uada.OpenVLAAttacker(model, Proc(), save_dir=d, optimizer=optimizer, resize_patch=False, **extra)
        att.plot_loss = lambda: None          # matplotlib / seaborn are not installed
        random.seed(42)
        np.random.seed(42)
        torch.manual_seed(42)
        rand = torch.rand

        def capturing_rand(*a, **k):          # the loop creates its patch with torch.rand(patch_size): keep a handle on that tensor
            t = rand(*a, **k)
            FINAL.append(t)
            return t
        uada.torch.rand = capturing_rand

        class Loader:
            """Yields fresh copies: on the CPU ``labels.to(device)`` aliases the loader's tensor and the reference's in-place
            ``mask_labels`` would otherwise corrupt the batch for the next pass over the loader (on a GPU ``.to`` copies)."""

            def __init__(self, bs):
                self.bs = bs

            def __len__(self):
                return len(self.bs)

            def __iter__(self):
                for b in self.bs:
                    yield {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items() if k != "obs"}

        strip = Loader
        step_arg = dict(alpha=LR) if kind == "TMA" else dict(lr=LR)
        att.patchattack_unconstrained(strip(train), strip(val), num_iter=NUM_ITER, target_action=np.zeros(7), patch_size=[3, P_HW, P_HW],
                                      accumulate_steps=1, maskidx=MASKIDX, warmup=0, filterGripTrainTo1=False,
                                      geometry=(kind != "TMA"),   # TMA.py:137-141 passes colorjitter= to a method without it: only geometry=False runs
                                      innerLoop=INNER, args=argparse.Namespace(wandb_project="false"), **step_arg)
        saved_last = torch.load(os.path.join(d, "last", "patch.pt"))
        saved_best = torch.load(os.path.join(d, "0", "patch.pt"))
    out = dict(weights) if kind == "UADA" else {}      # the three fixtures share the weights and batches of the UADA one
    if kind == "UADA":
        # filter_train (UADA.py:309-340) on crafted gripper labels: 3 of 5 closed, none, exactly 8, 11 of 12 (random.sample of 8)
        for c, grips in enumerate(([1, 0, 1, 1, 0], [0, 0, 0], [1] * 8, [1] * 11 + [0])):
            n, T2 = len(grips), 12
            lab_in = torch.full((n, T2), -100, dtype=torch.int64)
            for r, gv in enumerate(grips):
                lab_in[r, T2 - 8:T2 - 1] = torch.tensor([31800 + r] * 6 + [31744 if gv else 31872])
                lab_in[r, T2 - 1] = 2
            batch = {"labels": lab_in.clone(), "input_ids": torch.arange(n * T2).view(n, T2), "attention_mask": torch.ones(n, T2, dtype=torch.bool),
                     "pixel_values": list(range(n))}
            random.seed(5)
            fl, fm, fi, fp = att.filter_train(batch)
            out[f"ft{c}_grips"] = np.array(grips)
            out[f"ft{c}_labels"], out[f"ft{c}_mask"], out[f"ft{c}_ids"], out[f"ft{c}_pixels"] = fl.numpy(), fm.numpy(), fi.numpy(), np.array(fp)
    for name, bs in ((("train", train), ("val", val)) if kind == "UADA" else ()):
        for k, b in enumerate(bs):
            out[f"{name}{k}_obs"] = b["obs"]
            out[f"{name}{k}_ids"] = b["input_ids"].numpy()
            out[f"{name}{k}_mask"] = b["attention_mask"].numpy()
            out[f"{name}{k}_labels"] = b["labels"].numpy()
    lists = {"UADA": ["train_CE_loss", "train_MSE_distance_loss", "train_UAD", "val_CE_loss", "val_MSE_Distance", "val_UAD"],
             "UPA": ["train_CE_loss", "avg_angle_loss", "avg_distance_loss", "avg_reserve_loss"],
             "TMA": ["train_CE_loss", "train_inner_avg_loss", "val_CE_loss", "val_L1_loss", "val_ASR"]}[kind]
    for name in lists:
        out[name] = np.array([float(v) for v in getattr(att, name)])
    if optimizer == "adamW":
        out.update(grads=torch.stack(GRADS).numpy(), patches=torch.stack(PATCHES).numpy())
    out.update(saved_last=saved_last.numpy(), saved_best=saved_best.numpy(), final_patch=FINAL[-1].detach().numpy())
    suffix = ("" if kind == "UADA" else "_" + kind.lower()) + ("" if optimizer == "adamW" else "_" + optimizer)
    np.savez_compressed(os.path.join(HERE, f"reference_golden_loop{suffix}.npz"), **out)
    uada.torch.rand = rand
    print(kind, optimizer, "steps", len(GRADS), {n: out[n].tolist() for n in lists})


if __name__ == "__main__":
    for k in (sys.argv[1:] or ["UADA", "UPA", "TMA", "TMA:pgd"]):
        main(*k.split(":"))
