"""The drop-in OpenVLAAttacker classes (UADA / UPA / TMA) run end to end on the CUDA engine with a reduced-depth model:
reference constructor + patchattack_unconstrained signatures, synthetic collator-shaped batches (PIL images), patch.pt
written in the reference's format."""
import argparse
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200.config import tiny  # noqa: E402
from roboticattack_b200.synthetic import synthetic_batch  # noqa: E402
from roboticattack_b200.weights import random_state_dict  # noqa: E402


def loader(cfg, n, B, T, seed, as_pil):
    from PIL import Image
    out = []
    for i in range(n):
        b = synthetic_batch(cfg, B, T, seed=seed + i, ragged=(i % 2 == 1))
        px = [Image.fromarray(o.numpy()) for o in b["obs"]] if as_pil else b["obs"]
        out.append({"pixel_values": px, "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
    return out


@pytest.fixture(scope="module")
def model():
    cfg = tiny(img=56, llm_layers=2, vit_depth=3)
    return cfg, random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test")


ARGS = argparse.Namespace(wandb_project="false")


@pytest.mark.parametrize("kind", ["UADA", "UPA", "TMA"])
def test_attacker_api_runs(model, tmp_path, kind):
    import importlib
    cfg, sd = model
    mod = importlib.import_module(f"roboticattack_b200.white_patch.{kind}")
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42)
    kw = dict(alpha=0.8, belta=0.2) if kind == "UPA" else {}
    att = mod.OpenVLAAttacker(sd, None, save_dir=str(tmp_path), optimizer="adamW", resize_patch=False, cfg=cfg, **kw)
    att.val_batches = 2
    train, val = loader(cfg, 3, 2, 16, 100, as_pil=True), loader(cfg, 2, 2, 16, 200, as_pil=False)
    common = dict(num_iter=3, target_action=np.zeros(7), patch_size=[3, 12, 12], accumulate_steps=1, maskidx=[0, 1, 2], warmup=1,
                  filterGripTrainTo1=False, geometry=True, innerLoop=2, args=ARGS)
    if kind == "TMA":
        patch = att.patchattack_unconstrained(train, val, alpha=2e-3, **common)
    else:
        patch = att.patchattack_unconstrained(train, val, lr=2e-3, **common)
    assert patch.shape == (3, 12, 12) and patch.dtype == torch.float32 and patch.device.type == "cpu"
    assert 0.0 <= patch.min() and patch.max() <= 1.0
    saved = torch.load(os.path.join(tmp_path, "last", "patch.pt"), weights_only=True)
    assert saved.shape == (3, 12, 12) and saved.dtype == torch.float32
    assert len(att.train_CE_loss) >= 3 and all(np.isfinite(att.train_CE_loss))
    assert att.host.opt_step == 6          # 3 outer x 2 inner AdamW steps
    # validation side effects of the reference (UADA.py:193-292, UPA.py:193-275, TMA.py:202-383): images of the last
    # validation batch with the patch, metric lists as .pkl, restart state next to patch.pt
    vd = os.path.join(tmp_path, "last", "val_related_data")
    from PIL import Image
    im = Image.open(os.path.join(vd, "0.png"))
    assert im.size == (cfg.img, cfg.img) and os.path.exists(os.path.join(vd, "1.png"))
    assert os.path.exists(os.path.join(tmp_path, "last", "attack_state.pt"))
    assert os.path.exists(os.path.join(tmp_path, "0", "patch.pt")), "the first validation always improves on the initial best"
    pk = {"UADA": ["val_MSE_Distance", "val_UAD"], "UPA": ["avg_reserve_loss", "avg_angle_loss"], "TMA": ["val_L1_loss", "val_ASR"]}[kind]
    import pickle
    for name in pk:
        with open(os.path.join(tmp_path, f"{name}.pkl"), "rb") as f:
            vals = pickle.load(f)
        assert len(vals) == 1 and np.isfinite(vals[0]), (name, vals)
    if kind == "TMA":
        assert os.path.exists(os.path.join(vd, "continuous_actions_pred.pt"))
        assert 0.0 <= att.val_ASR[0] <= 1.0


def test_tma_gripper_validation_filters_on_clean_prediction(model, tmp_path):
    """maskidx == [6]: samples whose CLEAN gripper prediction is wrong are dropped before the attacked forward, and the
    0->other / 1->other / other->0 rates are logged (TMA.py:223-250,298-306,318-336)."""
    from roboticattack_b200.white_patch.TMA import OpenVLAAttacker
    cfg, sd = model
    random.seed(2)
    np.random.seed(2)
    torch.manual_seed(2)
    att = OpenVLAAttacker(sd, None, save_dir=str(tmp_path), optimizer="adamW", cfg=cfg)
    att.val_batches = 2
    logs = []
    att._log = lambda args, data, step: logs.append(data)
    data = loader(cfg, 2, 3, 16, 400, as_pil=True)
    att.patchattack_unconstrained(data, data, num_iter=1, patch_size=[3, 8, 8], alpha=2e-3, maskidx=[6], geometry=False,
                                  innerLoop=1, args=ARGS)
    val = [d for d in logs if "VAL_avg_CE_loss" in d]
    assert len(val) == 1 and "ALL_ASR_6" in val[0] and np.isfinite(val[0]["VAL_avg_L1_loss"])


def test_tma_pgd_and_accumulate(model, tmp_path):
    """sign-PGD branch (TMA.py:171-175) and accumulate_steps > 1 (steps gated on the outer index, TMA.py:165)."""
    from roboticattack_b200.white_patch.TMA import OpenVLAAttacker
    cfg, sd = model
    random.seed(1)
    np.random.seed(1)
    torch.manual_seed(1)
    att = OpenVLAAttacker(sd, None, save_dir=str(tmp_path), optimizer="pgd", cfg=cfg)
    train = loader(cfg, 4, 2, 16, 300, as_pil=False)
    p = att.patchattack_unconstrained(train, train, num_iter=4, patch_size=[3, 8, 8], alpha=1 / 255, accumulate_steps=2,
                                      maskidx=[0], geometry=False, innerLoop=1, args=ARGS)
    assert att.host.opt_step == 2          # only outer iterations 1 and 3 step
    assert torch.isfinite(p).all()


def test_random_patch_transform_autograd():
    """RandomPatchTransform under the reference's method names, with gradients to the patch through the CUDA backward."""
    from oracle import frontend as ofe
    from roboticattack_b200.config import NORM_MEAN, NORM_STD
    from roboticattack_b200.white_patch.appply_random_transform import RandomPatchTransform
    obs = torch.randint(0, 256, (3, 56, 56, 3), dtype=torch.uint8)
    patch = torch.rand(3, 10, 10, device="cuda", requires_grad=True)
    t = RandomPatchTransform("cuda:0", False)
    random.seed(3)
    np.random.seed(3)
    out = t.apply_random_patch_batch(obs, patch, geometry=True)
    assert out.shape == (3, 6, 56, 56) and out.dtype == torch.bfloat16
    gw = torch.randn_like(out, dtype=torch.float32)
    (out.float() * gw).sum().backward()
    random.seed(3)
    np.random.seed(3)
    xy, th = ofe.draw_placements(3, (56, 56), (10, 10), True)
    pr = patch.detach().cpu().clone().requires_grad_(True)
    ref = ofe.apply_patch_batch(obs, pr, xy, th, ofe.MODE_WARP, NORM_MEAN, NORM_STD).to(torch.bfloat16)
    (ref.float() * gw.cpu().bfloat16().float()).sum().backward()
    torch.testing.assert_close(patch.grad.cpu(), pr.grad, rtol=2e-2, atol=2e-2 * pr.grad.abs().max().item())
    assert t.im_process(obs).shape == (3, 6, 56, 56)
    assert t.paste_patch_fix(obs, patch).shape == (3, 6, 56, 56)
