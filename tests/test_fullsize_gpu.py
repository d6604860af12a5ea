"""Size-independent properties at the BASELINE.json sizes (OpenVLA-7B shapes, per-GPU bs = 8, patch 3x50x50, T = 33):
the oracle cannot run here in seconds, so the full-size engine is checked through properties the path must have.

  * sharding: for the DDP loss (mean over tokens, equal token count per sample) the gradient of the batch of 8 equals the
    mean of the gradients of its two halves of 4 -- exactly what the multi-GPU path relies on (UADA_ddp.py:157-166,206);
  * sample-permutation invariance of loss and gradient;
  * run-to-run reproducibility (the only non-deterministic reduction is the fp32 atomics of the front-end backward);
  * forward-only (validation) pass returns the same scalars as the training pass.
"""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200 import _lib, labels as lab  # noqa: E402
from roboticattack_b200.config import openvla_7b  # noqa: E402
from roboticattack_b200.engine import LossSpec, VLAEngine  # noqa: E402
from roboticattack_b200.synthetic import draw_placements, synthetic_batch  # noqa: E402


@pytest.fixture(scope="module")
def full():
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    cfg = openvla_7b()
    eng = VLAEngine(cfg, 8, 33)
    eng.load_random_weights(seed=0, init="reference")
    batch = synthetic_batch(cfg, 8, 33, seed=1234, ragged=True)
    batch["labels"] = lab.mask_labels_uada(batch["labels"].clone(), [0, 1, 2])
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(8, (224, 224), (50, 50), True, steps=1)
    torch.manual_seed(42)
    patch = torch.rand(3, 50, 50).cuda()
    yield cfg, eng, batch, xy, theta, patch
    del eng
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def run(eng, batch, xy, theta, patch, idx, loss, forward_only=False):
    sub = {k: v[idx] for k, v in batch.items()}
    eng.ensure_plan(len(idx), batch["input_ids"].shape[1])
    eng.set_batch(sub["obs"], sub["input_ids"], sub["attention_mask"], sub["labels"])
    eng.set_placements(xy[:, idx], theta[:, idx])
    g = torch.zeros_like(patch)
    sc = torch.zeros(_lib.NUM_SCALARS, device="cuda")
    pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
    eng.fwd_bwd(patch, 0, _lib.FE_WARP, loss, g, sc, pred, forward_only=forward_only)
    torch.cuda.synchronize()
    return sc.cpu(), g.cpu(), pred.cpu()


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_fullsize_properties(full):
    cfg, eng, batch, xy, theta, patch = full
    ddp = LossSpec(_lib.LOSS_UADA_DDP, mse_weight=5.0)
    all8 = list(range(8))
    sc8, g8, pred8 = run(eng, batch, xy, theta, patch, all8, ddp)
    assert torch.isfinite(g8).all() and g8.abs().max() > 0 and np.isfinite(sc8[_lib.S_LOSS].item())
    assert sc8[_lib.S_NTOK].item() == 8 * 4 and sc8[_lib.S_NACT].item() == 8 * 3      # maskidx 0,1,2 + EOS per sample

    # reproducibility
    sc8b, g8b, pred8b = run(eng, batch, xy, theta, patch, all8, ddp)
    assert torch.equal(pred8, pred8b) and sc8[_lib.S_LOSS].item() == sc8b[_lib.S_LOSS].item()
    assert rel(g8b, g8) < 1e-5

    # sharding: grad(8) == mean(grad(first 4), grad(last 4)); loss likewise
    scA, gA, _ = run(eng, batch, xy, theta, patch, [0, 1, 2, 3], ddp)
    scB, gB, _ = run(eng, batch, xy, theta, patch, [4, 5, 6, 7], ddp)
    np.testing.assert_allclose(0.5 * (scA[_lib.S_LOSS].item() + scB[_lib.S_LOSS].item()), sc8[_lib.S_LOSS].item(), rtol=2e-3)
    r = rel(0.5 * (gA + gB), g8)
    print(f"sharding: rel diff of mean-of-halves vs full batch gradient = {r:.4f}")
    assert r < 0.05          # different GEMM M -> different tile variants / accumulation order: bf16-level agreement

    # permutation invariance
    perm = [3, 7, 0, 5, 1, 6, 2, 4]
    scP, gP, _ = run(eng, batch, xy, theta, patch, perm, ddp)
    np.testing.assert_allclose(scP[_lib.S_LOSS].item(), sc8[_lib.S_LOSS].item(), rtol=1e-4)
    assert rel(gP, g8) < 0.02

    # validation pass == training pass scalars
    scF, gF, predF = run(eng, batch, xy, theta, patch, all8, ddp, forward_only=True)
    assert scF[_lib.S_LOSS].item() == sc8[_lib.S_LOSS].item() and torch.equal(predF, pred8) and gF.abs().max() == 0


def test_fullsize_update_moves_patch_inside_unit_box(full):
    cfg, eng, batch, xy, theta, patch = full
    loss = LossSpec(_lib.LOSS_UADA, mse_weight=5.0)
    p = patch.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sc, g, _ = run(eng, batch, xy, theta, p, list(range(8)), loss)
    eng.patch_update(p, g.cuda(), m, v, 1, 2e-3)
    torch.cuda.synchronize()
    d = (p - patch).abs().cpu()
    assert 0 < d.max().item() <= 2e-3 * 1.001 and p.min() >= 0 and p.max() <= 1
    # first transformers.AdamW step: |delta| = lr * |g| / (|g| + eps / sqrt(1 - beta2)) away from the clamp
    expect = 2e-3 * g.abs() / (g.abs() + 1e-6 / (1 - 0.999) ** 0.5)
    inside = (patch.cpu() > 0.01) & (patch.cpu() < 0.99)
    torch.testing.assert_close(d[inside], expect[inside], rtol=2e-3, atol=1e-7)
    assert torch.equal(torch.sign(p.cpu() - patch.cpu())[inside], -torch.sign(g)[inside])


def _loader(cfg, n, B, T, seed):
    out = []
    for i in range(n):
        b = synthetic_batch(cfg, B, T, seed=seed + i, ragged=(i % 2 == 1))
        out.append({"pixel_values": b["obs"], "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
    return out


def test_baseline_configs_run_at_full_size(full, tmp_path):
    """BASELINE.json configs 2-4 through the drop-in classes on OpenVLA-7B shapes (two outer x two inner iterations each):
    UADA bs=8 p=50; TMA (zero target) bs=8 geometry; UPA bs=16 p=70 with a maskidx sweep in its -CE mode and the default
    position-aware loss.  One engine (30 GB of weights) is shared; the activation arena is re-planned per (B, T)."""
    import argparse
    from roboticattack_b200.attacker import TMAAttacker, UADAAttacker, UPAAttacker
    cfg, eng, *_ = full
    args = argparse.Namespace(wandb_project="false")
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42)

    a = UADAAttacker(eng, None, save_dir=str(tmp_path / "uada"), optimizer="adamW", cfg=cfg)
    a.val_batches = 1
    p = a.patchattack_unconstrained(_loader(cfg, 2, 8, 33, 10), _loader(cfg, 1, 8, 33, 20), num_iter=2, patch_size=[3, 50, 50], lr=2e-3,
                                    maskidx=[0, 1, 2], warmup=0, geometry=True, innerLoop=2, args=args)
    assert p.shape == (3, 50, 50) and torch.isfinite(p).all() and len(a.train_CE_loss) == 4 and np.isfinite(a.train_CE_loss).all()

    t = TMAAttacker(eng, None, save_dir=str(tmp_path / "tma"), optimizer="adamW", cfg=cfg)
    p = t.patchattack_unconstrained(_loader(cfg, 2, 8, 33, 30), None, num_iter=2, target_action=np.zeros(7), patch_size=[3, 50, 50],
                                    alpha=2e-3, maskidx=[0, 1, 2], warmup=0, geometry=True, innerLoop=2, args=args)
    assert torch.isfinite(p).all() and np.isfinite(t.train_CE_loss).all()
    first, last = t.train_inner_avg_loss[0], t.train_CE_loss[-1]
    assert first > 0 and last > 0

    for maskidx in ([0], [3], [6]):
        u = UPAAttacker(eng, None, save_dir=str(tmp_path / f"upa{maskidx[0]}"), optimizer="adamW", alpha=0.8, belta=0.2, cfg=cfg)
        p = u.patchattack_unconstrained(_loader(cfg, 1, 16, 33, 40), None, num_iter=1, patch_size=[3, 70, 70], lr=2e-3, maskidx=maskidx,
                                        warmup=0, geometry=True, innerLoop=2, reverse_direction=False, args=args)
        assert p.shape == (3, 70, 70) and torch.isfinite(p).all() and np.isfinite(u.train_CE_loss).all()
    u = UPAAttacker(eng, None, save_dir=str(tmp_path / "upa"), optimizer="adamW", alpha=0.8, belta=0.2, cfg=cfg)
    p = u.patchattack_unconstrained(_loader(cfg, 1, 16, 33, 50), None, num_iter=1, patch_size=[3, 70, 70], lr=2e-3, maskidx=[0, 1, 2],
                                    warmup=0, geometry=True, innerLoop=2, reverse_direction=True, args=args)
    assert torch.isfinite(p).all() and np.isfinite(u.train_CE_loss).all()
    eng.ensure_plan(8, 33)


@pytest.mark.parametrize("B", [1, 2])
def test_fullsize_kv_decode_matches_teacher_forced_forward(full, B):
    """Greedy action decode at OpenVLA-7B shapes (predict_action, modeling_prismatic.py:506-536): every single-position step on
    the KV cache (fused skinny kernels of csrc/decode.cu, one recorded step replayed per token) must produce the logits of a
    full forward pass over the same prefix (tcgen05 GEMMs + flash attention).  Both are bf16 pipelines that differ in
    summation order only: logits agree to a few bf16 roundings, the chosen token is the argmax of the step's own logits, and
    where the two paths pick different tokens the full forward itself calls the decision a tie (random weights make flat
    logits, so such ties do occur)."""
    from roboticattack_b200.policy import ActionPolicy
    cfg, eng, batch, *_ = full
    pol = ActionPolicy(eng)
    obs = batch["obs"][:B].contiguous()
    prompt = batch["input_ids"][:B, :25].clone()
    n = 7
    V = cfg.llm.vocab
    toks = torch.from_numpy(pol.generate_action_tokens(obs, prompt, n))
    assert toks.shape == (B, n) and toks.min() >= 0 and toks.max() < V
    again = torch.from_numpy(pol.generate_action_tokens(obs, prompt, n))
    assert torch.equal(toks, again), "replaying the recorded decode step must be deterministic"
    worst = 0.0
    for k in range(n):
        part = torch.from_numpy(pol.generate_action_tokens(obs, prompt, k + 1))
        assert torch.equal(part, toks[:, :k + 1]), f"decode of {k + 1} tokens is not a prefix of the decode of {n}"
        kv = eng.tap("logits", dtype=torch.float32, max_elems=B * V).view(B, V).clone()
        assert torch.equal(kv.argmax(dim=1).cpu(), toks[:, k]), f"token {k} is not the argmax of its step's logits"
        # the same prefix through the full forward (one pass, no cache): generate 1 token from prompt + first k tokens
        forced = torch.cat([prompt, toks[:, :k]], dim=1)
        nxt = torch.from_numpy(pol.generate_action_tokens(obs, forced, 1, kv_cache=False))
        ref = eng.tap("logits", dtype=torch.float32, max_elems=B * V).view(B, V).clone()
        err = ((kv - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
        worst = max(worst, err)
        print(f"B={B} token {k}: relative logits error {err:.4g}, cos {torch.nn.functional.cosine_similarity(kv, ref, dim=1).min().item():.6f}")
        assert err < 5e-2, f"token {k}: kv-cache logits differ from the full forward by {err:.3g}"
        tol = 4 * 2.0 ** -8 * ref.abs().max().item()
        for b in range(B):
            if nxt[b, 0] != toks[b, k]:
                margin = (ref[b, nxt[b, 0]] - ref[b, toks[b, k]]).item()
                assert 0 <= margin <= tol, f"token {k} sample {b}: paths disagree at a margin of {margin:.4g} (> {tol:.4g})"
    print(f"kv decode vs teacher-forced forward, B={B}: worst relative logits error {worst:.3g}")
