"""Engine vs oracle at the BASELINE.json sizes: OpenVLA-7B shapes, random weights of the "test" initialisation (every branch
carries signal: LayerScale O(0.3), non-zero biases), the oracle run on the same GPU.

  config #2  UADA      bs 8,  patch 3x50x50, T = 33 ragged, geometry, maskidx 0,1,2      (UADA.py:133-159)
  config #3  TMA       bs 8,  patch 3x50x50, zero-action target, geometry               (TMA.py:132-175)
  config #4  UPA       bs 16, patch 3x70x70, position-aware loss alpha .8 / belta .2    (UPA.py:133-168)

Two oracle runs per config on the SAME bf16-valued weights: fp32 ("truth") and bf16 (the reference's own precision).
Asserted per config:
  * loss / CE (/ UAD) of the engine within max(3 x the oracle-bf16's own deviation from truth, 2 %);
  * forward taps against truth: front-end output, both tower outputs, multimodal embedding;
  * patch-gradient relative error vs truth <= 2.5 x the oracle-bf16's and cosine >= 0.999;
  * greedy action decode with the KV cache (predict_action, batch 2): every step's logits row against the fp32 oracle's
    logits for the same prefix, and every chosen token within that error of the oracle's best;
and for configs #2 (transformers.AdamW, lr 2e-3) and #3 (sign-PGD, alpha 1/255) a free-running 10-step trajectory against
the oracle-bf16 trajectory: the loss of EVERY step within 2 %, and the final patch within the stated distribution
(Adam's first steps and sign-PGD move a pixel by +-lr whatever |g| is, so one sign flip of a noise-level gradient entry
costs 2 lr: the bound is on the fraction of pixels within lr/2 and on the mean, not on L-inf).

Memory: the oracle phase (30 GB fp32 weights + up to ~90 GB of fp32 autograd state at bs 16) runs BEFORE the engine is
created (30 GB weight arena + 20 GB activations).
"""
import gc
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200 import _lib, labels as lab  # noqa: E402
from roboticattack_b200.config import NORM_MEAN, NORM_STD, openvla_7b  # noqa: E402
from roboticattack_b200.engine import LossSpec, VLAEngine  # noqa: E402
from roboticattack_b200.synthetic import draw_placements, synthetic_batch  # noqa: E402
from roboticattack_b200.weights import random_state_dict  # noqa: E402

STEPS = 10
CONFIGS = {
    "uada": dict(B=8, p=50, loss="uada", opt="adamw", lr=2e-3, traj=True),
    "tma": dict(B=8, p=50, loss="ce", opt="pgd", lr=1 / 255, traj=True),
    "upa": dict(B=16, p=70, loss="upa", opt="adamw", lr=2e-3, traj=False),
}
SPECS = {"uada": LossSpec(_lib.LOSS_UADA, 5.0), "upa": LossSpec(_lib.LOSS_UPA, 0, 0.8, 0.2), "ce": LossSpec(_lib.LOSS_CE, ce_scale=1.0)}


def make_inputs(cfg, name, c):
    batch = synthetic_batch(cfg, c["B"], 33, seed={"uada": 1234, "tma": 77, "upa": 99}[name], ragged=True)
    if name == "uada":
        batch["labels"] = lab.mask_labels_uada(batch["labels"].clone(), [0, 1, 2])
    elif name == "tma":
        batch["labels"] = lab.tma_labels(batch["labels"], lab.tma_target(np.zeros(7), [0, 1, 2]))
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(c["B"], (cfg.img, cfg.img), (c["p"], c["p"]), True, steps=STEPS)
    torch.manual_seed(42)
    patch = torch.rand(3, c["p"], c["p"])
    return batch, xy, theta, patch


def oracle_pass(sd, cfg, batch_d, patch, xy, theta, loss_kind, dtype, taps=False):
    from oracle import frontend as ofe, losses as ol, model as om
    p = patch.clone().requires_grad_(True)
    px = ofe.apply_patch_batch(batch_d["obs"], p, xy, theta, ofe.MODE_WARP, NORM_MEAN, NORM_STD)
    pxd = px.to(torch.bfloat16).to(dtype)          # UADA.py:142: the model sees the bf16-rounded image in every precision
    out = {}
    if taps:
        with torch.no_grad():
            feats = om.vision_backbone(sd, cfg, pxd)
            x0, _, _ = om.splice(sd, cfg, om.projector(sd, feats), batch_d["input_ids"], batch_d["attention_mask"], batch_d["labels"])
            out.update(px=px.detach().bfloat16().float().cpu(), feats=feats.float().cpu(), x0=x0.float().cpu())
    o = om.forward(sd, cfg, batch_d["input_ids"], batch_d["attention_mask"], pxd, batch_d["labels"])
    logits = o.logits.float()
    uad = 0.0
    if loss_kind == "uada":
        mse, uad = ol.weighted_loss_uada(logits, batch_d["labels"], 5)
        loss = mse + 1 / o.loss
    elif loss_kind == "upa":
        loss, _, _ = ol.weighted_loss_upa(logits, batch_d["labels"], 0.8, 0.2, cfg.num_patches)
    else:
        loss = o.loss
    loss.backward()
    out.update(loss=loss.item(), ce=o.loss.item(), uad=float(uad), grad=p.grad.detach().float().cpu())
    return out


def oracle_trajectory(sd, cfg, batch_d, patch0, xy, theta, loss_kind, opt, lr):
    from oracle import optim as oo
    p = patch0.clone()
    adam = oo.HFAdamW(p.shape, lr)
    losses = []
    for s in range(STEPS):
        r = oracle_pass(sd, cfg, batch_d, p, xy[s], theta[s], loss_kind, torch.bfloat16)
        losses.append(r["loss"])
        g = r["grad"].to(p.device)
        if opt == "adamw":
            adam.step(p, g)
            p.clamp_(0, 1)
        else:
            p = oo.pgd_step(p, g, lr)
    return losses, p.cpu()


def decode_vs_oracle(eng, sd32, cfg, dev, B=2, n=7, T0=25):
    """predict_action's greedy decode (modeling_prismatic.py:506-536) on the engine -- prefill + n - 1 replayed single-position
    steps on the KV cache -- against ONE fp32 oracle forward over prompt + the engine's tokens (causal: row T0 - 1 + k of the text
    holds the logits that chose token k)."""
    from oracle import frontend as ofe, model as om
    from roboticattack_b200.policy import ActionPolicy
    batch = synthetic_batch(cfg, B, 33, seed=5)
    prompt = batch["input_ids"][:, :T0].clone()
    pol = ActionPolicy(eng)
    V = cfg.llm.vocab
    toks = torch.from_numpy(pol.generate_action_tokens(batch["obs"], prompt, n))
    steps = []
    for k in range(n):       # the logits of step k are what the engine holds after a decode of k + 1 tokens
        part = torch.from_numpy(pol.generate_action_tokens(batch["obs"], prompt, k + 1))
        assert torch.equal(part, toks[:, :k + 1])
        steps.append(eng.tap("logits", dtype=torch.float32, max_elems=B * V).view(B, V).clone().cpu())
    with torch.device(dev), torch.no_grad():
        xy, theta = np.zeros((B, 2), dtype=np.int32), np.zeros((B, 2, 3), dtype=np.float32)
        px = ofe.apply_patch_batch(batch["obs"].to(dev), torch.zeros(3, 1, 1), xy, theta, ofe.MODE_NONE, NORM_MEAN, NORM_STD)
        ids = torch.cat([prompt, toks[:, :-1]], dim=1).to(dev)
        out = om.forward(sd32, cfg, ids, torch.ones_like(ids, dtype=torch.bool), px.to(torch.bfloat16).float(), None)
        ref = out.logits[:, -n:].float().cpu()          # rows of the last prompt token and of the first n - 1 generated tokens
    return {"tokens": toks, "engine_logits": torch.stack(steps, 1), "oracle_logits": ref}


@pytest.fixture(scope="module")
def results():
    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~150 GB of device memory (fp32 oracle at OpenVLA-7B shapes)")
    cfg = openvla_7b()
    dev = torch.device("cuda")
    sd32 = {k: v.float() for k, v in random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16, init="test").items()}
    res = {name: {"inputs": make_inputs(cfg, name, c)} for name, c in CONFIGS.items()}
    # ---------------- phase 1: the oracle ----------------
    with torch.device(dev):          # the oracle builds its masks / tables with factory functions
        for name, c in CONFIGS.items():
            batch, xy, theta, patch = res[name]["inputs"]
            bd = {k: v.to(dev) for k, v in batch.items()}
            pd = patch.to(dev)
            r = res[name]
            r["truth"] = oracle_pass(sd32, cfg, bd, pd, xy[0], theta[0], c["loss"], torch.float32, taps=True)
            gc.collect()
            torch.cuda.empty_cache()
            sd16 = {k: v.bfloat16() for k, v in sd32.items()}
            r["bf16"] = oracle_pass(sd16, cfg, bd, pd, xy[0], theta[0], c["loss"], torch.bfloat16)
            if c["traj"]:
                r["traj"] = oracle_trajectory(sd16, cfg, bd, pd, xy, theta, c["loss"], c["opt"], c["lr"])
            del sd16
            gc.collect()
            torch.cuda.empty_cache()
    # ---------------- phase 2: the engine ----------------
    eng = VLAEngine(cfg, 8, 33)
    eng.load_state_dict({k: v.bfloat16() for k, v in sd32.items()})
    res["decode"] = decode_vs_oracle(eng, sd32, cfg, dev)      # no autograd state: the fp32 oracle fits next to the engine
    del sd32
    gc.collect()
    torch.cuda.empty_cache()
    for name, c in CONFIGS.items():
        batch, xy, theta, patch = res[name]["inputs"]
        B = c["B"]
        eng.ensure_plan(B, 33)
        eng.set_batch(batch["obs"], batch["input_ids"], batch["attention_mask"], batch["labels"])
        eng.set_placements(xy, theta)
        pe = patch.cuda()
        g = torch.zeros_like(pe)
        sc = torch.zeros(_lib.NUM_SCALARS, device="cuda")
        pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
        eng.fwd_bwd(pe, 0, _lib.FE_WARP, SPECS[c["loss"]], g, sc, pred)
        torch.cuda.synchronize()
        e = {"loss": sc[_lib.S_LOSS].item(), "ce": sc[_lib.S_CE].item(), "uad": sc[_lib.S_UAD].item(), "grad": g.cpu()}
        e["px"] = eng.tap("px").view(B, 6, cfg.img, cfg.img).float().cpu()
        e["dino"] = eng.tap("dino_out").view(B, cfg.dino.tokens, cfg.dino.dim)[:, cfg.dino.num_prefix:].float().cpu()
        e["sig"] = eng.tap("siglip_out").view(B, cfg.siglip.tokens, cfg.siglip.dim).float().cpu()
        e["x0"] = eng.tap("llm_x0").view(B, eng.L, cfg.llm.hidden).float().cpu()
        if c["traj"]:
            m, v = torch.zeros_like(pe), torch.zeros_like(pe)
            hist = torch.zeros(STEPS, _lib.NUM_SCALARS, device="cuda")
            eng.set_step_state(0, 0)
            for s in range(STEPS):     # the product's own step: vla_attack_step (eager, recording, then graph replays)
                eng.attack_step(pe, m, v, g, hist, pred, _lib.FE_WARP, SPECS[c["loss"]], c["lr"],
                                opt_kind=_lib.OPT_ADAMW if c["opt"] == "adamw" else _lib.OPT_PGD)
            torch.cuda.synchronize()
            e["traj"] = (hist[:, _lib.S_LOSS].cpu().tolist(), pe.cpu())
        res[name]["engine"] = e
    del eng
    gc.collect()
    torch.cuda.empty_cache()
    return cfg, res


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_loss_and_gradient_vs_oracle(results, name):
    cfg, res = results
    t, o, e = res[name]["truth"], res[name]["bf16"], res[name]["engine"]
    cos = torch.nn.functional.cosine_similarity(e["grad"].flatten(), t["grad"].flatten(), dim=0).item()
    e_or, e_en = rel(o["grad"], t["grad"]), rel(e["grad"], t["grad"])
    print(f"[{name}] loss truth {t['loss']:.6f} oracle-bf16 {o['loss']:.6f} engine {e['loss']:.6f} | CE {t['ce']:.5f} {o['ce']:.5f} {e['ce']:.5f} | "
          f"grad rel-err vs truth: oracle-bf16 {e_or:.4f} engine {e_en:.4f}, cos(engine, truth) {cos:.5f}, |g|max {t['grad'].abs().max().item():.3g}")
    assert abs(e["loss"] - t["loss"]) <= max(3 * abs(o["loss"] - t["loss"]), 2e-2 * abs(t["loss"]))
    assert abs(e["ce"] - t["ce"]) <= max(3 * abs(o["ce"] - t["ce"]), 2e-2 * abs(t["ce"]))
    if name == "uada":
        assert abs(e["uad"] - t["uad"]) <= max(3 * abs(o["uad"] - t["uad"]), 5e-2 * abs(t["uad"]) + 1e-3)
    assert torch.isfinite(e["grad"]).all() and e["grad"].abs().max() > 0
    assert e_en <= max(2.5 * e_or, 0.05), f"engine gradient error {e_en:.4f} vs oracle-bf16 {e_or:.4f}"
    assert cos >= 0.999, cos


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_taps_vs_oracle(results, name):
    cfg, res = results
    t, e = res[name]["truth"], res[name]["engine"]
    frac = ((e["px"] - t["px"]).abs() > 0).float().mean().item()
    r_d, r_s = rel(e["dino"], t["feats"][..., :cfg.dino.dim]), rel(e["sig"], t["feats"][..., cfg.dino.dim:])
    L = e["x0"].shape[1]
    r_x = rel(e["x0"], t["x0"][:, :L])
    print(f"[{name}] px differing fraction {frac:.5f}; rel-err vs fp32 truth: dino_out {r_d:.4f} siglip_out {r_s:.4f} llm_x0 {r_x:.4f}")
    # bf16 rounding-boundary flips + patch-border pixels (see test_kernels_gpu.py::test_frontend_vs_oracle, which compares with
    # ATen's CPU grid_sample; here the oracle's front end runs ATen's CUDA path, whose coordinate maths differ from the CPU
    # one by an fp32 ulp as well): few pixels differ, by one bf16 ulp, never by more than two (0.02 at the -100 border)
    diff = (e["px"] - t["px"]).abs()
    assert frac < 4e-3
    assert (diff > 2.0 ** -8 * t["px"].abs().clamp_min(1.0) * 1.01).float().mean().item() < 2e-4
    assert (diff <= torch.maximum(torch.full_like(diff, 0.02), 2.02 * 2.0 ** -8 * t["px"].abs())).all()
    assert r_d < 0.03 and r_s < 0.03 and r_x < 0.03


@pytest.mark.parametrize("name", [n for n, c in CONFIGS.items() if c["traj"]])
def test_ten_step_trajectory_vs_oracle(results, name):
    cfg, res = results
    lr = CONFIGS[name]["lr"]
    o_losses, o_patch = res[name]["traj"]
    e_losses, e_patch = res[name]["engine"]["traj"]
    dev_ = (e_patch - o_patch).abs()
    frac, mean, linf = (dev_ < lr / 2).float().mean().item(), dev_.mean().item(), dev_.max().item()
    print(f"[{name}] oracle losses {[f'{x:.4f}' for x in o_losses]}")
    print(f"[{name}] engine losses {[f'{x:.4f}' for x in e_losses]}")
    print(f"[{name}] final patch: Linf {linf:.4g} ({linf / lr:.2f} lr) mean {mean:.4g} ({mean / lr:.3f} lr) fraction within lr/2: {frac:.4f}")
    for s, (a, b) in enumerate(zip(e_losses, o_losses)):
        assert abs(a - b) <= 2e-2 * abs(b) + 1e-3, f"step {s}: engine {a} oracle {b}"
    moved = (o_patch - res[name]["inputs"][3]).abs().max().item()
    assert moved > 2 * lr, "the trajectory must move the patch"
    if CONFIGS[name]["opt"] == "adamw":
        assert frac >= 0.95 and mean <= lr / 4, (frac, mean)
    else:   # sign-PGD: every step is +-alpha; entries whose gradient is below bf16 noise flip freely
        assert frac >= 0.75 and mean <= lr / 2, (frac, mean)


def test_kv_cache_decode_vs_oracle(results):
    cfg, res = results
    d = res["decode"]
    toks, e, o = d["tokens"], d["engine_logits"], d["oracle_logits"]
    B, n, V = e.shape
    assert o.shape == e.shape
    worst, worst_gap = 0.0, 0.0
    for k in range(n):
        r = ((e[:, k] - o[:, k]).norm(dim=1) / o[:, k].norm(dim=1)).max().item()
        worst = max(worst, r)
        for b in range(B):
            assert e[b, k].argmax().item() == toks[b, k].item()
            # where the engine and the oracle choose different tokens the oracle itself must call it a near-tie: the oracle's
            # logit of the engine's token within a few bf16 roundings (2^-8 of the row's scale) of the oracle's best
            gap = (o[b, k].max() - o[b, k, toks[b, k]]).item() / (2.0 ** -8 * o[b, k].abs().max().item())
            worst_gap = max(worst_gap, gap)
    agree = (e.argmax(-1) == o.argmax(-1)).float().mean().item()
    print(f"[decode] kv-cache logits vs fp32 oracle, {n} steps x {B} samples: worst relative error {worst:.4f}; same token as the "
          f"oracle's argmax in {agree:.0%} of the decisions, worst oracle gap of a chosen token {worst_gap:.2f} bf16 ulps of the row scale")
    assert worst < 0.03, worst   # bf16 engine vs fp32 truth (measured 0.014); the other full-size taps sit at 0.010-0.011
    assert worst_gap <= 8, worst_gap   # measured 2.3
