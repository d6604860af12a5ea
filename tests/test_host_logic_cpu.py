"""Host-side logic of the product package, checked against the oracle and the golden vectors (CPU only)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from oracle import losses as ol
from roboticattack_b200 import labels as lab
from roboticattack_b200.attacker import cosine_with_warmup
from roboticattack_b200.config import flops_per_sample, openvla_7b, tiny
from roboticattack_b200.synthetic import draw_placements, synthetic_batch


def test_labels_match_reference_golden(golden):
    for mi in range(4):
        out = lab.mask_labels_uada(torch.from_numpy(golden["loss_labels"]).clone(), golden[f"uada_mask{mi}_idx"].tolist())
        np.testing.assert_array_equal(out.numpy(), golden[f"uada_mask{mi}_labels"])
    for mi in range(3):
        out = lab.mask_labels_upa(torch.from_numpy(golden["loss_labels"]).clone(), golden[f"upa_mask{mi}_idx"].tolist())
        np.testing.assert_array_equal(out.numpy(), golden[f"upa_mask{mi}_labels"])
    np.testing.assert_array_equal(lab.decode_token_ids_to_actions(golden["tok_ids"]), golden["tok_decode"])
    np.testing.assert_array_equal(lab.action_to_token_ids(golden["tok_actions"]), golden["tok_encode_ids"])


def test_tma_target_matches_oracle():
    b = synthetic_batch(tiny(), 3, 16, ragged=True)
    for maskidx in ([0], [0, 1, 2], [6, 7], list(range(8))):
        t = lab.tma_target(np.zeros(7), maskidx)
        t_or = ol.tma_target(ol.encode_actions_to_token_ids(np.zeros(7)), maskidx)
        assert torch.equal(t, t_or)
        assert torch.equal(lab.tma_labels(b["labels"], t), ol.tma_labels(b["labels"], t_or))
    assert lab.tma_target(np.zeros(7), list(range(8)))[:7].tolist() == [31872] * 7


def test_schedule_matches_transformers_golden(golden):
    lrs = np.array([2e-3 * cosine_with_warmup(s, 20, 2000) for s in range(2000)])
    np.testing.assert_allclose(lrs, golden["sched_lrs"], rtol=1e-12, atol=1e-18)


def test_placement_stream_matches_reference_protocol(golden):
    """draw_placements reproduces the reference's RNG order: the oracle front end fed with these draws equals the
    output of the reference's own apply_random_patch_batch under the same seeds."""
    obs = torch.from_numpy(golden["fe_s64_obs"])
    patch = torch.from_numpy(golden["fe_s64_patch"])
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(obs.shape[0], (64, 64), (16, 16), True, steps=1)
    from roboticattack_b200.config import NORM_MEAN, NORM_STD
    y = ofe.apply_patch_batch(obs, patch, xy[0], theta[0], ofe.MODE_WARP, NORM_MEAN, NORM_STD)
    np.testing.assert_allclose(y.numpy(), golden["fe_s64_warp_out"], rtol=0, atol=1e-6)
    # multi-step draw == consecutive single draws
    random.seed(1)
    np.random.seed(1)
    a = draw_placements(4, (224, 224), (50, 50), True, steps=3)
    random.seed(1)
    np.random.seed(1)
    b = [ofe.draw_placements(4, (224, 224), (50, 50), True) for _ in range(3)]
    assert np.array_equal(a[0], np.stack([x[0] for x in b])) and np.array_equal(a[1], np.stack([x[1] for x in b]))


def test_synthetic_batch_layout():
    b = synthetic_batch(openvla_7b(), 4, 33, ragged=True)
    assert b["obs"].shape == (4, 224, 224, 3) and b["obs"].dtype == torch.uint8
    for i in range(4):
        n = int(b["attention_mask"][i].sum())
        assert b["input_ids"][i, 0] == 1 and b["input_ids"][i, n - 1] == 2 and b["input_ids"][i, n - 9] == 29871
        assert (b["labels"][i, :n - 8] == -100).all() and (b["labels"][i, n:] == -100).all()
        assert ((b["labels"][i, n - 8:n - 1] >= 31744) & (b["labels"][i, n - 8:n - 1] < 32000)).all()
        assert (b["input_ids"][i, n:] == 32000).all()


def test_flop_model_matches_survey():
    f = flops_per_sample(openvla_7b(), 33)
    assert abs(f["iter"] / 1e9 - 8380.8) < 0.5 and abs(f["f_lin"] / 1e9 - 4136.17) < 0.1


def test_resume_reproduces_uninterrupted_run(tmp_path):
    """a run restarted from <save_dir>/last/attack_state.pt (patch, Adam moments, step counters and the three
    host RNG streams) ends with the same patch, bit for bit, as the uninterrupted run (CPU oracle engine standing in for the GPU)."""
    import os
    import random
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.attacker import UADAAttacker
    from roboticattack_b200.config import tiny
    from roboticattack_b200.synthetic import synthetic_batch
    from roboticattack_b200.weights import random_state_dict
    cfg = tiny(img=28, llm_layers=1, vit_depth=2)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    batches = []
    for i in range(4):
        b = synthetic_batch(cfg, 2, 14, seed=50 + i)
        batches.append({"pixel_values": b["obs"], "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
    kw = dict(num_iter=4, patch_size=[3, 8, 8], lr=2e-3, maskidx=[0, 1], warmup=1, geometry=True, innerLoop=2)

    def make(save_dir, resume=None):
        a = UADAAttacker(sd, save_dir=save_dir, optimizer="adamW", cfg=cfg, device="cpu", engine_factory=OracleEngine, resume=resume)
        a.val_every = 1
        return a

    def seed():
        random.seed(42)
        np.random.seed(42)
        torch.manual_seed(42)

    seed()
    full = make(str(tmp_path / "full")).patchattack_unconstrained(batches, None, **kw)
    seed()
    part_dir = str(tmp_path / "part")

    class Preempted(Exception):
        pass

    def crashing_loader():            # the job dies while fetching the third batch
        yield batches[0]
        yield batches[1]
        raise Preempted()

    with pytest.raises(Preempted):
        make(part_dir).patchattack_unconstrained(crashing_loader(), None, **kw)
    assert os.path.exists(os.path.join(part_dir, "last", "patch.pt")) and os.path.exists(os.path.join(part_dir, "last", "attack_state.pt"))
    random.seed(1), np.random.seed(1), torch.manual_seed(1)      # a restarted process has unrelated RNG state
    resumed = make(str(tmp_path / "resumed"), resume=os.path.join(part_dir, "last")).patchattack_unconstrained(batches[2:], None, **kw)
    assert torch.equal(full, resumed)
    assert not torch.equal(full, torch.load(os.path.join(part_dir, "last", "patch.pt")))


def test_resume_mid_accumulation_keeps_the_partial_gradient(tmp_path):
    """TMA / UPA with accumulate_steps > 1 step only on every accumulate_steps-th OUTER iteration (TMA.py:165); a checkpoint
    written in between must carry the partial gradient sum, or the resumed run silently drops it."""
    import os
    import random
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.attacker import TMAAttacker
    from roboticattack_b200.config import tiny
    from roboticattack_b200.synthetic import synthetic_batch
    from roboticattack_b200.weights import random_state_dict
    cfg = tiny(img=28, llm_layers=1, vit_depth=2)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    batches = []
    for i in range(4):
        b = synthetic_batch(cfg, 2, 14, seed=70 + i)
        batches.append({"pixel_values": b["obs"], "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
    kw = dict(num_iter=4, patch_size=[3, 8, 8], alpha=2e-3, maskidx=[0, 1, 2], warmup=0, geometry=True, innerLoop=1, accumulate_steps=2)

    def make(save_dir, resume=None):
        a = TMAAttacker(sd, save_dir=save_dir, optimizer="adamW", cfg=cfg, device="cpu", engine_factory=OracleEngine, resume=resume)
        a.val_every = 1
        return a

    def seed(v=42):
        random.seed(v), np.random.seed(v), torch.manual_seed(v)

    seed()
    full = make(str(tmp_path / "full")).patchattack_unconstrained(batches, None, **kw)

    class Preempted(Exception):
        pass

    def crashing_loader():            # dies while fetching the second batch: the state on disk is the accumulate-only iteration 0
        yield batches[0]
        raise Preempted()

    seed()
    part_dir = str(tmp_path / "part")
    with pytest.raises(Preempted):
        make(part_dir).patchattack_unconstrained(crashing_loader(), None, **kw)
    st = torch.load(os.path.join(part_dir, "last", "attack_state.pt"), weights_only=True)      # no pickled objects inside
    assert st["outer_iter"] == 0 and st["opt_step"] == 0 and st["accumulate"].abs().sum() > 0
    seed(1)
    resumed = make(str(tmp_path / "resumed"), resume=os.path.join(part_dir, "last")).patchattack_unconstrained(batches[1:], None, **kw)
    assert torch.equal(full, resumed)


def test_filter_train_matches_reference_branches():
    """filterGripTrainTo1 (UADA.py:311-341): 2..7 gripper-closed samples -> keep them; > 8 -> random.sample of 8; else unchanged."""
    import random
    from roboticattack_b200.attacker import _AttackerBase

    def batch(grips):
        B, T = len(grips), 12
        labels = torch.full((B, T), -100, dtype=torch.int64)
        for b, g in enumerate(grips):
            labels[b, T - 8:T - 1] = torch.tensor([31800] * 6 + [31744 if g else 31872])
            labels[b, T - 1] = 2
        return {"labels": labels, "input_ids": torch.arange(B * T).view(B, T), "attention_mask": torch.ones(B, T, dtype=torch.bool),
                "pixel_values": [f"img{b}" for b in range(B)]}

    d = batch([1, 0, 1, 1, 0])
    f = _AttackerBase.filter_train(d)
    assert f["pixel_values"] == ["img0", "img2", "img3"] and f["labels"].shape[0] == 3 and torch.equal(f["input_ids"], d["input_ids"][[0, 2, 3]])
    for grips in ([0, 0, 0], [1, 0, 0], [1] * 8):            # 0, 1 or exactly 8 hits: every branch of the reference falls through
        d = batch(grips)
        assert _AttackerBase.filter_train(d) is d
    d = batch([1] * 11 + [0])
    random.seed(5)
    f = _AttackerBase.filter_train(d)
    random.seed(5)
    assert f["pixel_values"] == [f"img{i}" for i in random.sample(list(range(11)), k=8)]


def test_reference_helper_methods(tmp_path):
    """The reference's helper methods under their own names (UADA.py:295-418, UPA.py:309-364, TMA.py:385-483), checked
    against the oracle / element-wise restatements of the reference loops."""
    import pickle
    from oracle import losses as ol
    from roboticattack_b200.attacker import TMAAttacker, UADAAttacker, UPAAttacker
    rng = np.random.default_rng(0)
    uada = UADAAttacker(None, save_dir=str(tmp_path), optimizer="adamW", device="cpu")
    tma = TMAAttacker(None, save_dir=str(tmp_path), optimizer="adamW", device="cpu")
    upa = UPAAttacker(None, save_dir=str(tmp_path), optimizer="adamW", alpha=0.8, belta=0.2, device="cpu")

    pred_ids, gt_ids = torch.from_numpy(rng.integers(31744, 32000, 21)), torch.from_numpy(rng.integers(31744, 32000, 21))
    np.testing.assert_allclose(float(uada.cal_UAD(pred_ids, gt_ids)), float(ol.cal_uad(pred_ids, gt_ids)), rtol=1e-12)

    pred, gt = torch.from_numpy(rng.uniform(-1, 1, 12)), torch.from_numpy(rng.uniform(-1, 1, 12))
    maskidx = [0, 2, 5]
    rd = uada.calculate_relative_distance(pred, gt, maskidx, {str(k): [] for k in maskidx})
    p2, g2 = pred.view(4, 3), gt.view(4, 3)
    for j, dof in enumerate(maskidx):        # the reference's scalar loop (UADA.py:357-368)
        want = [float(abs(p2[i, j] - g2[i, j]) / max(1 - g2[i, j], g2[i, j] + 1)) for i in range(4)]
        np.testing.assert_allclose(rd[str(dof)], want, rtol=1e-12)
    np.testing.assert_allclose(float(tma.calculate_relative_distance_target(pred, gt)),
                               np.mean([float(abs(p - g) / max(1 - g, g + 1)) for p, g in zip(pred, gt)]), rtol=1e-12)

    gt_g = torch.tensor([31872, 31872, 31744, 31744, 31900, 31800, 31872])
    pr_g = torch.tensor([31872, 31744, 31744, 31872, 31872, 31800, 31999])
    assert tma.calculate_01_ASR(pr_g, gt_g) == (2, 3, 1, 2, 1, 2)

    labels = torch.full((2, 10), -100, dtype=torch.int64)
    labels[0, 3:] = torch.arange(31750, 31757)
    labels[1, 2:9] = torch.arange(31760, 31767)
    out = tma.modifiy_labels(labels.clone(), {"0": 31999, "1": -100, "6": 31744})
    assert out[0, 3] == 31999 and out[0, 4] == 31751 and out[0, 9] == 31744
    assert out[1, 2] == 31999 and out[1, 3] == 31761 and out[1, 8] == 31744 and (out[1, :2] == -100).all()

    g = torch.tensor([[-100, 31872, 31900, 31800, 2]])
    torch.manual_seed(0)
    assert (upa.change_target(g.clone())[0, 1:] == 31999).all()     # the literal behaviour of UPA.py:358-364 (labels.change_target)

    uada.train_CE_loss, uada.val_UAD = [1.0, 2.0], [0.5]
    uada.train_MSE_distance_loss = uada.train_UAD = uada.val_CE_loss = uada.val_MSE_Distance = []
    uada.save_info(str(tmp_path / "info"))
    with open(tmp_path / "info" / "train_CE_loss.pkl", "rb") as f:
        assert pickle.load(f) == [1.0, 2.0]
    assert sorted(os.listdir(tmp_path / "info")) == sorted(f"{n}.pkl" for n in UADAAttacker.SAVE_INFO_LISTS)
    with pytest.raises(Exception):
        uada.weighted_loss(torch.zeros(1, 20, 32064), torch.full((1, 4), -100))   # CPU logits: no CPU fallback


def test_action_tokenizer_class(golden):
    """``ActionTokenizer`` (prismatic/vla/action_tokenizer.py:28-72) with and without a tokenizer object."""
    tok = lab.ActionTokenizer()
    assert tok.action_token_begin_idx == 31743 and tok.vocab_size == 256
    a = np.linspace(-1.2, 1.2, 41)
    np.testing.assert_array_equal(tok(a), ol.encode_actions_to_token_ids(a))
    ids = np.arange(31740, 32003)
    np.testing.assert_array_equal(tok.decode_token_ids_to_actions(ids), ol.decode_token_ids_to_actions(ids))

    class Tk:
        vocab_size = 32000
        def decode(self, ids):
            return " ".join(str(int(i)) for i in ids)
        def batch_decode(self, rows):
            return [self.decode(r) for r in rows]
    t2 = lab.ActionTokenizer(Tk())
    assert t2(np.zeros(7)) == " ".join(["31872"] * 7)
    assert t2(np.zeros((2, 3))) == ["31872 31872 31872"] * 2


@pytest.mark.parametrize("kind,opt", [("UADA", "adamW"), ("UPA", "adamW"), ("TMA", "adamW"), ("TMA", "pgd")])
def test_attacker_loops_on_the_cpu_oracle_engine(tmp_path, kind, opt):
    """Host side of the three attacks (outer loop, label preparation, schedule, validation, best-patch selection, side
    effects: UADA.py:104-292, UPA.py:104-275, TMA.py:93-383) driven end to end with the CPU oracle engine standing in for
    the CUDA engine -- the same flow tests/test_attackers_gpu.py runs on the GPU."""
    import importlib
    import pickle
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.weights import random_state_dict
    cfg = tiny(img=28, llm_layers=1, vit_depth=2)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")

    def loader(n, seed):
        out = []
        for i in range(n):
            b = synthetic_batch(cfg, 2, 14, seed=seed + i, ragged=(i % 2 == 1))
            out.append({"pixel_values": b["obs"], "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
        return out

    mod = importlib.import_module(f"roboticattack_b200.white_patch.{kind}")
    random.seed(42), np.random.seed(42), torch.manual_seed(42)
    kw = dict(alpha=0.8, belta=0.2) if kind == "UPA" else {}
    att = mod.OpenVLAAttacker(sd, None, save_dir=str(tmp_path), optimizer=opt, resize_patch=False, cfg=cfg, device="cpu",
                              engine_factory=OracleEngine, **kw)
    att.val_batches = 2
    common = dict(num_iter=2, target_action=np.zeros(7), patch_size=[3, 8, 8], accumulate_steps=1, maskidx=[0, 1, 2], warmup=1,
                  filterGripTrainTo1=False, geometry=True, innerLoop=2, args=None)
    step_arg = dict(alpha=2e-3) if kind == "TMA" else dict(lr=2e-3)
    patch = att.patchattack_unconstrained(loader(2, 100), loader(2, 200), **step_arg, **common)
    assert patch.shape == (3, 8, 8) and patch.dtype == torch.float32 and 0.0 <= patch.min() and patch.max() <= 1.0
    saved = torch.load(os.path.join(tmp_path, "last", "patch.pt"), weights_only=True)
    assert saved.shape == (3, 8, 8) and os.path.exists(os.path.join(tmp_path, "last", "attack_state.pt"))
    assert os.path.exists(os.path.join(tmp_path, "0", "patch.pt")), "the first validation always improves on the initial best"
    assert att.host.opt_step == 4 and len(att.train_CE_loss) >= 2 and all(np.isfinite(att.train_CE_loss))
    names = {"UADA": ["val_MSE_Distance", "val_UAD"], "UPA": ["avg_reserve_loss", "avg_angle_loss"], "TMA": ["val_L1_loss", "val_ASR"]}[kind]
    for name in names:
        with open(os.path.join(tmp_path, f"{name}.pkl"), "rb") as f:
            vals = pickle.load(f)
        assert len(vals) == 1 and np.isfinite(vals[0]), (name, vals)
    att.save_info(str(tmp_path / "info"))
    assert sorted(os.listdir(tmp_path / "info")) == sorted(f"{n}.pkl" for n in att.SAVE_INFO_LISTS)


def test_predict_action_postprocessing_matches_reference_method(monkeypatch):
    """``ActionPolicy.predict_action`` (the product's host code) against the reference's own
    ``OpenVLAForActionPrediction.predict_action`` run with a stand-in ``generate`` (tests/golden/make_golden_glue.py): the
    empty token 29871 is appended when missing, ids map to bin centres through ``32000 - id`` with clipping, and the
    result is un-normalised with q01 / q99 except where the mask is False."""
    from roboticattack_b200.policy import ActionPolicy
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_glue.npz"))
    stats = {"toy": {"action": {"q01": gold["pa_q01"].tolist(), "q99": gold["pa_q99"].tolist(), "mask": gold["pa_mask"].tolist()}}}
    pol = ActionPolicy(engine=None, norm_stats=stats)
    for i, toks in enumerate(gold["pa_tokens"]):
        seen = {}

        def fake(images_u8, input_ids, n_tokens=7, toks=toks, seen=seen):
            seen["ids"], seen["n"] = input_ids.numpy().copy(), n_tokens
            return toks[None]
        monkeypatch.setattr(pol, "generate_action_tokens", fake)
        act = pol.predict_action(np.zeros((28, 28, 3), dtype=np.uint8), torch.from_numpy(gold[f"pa_prompt{i}"]), unnorm_key="toy")
        np.testing.assert_allclose(act, gold["pa_actions"][i], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(seen["ids"], gold[f"pa_generate_input{i}"])
        assert seen["n"] == 7


@pytest.mark.parametrize("kind,opt", [("UADA", "adamW"), ("UPA", "adamW"), ("TMA", "adamW"), ("TMA", "pgd")])
def test_attack_loops_track_the_reference_loops(tmp_path, kind, opt):
    """The product's ``OpenVLAAttacker.patchattack_unconstrained`` (host loops of UADA / UPA / TMA: label preparation, host
    RNG protocol of the placements through training AND the validation pass (1000 / 100 / 100 batches), cosine schedule, the
    loss composition, AdamW step (+ UPA's L1 clip), clamp, validation metrics, best / last patch files) against the
    REFERENCE's own loops (UADA.py:93-292, UPA.py:92-275, TMA.py:82-383) run on the CPU with the reference's own model class
    and front end (tests/golden/make_golden_loop.py -> reference_golden_loop*.npz).  The product runs on the CPU oracle engine in
    bf16 like the reference model; what remains is bf16 round-off (the two sides order a few roundings differently; Adam's
    first steps are sign-like, so a single flipped sign moves a pixel by 2 lr).  TMA runs with geometry=False, the only
    branch of the reference that executes (TMA.py:137-141 passes ``colorjitter=`` to a method that does not take it)."""
    import argparse
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.config import LlamaConfig, OpenVLAConfig, ViTConfig
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gdir, "reference_golden_loop.npz"))                       # weights + batches (+ the UADA trajectory)
    suffix = ("" if kind == "UADA" else "_" + kind.lower()) + ("" if opt == "adamW" else "_" + opt)
    gk = g if suffix == "" else np.load(os.path.join(gdir, f"reference_golden_loop{suffix}.npz"))

    def big_matrix(rows, cols, a, b):       # embedding / lm_head by formula, as in the generator
        i, j = torch.arange(rows)[:, None], torch.arange(cols)[None, :]
        return (((i * a + j * b) % 257) - 128).float() / 4096.0

    sd = {k[2:]: torch.from_numpy(g[k]).view(torch.bfloat16).float() for k in g.files if k.startswith("w:")}
    sd["language_model.model.embed_tokens.weight"] = big_matrix(32064, 32, 131, 71)
    sd["language_model.lm_head.weight"] = big_matrix(32064, 32, 89, 153)
    cfg = OpenVLAConfig(dino=ViTConfig(dim=32, depth=3, heads=2, mlp_hidden=64, num_prefix=5, layerscale=True, img=28),
                        siglip=ViTConfig(dim=40, depth=3, heads=2, mlp_hidden=72, num_prefix=0, layerscale=False, img=28),
                        llm=LlamaConfig(hidden=32, layers=2, heads=2, ffn=64, vocab=32064), name="loop")

    def loader(name, n):
        return [{"pixel_values": torch.from_numpy(g[f"{name}{k}_obs"]), "input_ids": torch.from_numpy(g[f"{name}{k}_ids"]),
                 "attention_mask": torch.from_numpy(g[f"{name}{k}_mask"]), "labels": torch.from_numpy(g[f"{name}{k}_labels"])}
                for k in range(n)]

    lr = 2e-3
    mod = importlib.import_module(f"roboticattack_b200.white_patch.{kind}")
    kw = dict(alpha=0.8, belta=0.2) if kind == "UPA" else {}
    att = mod.OpenVLAAttacker(sd, None, save_dir=str(tmp_path), optimizer=opt, cfg=cfg, device="cpu",
                              engine_factory=lambda c, B, T, device="cpu": OracleEngine(c, B, T, device, dtype=torch.bfloat16), **kw)
    random.seed(42), np.random.seed(42), torch.manual_seed(42)
    step_arg = dict(alpha=lr) if kind == "TMA" else dict(lr=lr)
    patch = att.patchattack_unconstrained(loader("train", 3), loader("val", 2), num_iter=3, target_action=np.zeros(7), patch_size=[3, 8, 8],
                                          accumulate_steps=1, maskidx=[0, 1, 2], warmup=0, filterGripTrainTo1=False,
                                          geometry=(kind != "TMA"), innerLoop=2, args=argparse.Namespace(wandb_project="false"), **step_arg)
    tol = {"train_UAD": dict(atol=0.08), "val_UAD": dict(atol=5e-3),       # argmax flips of near-ties under bf16 noise
           "val_MSE_Distance": dict(rtol=1e-4), "val_CE_loss": dict(rtol=1e-4), "val_L1_loss": dict(rtol=1e-6), "val_ASR": dict(atol=0),
           "avg_angle_loss": dict(rtol=2e-3), "avg_distance_loss": dict(rtol=2e-3), "avg_reserve_loss": dict(rtol=2e-3)}
    checked = 0
    for name in gk.files:
        if name.startswith(("w:", "ft")) or name in ("grads", "patches", "saved_last", "saved_best", "final_patch") \
                or name[:3] in ("tra", "val") and name[-4:] in ("_obs", "_ids") or name.endswith("_mask") or name.endswith("_labels"):
            continue
        np.testing.assert_allclose([float(v) for v in getattr(att, name)], gk[name], err_msg=name, **tol.get(name, dict(rtol=1e-3)))
        checked += 1
    assert checked >= 4
    ref_final = torch.from_numpy(gk["final_patch"])
    # max: a pixel whose Adam / PGD step flips sign in two full-lr steps ends 4 lr away; mean: the bulk agrees to a fraction of lr
    assert (patch - ref_final).abs().max().item() < 4.5 * lr and (patch - ref_final).abs().mean().item() < lr / 10
    for sub, key in (("last", "saved_last"), ("0", "saved_best")):
        saved = torch.load(os.path.join(tmp_path, sub, "patch.pt"), weights_only=True)
        assert (saved - torch.from_numpy(gk[key])).abs().max().item() < 4.5 * lr
    assert (ref_final - torch.from_numpy(gk["saved_last"])).abs().max().item() > lr / 2, "the golden trajectory moves after iteration 0"


def test_lookahead_loader_order_restart_and_errors():
    """the next batch is fetched (and its PIL images converted) by ``prefetch()`` -- which the attack loops
    call between launching the inner steps and reading their scalars -- on the calling thread; order,
    restart-on-exhaustion and exception propagation are those of plain iteration."""
    from PIL import Image
    from roboticattack_b200.attacker import LookaheadLoader

    class Loader:
        def __init__(self, n, fail_at=None):
            self.n, self.fail_at, self.fetched = n, fail_at, []

        def __iter__(self):
            for k in range(self.n):
                if k == self.fail_at:
                    raise RuntimeError("loader died")
                self.fetched.append(k)
                yield {"pixel_values": [Image.fromarray(np.full((28, 28, 3), k, dtype=np.uint8))], "k": k}

    src = Loader(3)
    pf = LookaheadLoader(src, 28)
    assert src.fetched == [], "nothing is drawn from the loader (or its RNG) before the first next()"
    ks = []
    for i in range(7):                       # more than one pass: the loader restarts like the reference's while-loop
        b = pf.next()
        assert b["pixel_values"].dtype == torch.uint8 and tuple(b["pixel_values"].shape) == (1, 28, 28, 3)
        assert int(b["pixel_values"][0, 0, 0, 0]) == b["k"]
        ks.append(b["k"])
        n_before = len(src.fetched)
        pf.prefetch()                        # "while the device runs the inner loop"
        pf.prefetch()                        # idempotent until consumed
        assert len(src.fetched) == n_before + 1
    pf.close()
    assert ks == [0, 1, 2, 0, 1, 2, 0]
    pf = LookaheadLoader(Loader(3, fail_at=2), 28)
    assert pf.next()["k"] == 0
    pf.prefetch()
    assert pf.next()["k"] == 1
    pf.prefetch()                            # the failure is kept for the next() that would have produced the batch
    with pytest.raises(RuntimeError, match="loader died"):
        pf.next()
    pf = LookaheadLoader(Loader(2), 28, restart=False)
    assert [pf.next()["k"], pf.next()["k"]] == [0, 1]
    with pytest.raises(StopIteration):
        pf.next()
    off = Loader(3)
    pf = LookaheadLoader(off, 28, lookahead=False)
    pf.next()
    pf.prefetch()
    assert off.fetched == [0], "VLA_PREFETCH=0: batches are fetched only when asked for"


def test_seeded_run_is_reproducible_with_a_shuffled_torch_dataloader(tmp_path, monkeypatch):
    """A torch DataLoader draws its base seed and its RandomSampler seed from the GLOBAL torch RNG inside iter() / next().
    With the lookahead on the calling thread a seeded run (initial patch, batches, placements, final patch) is a pure
    function of the seeds -- with and without the lookahead."""
    import argparse
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.attacker import UPAAttacker
    from roboticattack_b200.config import tiny
    from roboticattack_b200.synthetic import synthetic_batch
    from roboticattack_b200.weights import random_state_dict
    cfg = tiny(img=28, llm_layers=1, vit_depth=2)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    big = synthetic_batch(cfg, 6, 14, seed=3)

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 6

        def __getitem__(self, i):
            return {"pixel_values": big["obs"][i], "input_ids": big["input_ids"][i], "attention_mask": big["attention_mask"][i],
                    "labels": big["labels"][i]}

    def run(prefetch):
        monkeypatch.setenv("VLA_PREFETCH", prefetch)
        random.seed(7)
        np.random.seed(7)
        torch.manual_seed(7)
        loader = torch.utils.data.DataLoader(DS(), batch_size=2, shuffle=True)
        vloader = torch.utils.data.DataLoader(DS(), batch_size=2, shuffle=True)
        a = UPAAttacker(sd, None, save_dir=str(tmp_path / prefetch), optimizer="adamW", alpha=0.8, belta=0.2, cfg=cfg, device="cpu",
                        engine_factory=OracleEngine)
        a.val_batches, a.val_every = 1, 2
        p = a.patchattack_unconstrained(loader, vloader, num_iter=4, patch_size=[3, 6, 6], lr=2e-3, maskidx=[0, 1, 2], warmup=0,
                                        geometry=True, innerLoop=2, args=argparse.Namespace(wandb_project="false"))
        return p, list(a.train_CE_loss)

    p1, l1 = run("1")
    p1b, l1b = run("1")
    p0, l0 = run("0")
    assert torch.equal(p1, p1b) and l1 == l1b, "seeded run is not reproducible"
    assert torch.equal(p1, p0) and l1 == l0, "the lookahead changed the RNG order"


def test_filter_train_matches_reference_method():
    """``filter_train`` against the reference's own method (UADA.py:309-340, called from tests/golden/make_golden_loop.py):
    2..7 closed-gripper samples are kept, more than 8 are sampled down to 8 with ``random.sample``, anything else passes."""
    from roboticattack_b200.attacker import _AttackerBase
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_loop.npz"))
    for c in range(4):
        grips = g[f"ft{c}_grips"]
        n, T2 = len(grips), 12
        labels = torch.full((n, T2), -100, dtype=torch.int64)
        for r, gv in enumerate(grips):
            labels[r, T2 - 8:T2 - 1] = torch.tensor([31800 + r] * 6 + [31744 if gv else 31872])
            labels[r, T2 - 1] = 2
        batch = {"labels": labels, "input_ids": torch.arange(n * T2).view(n, T2), "attention_mask": torch.ones(n, T2, dtype=torch.bool),
                 "pixel_values": list(range(n))}
        random.seed(5)
        out = _AttackerBase.filter_train(batch)
        np.testing.assert_array_equal(out["labels"].numpy(), g[f"ft{c}_labels"])
        np.testing.assert_array_equal(out["attention_mask"].numpy(), g[f"ft{c}_mask"])
        np.testing.assert_array_equal(out["input_ids"].numpy(), g[f"ft{c}_ids"])
        assert list(out["pixel_values"]) == g[f"ft{c}_pixels"].tolist()
