"""Host-side logic of the product package, checked against the oracle and the golden vectors (CPU only)."""
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
