"""Synthetic bridge_orig-shaped inputs (no dataset is available offline); SURVEY.md section 8d.

Layout follows the reference's data path: RLDS resizes observations to 224x224 (white_patch/openvla_dataloader.py:122;
the reference's own DummyDataset uses uniform noise, prismatic/vla/datasets/datasets.py:218); token layout
``<s> prompt... 29871 a1..a7 </s>`` with labels -100 except the last 8 (datasets.py:56-69); right padding with
32000 / -100 / False (prismatic/util/data_utils.py:184-191).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from .config import IGNORE_INDEX, PAD_TOKEN_ID, OpenVLAConfig


def synthetic_batch(cfg: OpenVLAConfig, batch: int, text_len: int = 33, seed: int = 1234, ragged: bool = False):
    """-> dict(obs uint8 [B,H,W,3], input_ids i64 [B,T], attention_mask bool [B,T], labels i64 [B,T])."""
    assert text_len >= 11, "need BOS + >=1 prompt id + 29871 + 7 action ids + EOS"
    g = torch.Generator().manual_seed(seed)
    H = cfg.img
    obs = torch.randint(0, 256, (batch, H, H, 3), generator=g, dtype=torch.uint8)
    ids = torch.full((batch, text_len), PAD_TOKEN_ID, dtype=torch.int64)
    labels = torch.full((batch, text_len), IGNORE_INDEX, dtype=torch.int64)
    mask = torch.zeros(batch, text_len, dtype=torch.bool)
    for b in range(batch):
        n = text_len - (int(torch.randint(0, 6, (1,), generator=g)) if (ragged and b > 0) else 0)
        n = max(n, 11)
        prompt = torch.randint(3, 31744, (n - 10,), generator=g)
        actions = torch.randint(31744, 32000, (7,), generator=g)
        row = torch.cat([torch.tensor([1]), prompt, torch.tensor([29871]), actions, torch.tensor([2])])
        ids[b, :n] = row
        mask[b, :n] = True
        labels[b, n - 8:n] = row[n - 8:]
    return {"obs": obs, "input_ids": ids, "attention_mask": mask, "labels": labels}


def draw_placements(batch, img_hw, patch_hw, geometry, steps=1):
    """Host RNG protocol of ``RandomPatchTransform.apply_random_patch_batch`` (appply_random_transform.py:104-136),
    drawn for ``steps`` consecutive calls: per image ``x = random.randint(0, W-pw)``, ``y = random.randint(0, H-ph)``,
    then, if ``geometry``: ``np.random.rand() < 0.2`` -> identity, else angle ~ U(-30,30), shx, shy ~ U(-0.2,0.2) and
    theta = (S.R)[:2] in float32 (:80-91, :26-41).  Returns xy int32 [steps,B,2], theta float32 [steps,B,2,3]."""
    H, W = img_hw
    ph, pw = patch_hw
    xy = np.zeros((steps, batch, 2), dtype=np.int32)
    theta = np.zeros((steps, batch, 2, 3), dtype=np.float32)
    for s in range(steps):
        for b in range(batch):
            x = random.randint(0, W - pw)
            y = random.randint(0, H - ph)
            xy[s, b] = (x, y)
            m = np.eye(3, dtype=np.float32)
            if geometry and not (np.random.rand() < 0.2):
                angle = np.random.uniform(-30, 30)
                shx = np.random.uniform(-0.2, 0.2)
                shy = np.random.uniform(-0.2, 0.2)
                t = np.deg2rad(angle)
                c, sn = np.cos(t), np.sin(t)
                R = np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]], dtype=np.float32)
                S = np.array([[1, shx, 0], [shy, 1, 0], [0, 0, 1]], dtype=np.float32)
                m = np.dot(S, R)
            theta[s, b] = m[:2, :]
    return xy, theta
