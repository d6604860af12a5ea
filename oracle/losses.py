"""Label masking, loss heads and metrics of the attack engines, restated.

Follows: ``mask_labels`` UADA.py:371-379 (= UADA_ddp.py:89-97) and UPA.py:344-356; ``weighted_loss``
UADA.py:381-406 / UADA_ddp.py:99-124 / UPA.py:367-387; ``cal_UAD`` UADA.py:408-418; TMA target substitution
TMA.py:93-100,124-129; ``ActionTokenizer`` prismatic/vla/action_tokenizer.py:28-68;
``calculate_relative_distance`` UADA.py:355-369.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from roboticattack_b200.config import ACTION_TOKEN_BEGIN_IDX, VOCAB_TOKENIZER

BINS = np.linspace(-1, 1, 256)                    # action_tokenizer.py:30
BIN_CENTERS = (BINS[:-1] + BINS[1:]) / 2.0         # :31


def decode_token_ids_to_actions(ids: np.ndarray) -> np.ndarray:   # action_tokenizer.py:49-68
    d = VOCAB_TOKENIZER - ids
    d = np.clip(d - 1, a_min=0, a_max=BIN_CENTERS.shape[0] - 1)
    return BIN_CENTERS[d]


def encode_actions_to_token_ids(action: np.ndarray) -> np.ndarray:   # :38-47 (ids before tokenizer.decode)
    action = np.clip(action, a_min=-1.0, a_max=1.0)
    return VOCAB_TOKENIZER - np.digitize(action, BINS)


def mask_labels_uada(labels, maskidx):   # UADA.py:371-379 -- in place, like the reference
    mask = labels > ACTION_TOKEN_BEGIN_IDX
    masked = labels[mask]
    masked = masked.view(masked.shape[0] // 7, 7)
    template = torch.ones_like(masked) * -100
    for idx in maskidx:
        template[:, idx] = masked[:, idx]
    labels[labels > 2] = template.view(-1)
    return labels


def mask_labels_upa(labels, maskidx):    # UPA.py:344-356
    mask = labels > ACTION_TOKEN_BEGIN_IDX
    masked = labels[mask]
    masked = masked.view(masked.shape[0] // 7, 7)
    for idx in range(7):
        if idx not in maskidx:
            masked[:, idx] = -100
    new = []
    for j in range(labels.shape[0]):
        t = labels[j]
        t[t > 2] = masked[j]
        new.append(t.unsqueeze(0))
    return torch.cat(new, dim=0)


def tma_target(target_action_ids, maskidx):   # TMA.py:93-100: 7 action ids + EOS, entries not in maskidx -> -100
    t = list(int(v) for v in target_action_ids) + [2]
    t = torch.tensor(t)
    for idx in range(len(t)):
        if idx not in maskidx:
            t[idx] = -100
    return t


def tma_labels(labels, target):   # TMA.py:124-129
    new = []
    for j in range(labels.shape[0]):
        t = labels[j].clone()
        t[t != -100] = target
        new.append(t.unsqueeze(0))
    return torch.cat(new, dim=0)


def cal_uad(pred_ids, gt_ids):   # UADA.py:408-418
    gt = torch.tensor(decode_token_ids_to_actions(gt_ids.clone().detach().cpu().numpy()))
    pred = torch.tensor(decode_token_ids_to_actions(pred_ids.clone().detach().cpu().numpy()))
    max_distance = torch.where(gt > 0, torch.abs(gt - (-1)), torch.abs(gt - 1))
    return (torch.abs(pred - gt) / max_distance).mean()


def weighted_loss_uada(logits, labels, mse_weights=5):   # UADA.py:381-396 (weights 5), UADA_ddp.py:99-114
    temp_label = labels[:, 1:]
    action_mask = temp_label > 2
    temp_logits = logits[:, :, 31744:32000]
    action_logits = temp_logits[:, -temp_label.shape[-1] - 1:-1, :]
    action_logits = action_logits[action_mask]
    reweigh = torch.arange(1, 257) / 256
    temp_prob = F.softmax(action_logits, dim=-1)
    reweighted_prob = (temp_prob * reweigh).sum(dim=-1)
    hard = temp_label[action_mask]
    hard[hard > 31872] = 31999
    hard[hard <= 31872] = 31744
    hard[hard == 31999] = 1 / 256      # int64 tensor: truncates to 0 (reproduced on purpose)
    hard[hard == 31744] = 1
    uad = cal_uad(action_logits.argmax(dim=-1) + 31744, temp_label[action_mask])
    loss = F.mse_loss(mse_weights * reweighted_prob.contiguous(), mse_weights * hard.float().contiguous())
    return loss, uad


def weighted_loss_upa(logits, labels, alpha, belta, num_patches):   # UPA.py:367-387
    temp_label = labels[:, 1:]
    action_mask = temp_label != -100
    temp_logits = logits[:, :, 31744:32000]
    action_logits = temp_logits[:, num_patches:-1]
    reweigh = torch.arange(1, 257)
    temp_prob = F.softmax(action_logits, dim=-1)
    reweighted_prob = (temp_prob * reweigh).sum(dim=-1)
    xyz_r = torch.cat([row[action_mask[i]].unsqueeze(0) for i, row in enumerate(reweighted_prob)], dim=0)[:, :3]
    xyz_l = (torch.cat([row[action_mask[i]].unsqueeze(0) for i, row in enumerate(temp_label)], dim=0) - 31743)[:, :3]
    xyz_r = (xyz_r - 1) / 255
    xyz_l = (xyz_l - 1) / 255
    cos = F.cosine_similarity(xyz_r, xyz_l, dim=1)
    angle_loss = (cos + 1).mean()
    distance_loss = 1 / (torch.norm(xyz_r - xyz_l, p=2, dim=1).mean() + 1e-3)
    return alpha * angle_loss + belta * distance_loss, angle_loss, distance_loss


def relative_distance(pred, gt):   # UADA.py:355-369, one value per (sample, dof)
    dist_up, dist_lo = 1 - gt, gt - (-1)
    return (pred - gt).abs() / torch.maximum(dist_up, dist_lo)
