"""Host-side label preparation and metrics of the attack engines (run once per OUTER iteration, on CPU tensors).

Reference: ``mask_labels`` UADA.py:371-379 (= UADA_ddp.py:89-97) and UPA.py:344-356; TMA target substitution
TMA.py:93-100,124-129; ``change_target`` UPA.py:358-364; ``ActionTokenizer`` prismatic/vla/action_tokenizer.py:28-68;
``calculate_relative_distance`` UADA.py:355-369; ``calculate_relative_distance_target`` (TMA).
"""
from __future__ import annotations

import numpy as np
import torch

from .config import ACTION_TOKEN_BEGIN_IDX, IGNORE_INDEX, VOCAB_TOKENIZER

_BINS = np.linspace(-1, 1, 256)
_BIN_CENTERS = (_BINS[:-1] + _BINS[1:]) / 2.0


def decode_token_ids_to_actions(ids) -> np.ndarray:
    d = VOCAB_TOKENIZER - np.asarray(ids)
    d = np.clip(d - 1, a_min=0, a_max=_BIN_CENTERS.shape[0] - 1)
    return _BIN_CENTERS[d]


def action_to_token_ids(action) -> np.ndarray:
    """ids ActionTokenizer.__call__ encodes before ``tokenizer.decode`` (the tokenizer round trip of TMA.py:93 maps
    them back to themselves)."""
    action = np.clip(np.asarray(action, dtype=np.float64), a_min=-1.0, a_max=1.0)
    return VOCAB_TOKENIZER - np.digitize(action, _BINS)


class ActionTokenizer:
    """Interface of ``prismatic/vla/action_tokenizer.py:28-72`` on the two functions above: uniform bins over
    [min_action, max_action], ids counted down from the end of the tokenizer vocabulary.  ``tokenizer`` may be None (the
    attack engines only need ids): ``__call__`` then returns the token ids instead of the decoded string."""

    def __init__(self, tokenizer=None, bins: int = 256, min_action: int = -1, max_action: int = 1):
        self.tokenizer, self.n_bins, self.min_action, self.max_action = tokenizer, bins, min_action, max_action
        self.bins = np.linspace(min_action, max_action, bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        self._vocab = int(getattr(tokenizer, "vocab_size", VOCAB_TOKENIZER))
        self.action_token_begin_idx = int(self._vocab - (bins + 1))

    @property
    def vocab_size(self) -> int:
        return self.n_bins

    def token_ids(self, action) -> np.ndarray:
        action = np.clip(np.asarray(action, dtype=np.float64), a_min=float(self.min_action), a_max=float(self.max_action))
        return self._vocab - np.digitize(action, self.bins)

    def __call__(self, action):
        ids = self.token_ids(action)
        if self.tokenizer is None:
            return ids
        return self.tokenizer.decode(list(ids)) if ids.ndim == 1 else self.tokenizer.batch_decode(ids.tolist())

    def decode_token_ids_to_actions(self, action_token_ids) -> np.ndarray:
        d = self._vocab - np.asarray(action_token_ids)
        return self.bin_centers[np.clip(d - 1, a_min=0, a_max=self.bin_centers.shape[0] - 1)]


def _action_rows(labels):
    mask = labels > ACTION_TOKEN_BEGIN_IDX
    acts = labels[mask]
    if acts.numel() % 7 != 0:
        raise ValueError(f"expected 7 action tokens per sample, found {acts.numel()} in {labels.shape[0]} samples")
    return mask, acts.view(-1, 7)


def mask_labels_uada(labels: torch.Tensor, maskidx) -> torch.Tensor:
    """Keep the action labels of the DoF in ``maskidx`` (EOS stays supervised), others -> -100; in place."""
    _, acts = _action_rows(labels)
    template = torch.full_like(acts, IGNORE_INDEX)
    for idx in maskidx:
        template[:, idx] = acts[:, idx]
    labels[labels > 2] = template.view(-1)
    return labels


def mask_labels_upa(labels: torch.Tensor, maskidx) -> torch.Tensor:
    _, acts = _action_rows(labels)
    acts = acts.clone()
    for idx in range(7):
        if idx not in maskidx:
            acts[:, idx] = IGNORE_INDEX
    for j in range(labels.shape[0]):
        row = labels[j]
        row[row > 2] = acts[j]
    return labels


def tma_target(target_action, maskidx) -> torch.Tensor:
    """8-vector: ids of the 7 target actions + EOS, entries whose index is not in ``maskidx`` -> -100 (TMA.py:93-100)."""
    t = torch.tensor(list(action_to_token_ids(target_action).astype(np.int64)) + [2], dtype=torch.int64)
    for idx in range(len(t)):
        if idx not in maskidx:
            t[idx] = IGNORE_INDEX
    return t


def tma_labels(labels: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    new = labels.clone()
    for j in range(labels.shape[0]):
        row = new[j]
        sel = row != IGNORE_INDEX
        if int(sel.sum()) != target.numel():
            raise ValueError("TMA expects 8 supervised tokens per sample")
        row[sel] = target
    return new


def change_target(gt: torch.Tensor) -> torch.Tensor:
    """UPA ``guide`` labels (UPA.py:358-364), statement by statement.  The reference mutates ``gt`` between its
    three masked assignments, so every supervised token (EOS included) ends up as 31999; that literal behaviour is
    what is reproduced here (``guide`` defaults to False and no shipped script enables it)."""
    mask = gt != IGNORE_INDEX
    zero = mask & (gt == 31872)
    rnd = torch.randint(0, 2, (int(zero.sum()),), dtype=torch.bool)
    gt[zero] = torch.where(rnd, torch.tensor(31744), torch.tensor(31999))
    gt[mask & (gt > 31872)] = 31744
    gt[mask & (gt < 31872)] = 31999
    return gt


def relative_distance(pred_actions: torch.Tensor, gt_actions: torch.Tensor) -> torch.Tensor:
    """|pred - gt| / max(1 - gt, gt + 1) elementwise (UADA.py:355-369)."""
    return (pred_actions - gt_actions).abs() / torch.maximum(1 - gt_actions, gt_actions + 1)
