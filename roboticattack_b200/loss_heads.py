"""The reference's ``weighted_loss`` methods on caller-supplied logits, computed by the CUDA loss-head kernel.

Inside the attack loop the loss heads run fused behind the lm_head GEMM (``vla_fwd_bwd``); these wrappers expose the same
kernel (``vla_loss_head``, include/vla_b200.h) under the reference's calling convention -- full logits ``[B, L, V]`` of the
multimodal sequence and text labels ``[B, T]`` -- with autograd back to the logits, for code that drives its own model:

* ``uada_weighted_loss``  -- UADA.py:381-406 / UADA_ddp.py:99-124 (returns the MSE distance and the UAD metric)
* ``upa_weighted_loss``   -- UPA.py:367-387 (returns total, angle and distance terms)

CUDA only: a CPU tensor raises (the CPU restatement of these functions lives in ``oracle/`` and is test infrastructure).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .config import IGNORE_INDEX


def supervised_rows(labels: torch.Tensor, seq_len: int):
    """Rows of the flattened ``[B * seq_len, V]`` logits that predict a supervised label, in (sample, position) order, and
    the kernel's ``meta`` table {label, sample, index among the sample's supervised rows}.  The label at text position t
    is predicted by the logits at ``num_patches + t - 1`` (the shift of ``logits[:, -T:-1]`` in UADA.py:385)."""
    lab = labels.detach().cpu()
    B, T = lab.shape
    P = seq_len - T
    sup = lab[:, 1:] != IGNORE_INDEX
    b_idx, t_idx = sup.nonzero(as_tuple=True)
    rows = b_idx * seq_len + P + t_idx
    idx = sup.cumsum(dim=1)[b_idx, t_idx] - 1
    meta = torch.stack([lab[b_idx, t_idx + 1], b_idx, idx], dim=1).to(torch.int32)
    return rows, meta


class _LossHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, kind, mse_weight, alpha, belta, ce_scale):
        if not logits.is_cuda:
            raise _lib.VLAError("weighted_loss: the loss heads run on the GPU; pass CUDA logits")
        B, Lm, V = logits.shape
        rows, meta = supervised_rows(labels, Lm)
        R = int(rows.numel())
        if R == 0:
            raise _lib.VLAError("weighted_loss: no supervised label in the batch")
        rows = rows.to(logits.device)
        z = logits.reshape(B * Lm, V)[rows].float().contiguous()
        meta = meta.to(logits.device).contiguous()
        stats = torch.empty(8 * R, device=logits.device)
        dz = torch.empty(R, V, device=logits.device, dtype=torch.bfloat16)
        scalars = torch.zeros(_lib.NUM_SCALARS, device=logits.device)
        pred = torch.empty(R, dtype=torch.int32, device=logits.device)
        lp = _lib.LossParams(kind, float(mse_weight), float(alpha), float(belta), float(ce_scale))
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().vla_loss_head(_lib.ptr(z), _lib.ptr(meta), R, V, B, ctypes.byref(lp), _lib.ptr(stats), _lib.ptr(dz),
                                               _lib.ptr(scalars), _lib.ptr(pred), _lib.cur_stream()), "vla_loss_head")
        ctx.save_for_backward(rows, dz)
        ctx.shape, ctx.dtype = (B, Lm, V), logits.dtype
        ctx.mark_non_differentiable(scalars, pred)
        return scalars[_lib.S_LOSS].clone(), scalars, pred

    @staticmethod
    def backward(ctx, gloss, _gs, _gp):
        rows, dz = ctx.saved_tensors
        B, Lm, V = ctx.shape
        g = torch.zeros(B * Lm, V, device=dz.device, dtype=ctx.dtype)
        g[rows] = (dz.float() * gloss).to(ctx.dtype)
        return g.view(B, Lm, V), None, None, None, None, None, None


def uada_weighted_loss(logits, labels, mse_weight=5.0):
    """``(distance_loss, UAD)`` of UADA.py:381-406 with the weight 5 of the shipped code (``MSE_weights`` in UADA_ddp.py)."""
    loss, scalars, _ = _LossHead.apply(logits, labels, _lib.LOSS_UADA_DDP, mse_weight, 0.0, 0.0, 1.0)
    return loss, scalars[_lib.S_UAD]


def upa_weighted_loss(logits, labels, alpha, belta):
    """``(total_loss, angle_loss, distance_loss)`` of UPA.py:367-387; the two terms are Python floats like the reference's."""
    loss, scalars, _ = _LossHead.apply(logits, labels, _lib.LOSS_UPA, 0.0, alpha, belta, 1.0)
    s = scalars.cpu()
    return loss, s[_lib.S_AUX0].item(), s[_lib.S_AUX1].item()
