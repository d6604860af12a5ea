"""Patch update rules, restated.

``transformers.AdamW`` (pinned 4.40.1, removed from the installed 5.5.0; published algorithm restated):
defaults betas=(0.9, 0.999), eps=1e-6, weight_decay=0, correct_bias=True;
``exp_avg.mul_(b1).add_(g, alpha=1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2);
denom = exp_avg_sq.sqrt().add_(eps); step_size = lr*sqrt(1-b2^t)/(1-b1^t); p.addcdiv_(exp_avg, denom, value=-step_size)``
-- eps is added BEFORE the bias correction (unlike torch.optim.AdamW).  Used at UADA.py:107-115,155-157,
UADA_ddp.py:167-174,208-209, UPA.py:157-160 (after an L1 grad-norm clip to 1e-3), TMA.py:164-170.
``get_cosine_schedule_with_warmup`` (num_cycles=0.5) is stepped once per OUTER iteration (UADA.py:162-164).
sign-PGD: TMA.py:171-175.
"""
from __future__ import annotations

import math

import torch


class HFAdamW:
    def __init__(self, shape, lr, betas=(0.9, 0.999), eps=1e-6):
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = torch.zeros(shape)
        self.v = torch.zeros(shape)
        self.t = 0

    def step(self, p, g, lr=None):
        lr = self.lr if lr is None else lr
        b1, b2 = self.betas
        self.t += 1
        self.m.mul_(b1).add_(g, alpha=1.0 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = self.v.sqrt().add_(self.eps)
        step_size = lr * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        p.addcdiv_(self.m, denom, value=-step_size)
        return p


def cosine_with_warmup_lambda(step, warmup, total, num_cycles=0.5):
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def clip_grad_l1_(g, max_norm):   # torch.nn.utils.clip_grad_norm_([patch], max_norm, norm_type=1), UPA.py:157
    total = g.abs().sum()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    g.mul_(coef)
    return total


def pgd_step(p, g, alpha):        # TMA.py:173
    return (p - alpha * g.sign()).clamp(0, 1)
