"""TEST INFRASTRUCTURE ONLY (never imported by the product path): greedy action decode on the oracle model, following
``predict_action`` (prismatic/extern/hf/modeling_prismatic.py:506-536) with ``generate(do_sample=False)`` unrolled as one
full forward per token."""
import numpy as np
import torch

from . import frontend as ofe, model as om


@torch.no_grad()
def greedy_action_tokens(sd, cfg, obs_u8, input_ids, n_tokens, mean, std, dtype=torch.float32):
    """-> (tokens int64 [B,n], margins float [B,n] = top-1 minus top-2 logit of every decision)."""
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    B = input_ids.shape[0]
    xy = np.zeros((B, 2), dtype=np.int32)
    theta = np.zeros((B, 2, 3), dtype=np.float32)
    px = ofe.apply_patch_batch(obs_u8, torch.zeros(3, 1, 1), xy, theta, ofe.MODE_NONE, mean, std).to(dtype)
    ids = input_ids.clone()
    toks, margins = [], []
    for _ in range(n_tokens):
        mask = torch.ones_like(ids, dtype=torch.bool)
        out = om.forward(sdd, cfg, ids, mask, px, None)
        last = out.logits[:, -1].float()
        top2 = last.topk(2, dim=1)
        toks.append(top2.indices[:, 0])
        margins.append(top2.values[:, 0] - top2.values[:, 1])
        ids = torch.cat([ids, top2.indices[:, :1]], dim=1)
    return torch.stack(toks, 1), torch.stack(margins, 1)
