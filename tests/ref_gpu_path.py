"""Times the reference-STYLE GPU path on this box (last row): the oracle restatement of the reference's eager
PyTorch attack iteration (per-image front-end loop, autograd, full-sequence lm_head + fp32 logits, weighted_loss, HF-AdamW,
clamp) run on cuda:0 in bf16 with torch's own kernels (cuBLAS, SDPA) -- the denominator for "x times the reference GPU
path".  Two variants, as in the reference: weights frozen (UADA_ddp.py:50-51) and weights requiring grad (UADA.py never
freezes them, so autograd also computes 7.5 B unused weight gradients).  Measurement tool only; not on the product path.
usage: python tests/ref_gpu_path.py [--batch 8] [--iters 6]"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import frontend as ofe, losses as ol, model as om, optim as oo
from roboticattack_b200.config import NORM_MEAN, NORM_STD, openvla_7b
from roboticattack_b200.synthetic import draw_placements, synthetic_batch
from roboticattack_b200.weights import random_state_dict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--patch", type=int, default=50)
    ap.add_argument("--text-len", type=int, default=33)
    a = ap.parse_args()
    cfg = openvla_7b()
    dev = torch.device("cuda:0")
    sd = random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16, init="reference")
    b = synthetic_batch(cfg, a.batch, a.text_len, seed=1234)
    labels = ol.mask_labels_uada(b["labels"].clone(), [0, 1, 2]).to(dev)
    torch.set_default_device(dev)      # the oracle builds its masks / tables with factory functions
    ids, mask, obs = b["input_ids"].to(dev), b["attention_mask"].to(dev), b["obs"].to(dev)
    for frozen in (True, False):
        for t in sd.values():
            t.requires_grad_(not frozen and t.is_floating_point())
        torch.manual_seed(42)
        patch = torch.rand(3, a.patch, a.patch, device=dev)
        opt = oo.HFAdamW(patch.shape, 2e-3)
        random.seed(42)
        np.random.seed(42)
        ts = []
        for it in range(a.iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            xy, th = draw_placements(a.batch, (cfg.img, cfg.img), (a.patch, a.patch), True)
            p = patch.clone().requires_grad_(True)
            px = ofe.apply_patch_batch(obs, p, xy[0], th[0], ofe.MODE_WARP, NORM_MEAN, NORM_STD)
            out = om.forward(sd, cfg, ids, mask, px.to(torch.bfloat16), labels)
            mse, _ = ol.weighted_loss_uada(out.logits, labels, 5)
            loss = mse + 1 / out.loss
            loss.backward()
            opt.step(patch, p.grad)
            patch.clamp_(0, 1)
            _ = loss.item()
            if not frozen:
                for t in sd.values():
                    t.grad = None
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        best = min(ts[2:])
        print(f"reference-style eager GPU path, bs={a.batch}, weights {'frozen (UADA_ddp.py)' if frozen else 'requiring grad (UADA.py)'}: "
              f"{best * 1e3:.1f} ms/iteration = {1 / best:.2f} it/s   (all iterations: {[round(x * 1e3, 1) for x in ts]})", flush=True)


if __name__ == "__main__":
    main()
