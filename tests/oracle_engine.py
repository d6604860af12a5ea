"""CPU stand-in for VLAEngine built on the oracle (tests only): lets the host-side attack loops and the multi-process
(gloo) data-parallel plumbing run without a GPU."""
import math

import numpy as np
import torch

from oracle import frontend as ofe, losses as ol, model as om
from roboticattack_b200 import _lib
from roboticattack_b200.config import NORM_MEAN, NORM_STD


class GlooComm:
    """Stand-in for the NCCL communicator on CPU: sum over the ranks of the default (gloo) process group."""

    def __init__(self, world):
        self.world = world

    def all_reduce_(self, t):
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)


class OracleEngine:
    def __init__(self, cfg, batch, text_len, device="cpu", dtype=torch.float32):
        self.cfg, self.B, self.T, self.dtype = cfg, batch, text_len, dtype
        self.sd = None
        self.num_supervised = 0
        self._place, self._adam = 0, 0

    def load_state_dict(self, sd, strict=True):
        self.sd = {k: v.to(self.dtype) for k, v in sd.items()}

    def ensure_plan(self, B, T):
        self.B, self.T = B, T

    def set_batch(self, obs, ids, mask, labels):
        self.batch = (obs.cpu(), ids.clone(), mask.clone(), labels.clone())
        self.num_supervised = int((labels[:, 1:] != -100).sum())
        return self.num_supervised

    def set_placements(self, xy, theta):
        self.xy, self.theta = np.asarray(xy), np.asarray(theta)

    def fwd_bwd(self, patch, step_idx, fe_mode, loss, dpatch, scalars, pred_ids, forward_only=False):
        obs, ids, mask, labels = self.batch
        p = patch.detach().clone().requires_grad_(True)
        px = ofe.apply_patch_batch(obs, p, self.xy[step_idx], self.theta[step_idx], fe_mode, NORM_MEAN, NORM_STD)
        self._px = px.detach()
        out = om.forward(self.sd, self.cfg, ids, mask, px.to(self.dtype), labels)
        logits = out.logits.float()
        aux0 = aux1 = uad = 0.0
        if loss.kind in (_lib.LOSS_UADA, _lib.LOSS_UADA_DDP):
            mse, uad = ol.weighted_loss_uada(logits, labels, loss.mse_weight)
            total = mse + 1 / out.loss if loss.kind == _lib.LOSS_UADA else mse
            aux0, uad = mse.item(), float(uad)
        elif loss.kind == _lib.LOSS_UPA:
            total, a, d = ol.weighted_loss_upa(logits, labels, loss.alpha, loss.belta, self.cfg.num_patches)
            aux0, aux1 = a.item(), d.item()
        elif loss.kind == _lib.LOSS_CE:
            total = out.loss * loss.ce_scale
        else:
            total = -out.loss
        if not forward_only:
            total.backward()
            dpatch.copy_(p.grad)
        sup = labels[:, 1:] != -100
        P = self.cfg.num_patches
        self._sup_logits = logits[:, P:-1][sup].detach()          # [R, V], what VLAEngine.tap("logits") returns
        act = logits[:, P:-1][sup][:, 31744:32000].argmax(-1).int() + 31744
        tgt = labels[:, 1:][sup]
        pred_ids.copy_(torch.where(tgt > 2, act, torch.full_like(act, -1)))
        with torch.no_grad():
            scalars[_lib.S_LOSS] = total.item()
            scalars[_lib.S_CE] = out.loss.item()
            scalars[_lib.S_AUX0] = aux0
            scalars[_lib.S_AUX1] = aux1
            scalars[_lib.S_UAD] = uad

    # ---- the whole attack iteration, same contract as VLAEngine.attack_step (vla_attack_step) ----
    def set_step_state(self, placement_index, adam_step):
        self._place, self._adam = int(placement_index), int(adam_step)

    def get_step_state(self):
        return self._place, self._adam

    def make_comm(self, rank, world):
        return GlooComm(world) if world > 1 else None

    def attack_step(self, patch, m, v, dpatch, scalars_hist, pred_ids, fe_mode, loss, lr, opt_kind=_lib.OPT_ADAMW, clip_l1=0.0,
                    accumulate=None, comm=None, do_update=True, graph=True, betas=(0.9, 0.999), eps=1e-6):
        sc = scalars_hist[self._place]
        sc.zero_()
        self.fwd_bwd(patch, self._place, fe_mode, loss, dpatch, sc, pred_ids[:self.num_supervised])
        g = dpatch
        if accumulate is not None:
            accumulate.add_(dpatch)
            g = accumulate
        world = 1
        if comm is not None:
            comm.all_reduce_(g)
            world = comm.world
        if do_update:
            self.patch_update(patch, g, m, v, self._adam + 1, lr, kind=opt_kind, grad_scale=1.0 / world, clip_l1=clip_l1, scalars=sc,
                              betas=betas, eps=eps)
            self._adam += 1
            if accumulate is not None:
                accumulate.zero_()
        self._place += 1

    def tap(self, what, dtype=torch.float32, max_elems=None):
        if what == "px":                                           # normalised 6-channel front-end output of the last pass
            return self._px.reshape(-1).to(dtype)
        assert what == "logits", what
        return self._sup_logits.reshape(-1).to(dtype)

    def patch_update(self, patch, grad, m, v, step, lr, kind=_lib.OPT_ADAMW, grad_scale=1.0, clip_l1=0.0, scalars=None,
                     betas=(0.9, 0.999), eps=1e-6):
        with torch.no_grad():
            g = grad * grad_scale
            if scalars is not None:
                scalars[_lib.S_GRAD_MEAN] = g.mean().item()
            if clip_l1 > 0:
                g = g * min(1.0, clip_l1 / (g.abs().sum().item() + 1e-6))
            if kind == _lib.OPT_ADAMW:
                m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
                v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
                step_size = lr * math.sqrt(1 - betas[1] ** step) / (1 - betas[0] ** step)
                patch.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
            else:
                patch.sub_(lr * g.sign())
            patch.clamp_(0, 1)
