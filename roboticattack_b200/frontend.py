"""``RandomPatchTransform`` with the reference's method names, backed by the fused CUDA front end.

Reference: VLAAttacker/white_patch/appply_random_transform.py:8-197.  Differences: images may be a list of PIL images
(as the collator yields), a uint8 array/tensor [B,H,W,3]; the result is the bf16 tensor the model consumes
(the reference returns fp32 and callers immediately cast with ``.to(torch.bfloat16)``, UADA.py:142); gradients
flow to ``patch`` through a ``torch.autograd.Function`` that calls ``vla_patch_frontend_bwd``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .config import NORM_MEAN, NORM_STD
from .synthetic import draw_placements


def _as_uint8_batch(images, device):
    if isinstance(images, torch.Tensor):
        t = images
    else:
        t = torch.from_numpy(np.stack([np.asarray(im, dtype=np.uint8) for im in images]))
    assert t.dtype == torch.uint8 and t.dim() == 4 and t.shape[-1] == 3, "images must be uint8 [B,H,W,3]"
    return t.to(device).contiguous()


def fixed_affine(angle, shx, shy):
    """theta = (S . R)[:2] in float32 for a fixed rotation (degrees) and shear (appply_random_transform.py:26-41,69-75)."""
    t = np.radians(angle)
    c, s = np.cos(t), np.sin(t)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    S = np.array([[1, shx, 0], [shy, 1, 0], [0, 0, 1]], dtype=np.float32)
    return np.ascontiguousarray(np.dot(S, R)[:2, :], dtype=np.float32)


class _FrontendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patch, obs, xy, theta, mode, norm):
        B, H, W, _ = obs.shape
        out = torch.empty(B, 6, H, W, device=obs.device, dtype=torch.bfloat16)
        p = patch.detach().to(torch.float32).contiguous()
        _lib.check(_lib.lib().vla_patch_frontend_fwd(_lib.ptr(obs), _lib.ptr(p), _lib.ptr(xy), _lib.ptr(theta), _lib.ptr(out),
                                                     B, H, W, p.shape[1], p.shape[2], mode, norm, _lib.cur_stream()),
                   "vla_patch_frontend_fwd")
        ctx.save_for_backward(p, xy, theta)
        ctx.meta = (B, H, W, mode, norm)
        return out

    @staticmethod
    def backward(ctx, dout):
        p, xy, theta = ctx.saved_tensors
        B, H, W, mode, norm = ctx.meta
        dp = torch.empty_like(p)
        g = dout.to(torch.bfloat16).contiguous()
        _lib.check(_lib.lib().vla_patch_frontend_bwd(_lib.ptr(g), _lib.ptr(p), _lib.ptr(xy), _lib.ptr(theta), _lib.ptr(dp),
                                                     B, H, W, p.shape[1], p.shape[2], mode, norm, _lib.cur_stream()),
                   "vla_patch_frontend_bwd")
        return dp, None, None, None, None, None


class RandomPatchTransform:
    def __init__(self, device, resize_patch=False):
        self.device = torch.device(device)
        self.angle, self.shx, self.shy = 30, 0.2, 0.2
        if resize_patch:
            raise NotImplementedError("resize_patch=True is dead code in the reference (uses a variable before assignment)")
        self.resize_patch = resize_patch

    def normalize(self, images, mean, std):
        return (images - mean[None, :, None, None]) / std[None, :, None, None]

    def denormalize(self, images, mean, std):
        return images * std[None, :, None, None] + mean[None, :, None, None]

    # helpers kept under the reference's names (appply_random_transform.py:26-41,80-102); the engine itself never calls them
    def rotation_matrix(self, theta):
        theta = np.radians(theta)
        c, s = np.cos(theta), np.sin(theta)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)

    def shear_matrix(self, shx, shy):
        return np.array([[1, shx, 0], [shy, 1, 0], [0, 0, 1]], dtype=np.float32)

    def combined_transform_matrix(self):
        """Same RNG protocol as the reference: u < 0.2 -> identity, else angle, shx, shy in that order."""
        if np.random.rand() < 0.2:
            return torch.eye(3, dtype=torch.float32)
        angle = np.random.uniform(-self.angle, self.angle)
        shx = np.random.uniform(-self.shx, self.shx)
        shy = np.random.uniform(-self.shy, self.shy)
        return torch.tensor(np.dot(self.shear_matrix(shx, shy), self.rotation_matrix(angle)))

    def apply_affine_transform(self, image, transform_matrix):
        """Generic [C,H,W] warp (bilinear, border padding, align_corners=False) with torch ops; the attack path uses the fused
        CUDA front end instead."""
        import torch.nn.functional as F
        if image.ndim == 4:
            image = image.squeeze(0)
        aff = torch.as_tensor(transform_matrix, dtype=image.dtype, device=image.device)[:2, :].unsqueeze(0)
        grid = F.affine_grid(aff, image.unsqueeze(0).size(), align_corners=False)
        return F.grid_sample(image.unsqueeze(0), grid, align_corners=False, padding_mode="border")

    def _run(self, images, patch, mean, std, mode, geometry):
        obs = _as_uint8_batch(images, self.device)
        B, H, W, _ = obs.shape
        mean = NORM_MEAN if mean is None else [[float(v) for v in m] for m in mean]
        std = NORM_STD if std is None else [[float(v) for v in s] for s in std]
        norm = _lib.norm_array(mean, std)
        if mode == _lib.FE_NONE:
            xy = torch.zeros(B, 2, dtype=torch.int32, device=self.device)
            theta = torch.zeros(B, 2, 3, device=self.device)
            patch = torch.zeros(3, 1, 1, device=self.device)
        else:
            xy_np, th_np = draw_placements(B, (H, W), tuple(patch.shape[1:]), geometry, steps=1)
            xy = torch.from_numpy(xy_np[0]).to(self.device)
            theta = torch.from_numpy(th_np[0]).to(self.device)
        return _FrontendFn.apply(patch.to(self.device), obs, xy, theta, mode, norm)

    def apply_random_patch_batch(self, images, patch, mean=None, std=None, geometry=False):
        return self._run(images, patch, mean, std, _lib.FE_WARP if geometry else _lib.FE_PASTE20, geometry)

    def random_paste_patch(self, images, patch, mean=None, std=None):
        return self._run(images, patch, mean, std, _lib.FE_FIX, False)

    def paste_patch_fix(self, images, patch, mean=None, std=None, inference=False):
        return self._run(images, patch, mean, std, _lib.FE_FIX, False)

    def im_process(self, images, mean=None, std=None):
        return self._run(images, None, mean, std, _lib.FE_NONE, False)

    def simulation_random_patch(self, image, patch, geometry=False, colorjitter=False, angle=1, shx=0.1, shy=0.1, position=(0, 0)):
        """Eval-time paste (appply_random_transform.py:43-78): ``image`` uint8 ndarray [H,W,3] -> uint8 ndarray [H,W,3] with the
        patch quantised to uint8 and pasted at the FIXED ``position`` = (x, y), warped by the FIXED (angle, shx, shy) when
        ``geometry``.  ``colorjitter`` is accepted and ignored, as in the reference.  Runs ``vla_patch_sim_paste`` on the GPU."""
        img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8))[None].to(self.device)
        _, H, W, _ = img.shape
        p = patch.detach().to(self.device, torch.float32).contiguous()
        xy = torch.tensor([[int(position[0]), int(position[1])]], dtype=torch.int32, device=self.device)
        theta = torch.from_numpy(fixed_affine(angle, shx, shy)[None]).to(self.device)
        out = torch.empty_like(img)
        _lib.check(_lib.lib().vla_patch_sim_paste(_lib.ptr(img), _lib.ptr(p), _lib.ptr(xy), _lib.ptr(theta), _lib.ptr(out), 1, H, W,
                                                  p.shape[1], p.shape[2], int(bool(geometry)), _lib.cur_stream()),
                   "vla_patch_sim_paste")
        return out[0].cpu().numpy()
