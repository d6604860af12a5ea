"""Front end: patch paste + random affine warp + composite + dual normalise.

Follows ``VLAAttacker/white_patch/appply_random_transform.py`` (reference @ a0bef502):
``normalize`` :16-19, ``rotation_matrix``/``shear_matrix`` :26-41, ``combined_transform_matrix`` :80-91,
``apply_affine_transform`` :93-102, ``apply_random_patch_batch`` :104-136, ``random_paste_patch`` :138-158,
``paste_patch_fix`` :160-188, ``im_process`` :190-197.  The RNG draw order (python ``random`` for x,y; numpy for
the affine) is part of the contract: placements are drawn by ``draw_placements`` in exactly that order and the
deterministic part is ``apply_patch_batch``.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn.functional as F

MODE_WARP = 0      # apply_random_patch_batch(geometry=True): warp, keep where canvas >= -20
MODE_PASTE20 = 1   # apply_random_patch_batch(geometry=False): no warp, same `< -20` test
MODE_FIX = 2       # paste_patch_fix / random_paste_patch: no warp, `canvas != -100` test
MODE_NONE = 3      # im_process: no patch

ANGLE, SHX, SHY = 30, 0.2, 0.2   # :11-13


def rotation_matrix(theta_deg):   # :26-34
    theta = np.deg2rad(theta_deg)
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)


def shear_matrix(shx, shy):       # :36-41
    return np.array([[1, shx, 0], [shy, 1, 0], [0, 0, 1]], dtype=np.float32)


def combined_transform_matrix():  # :80-91
    if np.random.rand() < 0.2:
        return np.eye(3, dtype=np.float32)
    angle = np.random.uniform(-ANGLE, ANGLE)
    shx = np.random.uniform(-SHX, SHX)
    shy = np.random.uniform(-SHY, SHY)
    return np.dot(shear_matrix(shx, shy), rotation_matrix(angle))


def draw_placements(batch, img_hw, patch_hw, geometry):
    """Host RNG protocol of one ``apply_random_patch_batch`` call (:123-124,:127-129): per image
    ``x = random.randint(0, W-pw)``, ``y = random.randint(0, H-ph)``, then (geometry only) the affine draw.
    Returns xy int32 [B,2] (x, y) and theta float32 [B,2,3] (identity rows when geometry is False)."""
    H, W = img_hw
    ph, pw = patch_hw
    xy = np.zeros((batch, 2), dtype=np.int32)
    theta = np.zeros((batch, 2, 3), dtype=np.float32)
    for b in range(batch):
        x = random.randint(0, W - pw)
        y = random.randint(0, H - ph)
        xy[b] = (x, y)
        m = combined_transform_matrix() if geometry else np.eye(3, dtype=np.float32)
        theta[b] = m[:2, :]
    return xy, theta


def to_float_image(obs_u8: torch.Tensor) -> torch.Tensor:
    """torchvision ``ToTensor`` on a uint8 HWC image: CHW float32 / 255 (:108)."""
    return obs_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)


def apply_patch_batch(obs_u8, patch, xy, theta, mode, mean, std):
    """obs_u8 [B,H,W,3] uint8, patch [3,ph,pw] f32 (may require grad) -> [B,6,H,W] f32 (:104-136)."""
    B, H, W, _ = obs_u8.shape
    ims = to_float_image(obs_u8)
    m0 = torch.as_tensor(mean[0], dtype=torch.float32)
    s0 = torch.as_tensor(std[0], dtype=torch.float32)
    m1 = torch.as_tensor(mean[1], dtype=torch.float32)
    s1 = torch.as_tensor(std[1], dtype=torch.float32)
    out = []
    ph, pw = patch.shape[1:]
    for b in range(B):
        im = ims[b]
        if mode != MODE_NONE:
            canvas = torch.ones(3, H, W) * -100
            x, y = int(xy[b][0]), int(xy[b][1])
            canvas[:, y:y + ph, x:x + pw] = patch
            if mode == MODE_WARP:
                aff = torch.as_tensor(theta[b], dtype=torch.float32).unsqueeze(0)
                grid = F.affine_grid(aff, (1, 3, H, W), align_corners=False)
                canvas = F.grid_sample(canvas.unsqueeze(0), grid, align_corners=False, padding_mode="border")[0]
            if mode == MODE_FIX:
                im = torch.where(canvas != -100, canvas, im)
            else:
                im = torch.where(canvas < -20, im, canvas)
        im0 = (im[None] - m0[None, :, None, None]) / s0[None, :, None, None]
        im1 = (im[None] - m1[None, :, None, None]) / s1[None, :, None, None]
        out.append(torch.cat([im0, im1], dim=1))
    return torch.cat(out, dim=0)


def simulation_paste(image_u8, patch, geometry, angle=1, shx=0.1, shy=0.1, position=(0, 0)):
    """Restatement of ``simulation_random_patch`` (appply_random_transform.py:43-78): image uint8 ndarray [H,W,3], patch f32
    [3,ph,pw] in [0,1] -> uint8 ndarray [H,W,3].  ToPILImage on a float tensor is ``mul(255).byte()`` (truncation)."""
    image = torch.from_numpy(np.asarray(image_u8)).permute(2, 0, 1)
    C, H, W = image.shape
    canvas = torch.ones(C, H, W) * -100
    ph, pw = patch.shape[1:]
    q = patch.detach().mul(255).byte()
    x, y = position
    canvas[:, y:y + ph, x:x + pw] = q
    if geometry:
        m = torch.tensor(np.dot(shear_matrix(shx, shy), rotation_matrix(angle)))
        grid = F.affine_grid(m[:2, :].unsqueeze(0), (1, C, H, W), align_corners=False)
        canvas = F.grid_sample(canvas.unsqueeze(0), grid, align_corners=False, padding_mode="border")
    out = torch.where(canvas < 0, image, canvas)
    return out.squeeze(0).permute(1, 2, 0).numpy().astype(np.uint8)
