"""Greedy action decode on the engine: the consumer side of ``patch.pt`` in the closed-loop evaluation.

Mirrors ``OpenVLAForActionPrediction.predict_action`` (prismatic/extern/hf/modeling_prismatic.py:506-536): append the
empty token 29871 when missing, generate ``action_dim`` tokens greedily, map token ids to bin centres
(``vocab_size - id``, clip, centres) and un-normalise with the dataset statistics.  Like HF ``generate`` the decode
re-uses a KV cache: one prefill pass of the engine over image + prompt, then one single-position step per further token
(``vla_engine_decode_greedy``: for batch <= 4 five fused HBM-bound kernels per decoder layer, one recorded step replayed per token;
the tcgen05 GEMM with M = batch otherwise).  The image is expected to carry the patch already
(``RandomPatchTransform.simulation_random_patch``), so the front end runs in its no-patch mode (``im_process``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib, labels as lab
from .config import IGNORE_INDEX, PAD_TOKEN_ID
from .engine import LossSpec, VLAEngine

EMPTY_TOKEN = 29871       # the '' token that follows "Out:" in the training prompts (modeling_prismatic.py:512-516)
DUMMY_LABEL = 31872       # any valid label: it only marks the decode row as a row whose logits are wanted


class ActionPolicy:
    def __init__(self, engine: VLAEngine, norm_stats: Optional[dict] = None):
        self.engine = engine
        self.norm_stats = norm_stats or {}

    # -- token level ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate_action_tokens(self, images_u8: torch.Tensor, input_ids: torch.Tensor, n_tokens: int = 7, kv_cache: bool = True) -> np.ndarray:
        """images_u8 uint8 [B,H,W,3], input_ids int64 [B,T0] (un-padded prompts of equal length) -> int64 [B, n_tokens]:
        argmax over the full vocabulary at every step, as ``generate(do_sample=False)`` does.  One prefill pass over the prompt
        and the image, then ``n_tokens - 1`` single-position steps on the KV cache the prefill left in the engine's arena
        (``vla_engine_decode_greedy``); ``kv_cache=False`` re-runs the full forward for every token (round-1 path, kept as
        the cross-check of the cache)."""
        eng = self.engine
        ids0 = input_ids.to(torch.int64).cpu()
        B, T0 = ids0.shape
        T = T0 + n_tokens + 1                       # fixed plan: prompt + generated tokens + the marker position
        eng.ensure_plan(B, T)
        dev = eng.device
        zeros_patch = torch.zeros(3, 1, 1, device=dev)
        dpatch = torch.zeros_like(zeros_patch)
        scal = torch.zeros(_lib.NUM_SCALARS, device=dev)
        pred = torch.zeros(B, dtype=torch.int32, device=dev)
        xy = np.zeros((1, B, 2), dtype=np.int32)
        theta = np.zeros((1, B, 2, 3), dtype=np.float32)
        gen = torch.zeros(B, 0, dtype=torch.int64)
        obs = images_u8.contiguous()
        ce = LossSpec(_lib.LOSS_CE, ce_scale=1.0)

        def forward(known):
            """Full forward over the ``known`` tokens; the logits row of the last known position is the supervised row."""
            cur = known.shape[1]
            ids = torch.full((B, T), PAD_TOKEN_ID, dtype=torch.int64)
            ids[:, :cur] = known
            ids[:, cur] = DUMMY_LABEL
            mask = torch.zeros(B, T, dtype=torch.bool)
            mask[:, :cur + 1] = True
            labels = torch.full((B, T), IGNORE_INDEX, dtype=torch.int64)
            labels[:, cur] = DUMMY_LABEL             # makes row cur - 1 the (only) supervised row of each sample
            R = eng.set_batch(obs, ids, mask, labels)
            assert R == B, (R, B)
            eng.set_placements(xy, theta)
            eng.fwd_bwd(zeros_patch, 0, _lib.FE_NONE, ce, dpatch, scal, pred, forward_only=True)

        if kv_cache:
            forward(ids0)                            # prefill: image + prompt, KV cache left in the arena
            toks = torch.zeros(B, n_tokens, dtype=torch.int32, device=dev)
            eng.decode_greedy(T0, n_tokens, toks)
            return toks.cpu().to(torch.int64).numpy()
        for k in range(n_tokens):
            forward(torch.cat([ids0, gen], dim=1))
            logits = eng.tap("logits", dtype=torch.float32, max_elems=B * self.engine.cfg.llm.vocab).view(B, -1)
            nxt = logits.argmax(dim=1).cpu().to(torch.int64)
            gen = torch.cat([gen, nxt[:, None]], dim=1)
        return gen.numpy()

    # -- reference API -------------------------------------------------------------------------------------------
    def predict_action(self, image_u8, input_ids: torch.Tensor, unnorm_key: Optional[str] = None) -> np.ndarray:
        """``image_u8`` uint8 [H,W,3] (ndarray or tensor), ``input_ids`` [1,T0] -> un-normalised action [action_dim]."""
        ids = input_ids.to(torch.int64).cpu()
        if not torch.all(ids[:, -1] == EMPTY_TOKEN):
            ids = torch.cat([ids, torch.full((ids.shape[0], 1), EMPTY_TOKEN, dtype=torch.int64)], dim=1)
        stats = self.get_action_stats(unnorm_key)
        n = len(stats["q01"]) if stats else 7
        img = torch.as_tensor(np.ascontiguousarray(image_u8), dtype=torch.uint8)[None]
        toks = self.generate_action_tokens(img, ids, n)[0]
        normalized = lab.decode_token_ids_to_actions(toks)
        if not stats:
            return normalized
        mask = np.asarray(stats.get("mask", np.ones_like(stats["q01"], dtype=bool)))
        high, low = np.array(stats["q99"]), np.array(stats["q01"])
        return np.where(mask, 0.5 * (normalized + 1) * (high - low) + low, normalized)

    def get_action_stats(self, unnorm_key: Optional[str] = None) -> dict:
        if not self.norm_stats:
            return {}
        if unnorm_key is None:
            assert len(self.norm_stats) == 1, f"pass unnorm_key, one of {list(self.norm_stats)}"
            unnorm_key = next(iter(self.norm_stats))
        assert unnorm_key in self.norm_stats, f"unnorm_key must be one of {list(self.norm_stats)}"
        return self.norm_stats[unnorm_key]["action"]
