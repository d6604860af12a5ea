"""Python host side of the attack-iteration engine: owns the device arenas (torch tensors), loads weights and
drives ``vla_fwd_bwd`` / ``vla_patch_update`` through the C ABI.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_void_p
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .config import NORM_MEAN, NORM_STD, OpenVLAConfig


@dataclass
class LossSpec:
    """Which loss head the engine applies (reference call sites in include/vla_b200.h)."""
    kind: int = _lib.LOSS_UADA
    mse_weight: float = 5.0
    alpha: float = 0.8
    belta: float = 0.2
    ce_scale: float = 1.0

    def to_c(self) -> _lib.LossParams:
        return _lib.LossParams(self.kind, self.mse_weight, self.alpha, self.belta, self.ce_scale)


def rope_tables(L: int, head_dim: int, theta: float):
    """cos / sin exactly as HF ``LlamaRotaryEmbedding`` produces them for bf16 activations (fp32 maths, bf16 rounding)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.outer(torch.arange(L, dtype=torch.float32), inv_freq)
    cos = freqs.cos().to(torch.bfloat16).float().contiguous()
    sin = freqs.sin().to(torch.bfloat16).float().contiguous()
    return cos, sin


def _c_config(cfg: OpenVLAConfig) -> _lib.Config:
    c = _lib.Config()
    c.img, c.patch = cfg.dino.img, cfg.dino.patch
    d, s, l = cfg.dino, cfg.siglip, cfg.llm
    assert d.img == s.img and d.patch == s.patch
    c.dino_dim, c.dino_depth, c.dino_heads, c.dino_mlp = d.dim, d.depth, d.heads, d.mlp_hidden
    c.dino_prefix, c.dino_layerscale = d.num_prefix, int(d.layerscale)
    c.sig_dim, c.sig_depth, c.sig_heads, c.sig_mlp = s.dim, s.depth, s.heads, s.mlp_hidden
    c.sig_prefix, c.sig_layerscale = s.num_prefix, int(s.layerscale)
    c.vit_ln_eps = d.ln_eps
    c.llm_hidden, c.llm_layers, c.llm_heads, c.llm_ffn, c.vocab = l.hidden, l.layers, l.heads, l.ffn, l.vocab
    c.rms_eps = l.rms_eps
    for i in range(2):
        for j in range(3):
            c.norm_mean[i][j] = NORM_MEAN[i][j]
            c.norm_std[i][j] = NORM_STD[i][j]
    return c


class NcclComm:
    """A ``vla_comm`` handle.  It belongs to the engine that made it (``VLAEngine.make_comm`` hands out ONE communicator per
    engine and destroys it with the engine, after the CUDA graphs that recorded its all-reduce)."""

    def __init__(self, lib, handle, world, engine=None):
        self._lib, self.handle, self.world, self._engine = lib, handle, world, engine

    def all_reduce_(self, t):
        """In-place sum over ranks of a float32 device tensor on the current stream (``vla_allreduce_patch_grad``)."""
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
        _lib.check(self._lib.vla_allreduce_patch_grad(self.handle, _lib.ptr(t), t.numel(), _lib.cur_stream()), "vla_allreduce_patch_grad")

    def close(self):
        """Destroy the communicator (all ranks must call it).  The engine's recorded graphs go first: NCCL blocks the destruction
        of a communicator while a graph still holds its collectives."""
        if self.handle:
            eng = self._engine() if self._engine is not None else None
            if eng is not None and getattr(eng, "_h", None):
                eng.drop_graphs()
                eng._comm = None
            self._lib.vla_comm_destroy(self.handle)
            self.handle = None


class VLAEngine:
    """One engine per process / GPU.  ``batch`` is the per-GPU batch, ``text_len`` the padded text length T."""

    def __init__(self, cfg: OpenVLAConfig, batch: int, text_len: int, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.VLAError("VLAEngine needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg, self.B, self.T = cfg, batch, text_len
        self.L = text_len + cfg.num_patches - 1
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self._lib = _lib.lib()
        self._h = c_void_p()
        ccfg = _c_config(cfg)
        _lib.check(self._lib.vla_engine_create(byref(ccfg), byref(self._h)), "vla_engine_create")
        wbytes = self._lib.vla_engine_weight_bytes(self._h)
        sbytes = self._lib.vla_engine_workspace_bytes(self._h, batch, text_len)
        self.weight_arena = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        self.workspace = torch.empty(sbytes, dtype=torch.uint8, device=self.device)
        self.num_supervised = 0
        self._weights_ok = False
        self._bind(batch, text_len)

    def _bind(self, batch, text_len):
        # the engine drops the last text position (dead under the causal mask + shifted CE): L = T + P - 1
        self.B, self.T, self.L = batch, text_len, text_len + self.cfg.num_patches - 1
        _lib.check(self._lib.vla_engine_set_buffers(self._h, _lib.ptr(self.weight_arena), self.weight_arena.numel(),
                                                    _lib.ptr(self.workspace), self.workspace.numel(), batch, text_len),
                   "vla_engine_set_buffers")
        cos, sin = rope_tables(self.L, self.cfg.llm.head_dim, self.cfg.llm.rope_theta)
        _lib.check(self._lib.vla_engine_set_rope(self._h, _lib.ptr(cos), _lib.ptr(sin), self.L, _lib.cur_stream()),
                   "vla_engine_set_rope")
        torch.cuda.synchronize(self.device)

    def ensure_plan(self, batch, text_len):
        """Re-plan the activation arena when the collator hands over a different (B, T); weights stay in place."""
        if (batch, text_len) == (self.B, self.T):
            return
        need = self._lib.vla_engine_workspace_bytes(self._h, batch, text_len)
        if need > self.workspace.numel():
            torch.cuda.synchronize(self.device)
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._bind(batch, text_len)

    def shrink_plan(self, batch, text_len):
        """Re-plan for (batch, text_len) AND give the activation arena back down to that plan's size (``ensure_plan`` only ever
        grows it): e.g. after a one-off large-batch run."""
        torch.cuda.synchronize(self.device)
        need = self._lib.vla_engine_workspace_bytes(self._h, batch, text_len)
        self.workspace = None
        torch.cuda.empty_cache()
        self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._bind(batch, text_len)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            # the communicator is NOT destroyed here: ncclCommDestroy is a collective, and garbage collection is not a point all
            # ranks reach together; call comm.close() explicitly (or leave it to process exit)
            self._lib.vla_engine_destroy(h)
            self._h = None

    # ---- weights -------------------------------------------------------------------------------------------
    def load_state_dict(self, sd, strict=True):
        """Copy HF-named tensors into the engine arena (any device / float dtype; converted to bf16 on the GPU).
        Entries that are not on the attack hot path are skipped unless ``strict``."""
        stream = _lib.cur_stream()
        for name, t in sd.items():
            g = t.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
            rc = self._lib.vla_engine_load_weight(self._h, name.encode(), _lib.ptr(g), g.numel(), stream)
            if rc != 0:
                msg = self._lib.vla_last_error().decode()
                if strict or "unknown weight" not in msg:
                    raise _lib.VLAError(f"load_weight({name}): {msg}")
            del g
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.vla_engine_weights_ready(self._h), "vla_engine_weights_ready")
        self._weights_ok = True

    def load_random_weights(self, seed=0, init="reference"):
        """Random-init weights generated on the GPU one tensor at a time (no checkpoint is available offline)."""
        from .weights import param_shapes
        g = torch.Generator(device=self.device).manual_seed(seed)
        from .weights import random_tensor
        stream = _lib.cur_stream()
        for name, shape in param_shapes(self.cfg).items():
            t = random_tensor(name, shape, g, self.device, init).to(torch.bfloat16).contiguous()
            _lib.check(self._lib.vla_engine_load_weight(self._h, name.encode(), _lib.ptr(t), t.numel(), stream),
                       f"load_weight({name})")
            del t
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib.vla_engine_weights_ready(self._h), "vla_engine_weights_ready")
        self._weights_ok = True

    # ---- per outer iteration -------------------------------------------------------------------------------
    def set_batch(self, obs_u8, input_ids, attention_mask, labels):
        """obs_u8 [B,H,W,3] uint8 (host, ideally pinned, or device); ids / mask / labels [B,T] host tensors."""
        B, T = input_ids.shape
        ids = input_ids.to(torch.int64).contiguous().cpu()
        mask = attention_mask.to(torch.uint8).contiguous().cpu()
        lab = labels.to(torch.int64).contiguous().cpu()
        obs = obs_u8.contiguous()
        assert obs.dtype == torch.uint8 and tuple(obs.shape) == (B, self.cfg.img, self.cfg.img, 3), obs.shape
        _lib.check(self._lib.vla_engine_set_batch(self._h, _lib.ptr(obs), int(obs.is_cuda), _lib.ptr(ids), _lib.ptr(mask),
                                                  _lib.ptr(lab), B, T, _lib.cur_stream()), "vla_engine_set_batch")
        self.num_supervised = self._lib.vla_engine_num_supervised(self._h)
        return self.num_supervised

    def set_placements(self, xy, theta):
        """xy int32 [steps,B,2], theta float32 [steps,B,2,3] (numpy or torch, host)."""
        xy = np.ascontiguousarray(np.asarray(xy), dtype=np.int32)
        theta = np.ascontiguousarray(np.asarray(theta), dtype=np.float32)
        steps = xy.shape[0]
        assert xy.shape == (steps, self.B, 2) and theta.shape == (steps, self.B, 2, 3), (xy.shape, theta.shape)
        _lib.check(self._lib.vla_engine_set_placements(self._h, xy.ctypes.data_as(c_void_p), theta.ctypes.data_as(c_void_p),
                                                       steps, _lib.cur_stream()), "vla_engine_set_placements")

    # ---- per inner iteration -------------------------------------------------------------------------------
    def fwd_bwd(self, patch, step_idx, fe_mode, loss: LossSpec, dpatch, scalars, pred_ids, forward_only=False):
        """patch / dpatch f32 [3,ph,pw] (device); scalars f32 [8] (device); pred_ids i32 [>= num_supervised]."""
        assert self._weights_ok, "weights not loaded"
        lp = loss.to_c()
        _lib.check(self._lib.vla_fwd_bwd(self._h, _lib.ptr(patch), patch.shape[1], patch.shape[2], step_idx, fe_mode,
                                         byref(lp), _lib.ptr(dpatch), _lib.ptr(scalars), _lib.ptr(pred_ids),
                                         _lib.FLAG_FORWARD_ONLY if forward_only else 0, _lib.cur_stream()), "vla_fwd_bwd")

    def patch_update(self, patch, grad, m, v, step, lr, kind=_lib.OPT_ADAMW, grad_scale=1.0, clip_l1=0.0, scalars=None,
                     betas=(0.9, 0.999), eps=1e-6):
        _lib.check(self._lib.vla_patch_update(_lib.ptr(patch), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), patch.numel(),
                                              step, lr, betas[0], betas[1], eps, kind, grad_scale, clip_l1,
                                              _lib.ptr(scalars), _lib.cur_stream()), "vla_patch_update")

    # ---- whole attack iteration (one C-ABI call, one CUDA graph launch after the first two calls) ----------
    def set_step_state(self, placement_index: int, adam_step: int):
        """Device-side counters of ``attack_step``: which uploaded placement the next step uses, and optimiser steps so far."""
        _lib.check(self._lib.vla_engine_set_step_state(self._h, int(placement_index), int(adam_step), _lib.cur_stream()),
                   "vla_engine_set_step_state")

    def get_step_state(self):
        a, b = ctypes.c_int(), ctypes.c_int()
        _lib.check(self._lib.vla_engine_get_step_state(self._h, byref(a), byref(b)), "vla_engine_get_step_state")
        return a.value, b.value

    def make_comm(self, rank: int, world: int):
        """NCCL communicator for the patch-gradient all-reduce (``vla_comm``); the 128-byte id travels over the already
        initialised ``torch.distributed`` group (host-side plumbing only).  ``None`` for a single process."""
        if world <= 1:
            return None
        if getattr(self, "_comm", None) is not None and self._comm.handle and self._comm.world == world:
            return self._comm
        import weakref
        import torch.distributed as dist
        idbuf = (ctypes.c_uint8 * _lib.COMM_ID_BYTES)()
        if rank == 0:
            _lib.check(self._lib.vla_comm_unique_id(idbuf), "vla_comm_unique_id")
        box = [bytes(idbuf)]
        dist.broadcast_object_list(box, src=0)
        idbuf = (ctypes.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(box[0])
        h = c_void_p()
        torch.cuda.set_device(self.device)
        _lib.check(self._lib.vla_comm_create(idbuf, rank, world, byref(h)), "vla_comm_create")
        self._comm = NcclComm(self._lib, h, world, weakref.ref(self))
        return self._comm

    def drop_graphs(self):
        _lib.check(self._lib.vla_engine_drop_graphs(self._h), "vla_engine_drop_graphs")

    def attack_step(self, patch, m, v, dpatch, scalars_hist, pred_ids, fe_mode, loss: LossSpec, lr, opt_kind=_lib.OPT_ADAMW,
                    clip_l1=0.0, accumulate=None, comm=None, do_update=True, graph=True, betas=(0.9, 0.999), eps=1e-6):
        """front end -> model -> loss -> backward -> [accumulate] -> [all-reduce] -> update -> clamp, for the placement the
        device-side counter points at; the step's scalar record lands in ``scalars_hist[counter]``."""
        assert self._weights_ok, "weights not loaded"
        sp = _lib.StepParams(patch.shape[1], patch.shape[2], fe_mode, loss.to_c(), opt_kind, float(lr), betas[0], betas[1], eps,
                             float(clip_l1), (0 if graph else _lib.STEP_NO_GRAPH) | (0 if do_update else _lib.STEP_NO_UPDATE))
        _lib.check(self._lib.vla_attack_step(self._h, _lib.ptr(patch), _lib.ptr(m), _lib.ptr(v), _lib.ptr(dpatch),
                                             _lib.ptr(accumulate), byref(sp), comm.handle if comm is not None else None,
                                             _lib.ptr(scalars_hist), _lib.ptr(pred_ids), _lib.cur_stream()), "vla_attack_step")

    def full_vocab_pred(self):
        """int32 [num_supervised] (device): full-vocabulary argmax of every supervised logits row of the last pass."""
        out = torch.empty(max(self.num_supervised, 1), dtype=torch.int32, device=self.device)
        _lib.check(self._lib.vla_engine_full_vocab_pred(self._h, _lib.ptr(out), _lib.cur_stream()), "vla_engine_full_vocab_pred")
        return out[:self.num_supervised]

    def decode_greedy(self, prompt_len: int, n_tokens: int, tokens):
        """KV-cache greedy decode after a forward-only prefill (``vla_engine_decode_greedy``); tokens int32 [B, n_tokens] (device)."""
        assert tokens.dtype == torch.int32 and tokens.is_cuda and tuple(tokens.shape) == (self.B, n_tokens) and tokens.is_contiguous()
        _lib.check(self._lib.vla_engine_decode_greedy(self._h, int(prompt_len), int(n_tokens), _lib.ptr(tokens), _lib.cur_stream()),
                   "vla_engine_decode_greedy")

    @property
    def graph_kernel_nodes(self) -> int:
        return self._lib.vla_graph_kernel_nodes(self._h)

    def set_single_stream(self, on: bool):
        """True: both vision towers on the caller's stream (clean per-kernel timing); False (default): two streams."""
        _lib.check(self._lib.vla_engine_set_single_stream(self._h, int(on)), "vla_engine_set_single_stream")

    def tap(self, what: str, dtype=torch.bfloat16, max_elems=None):
        if max_elems is None:   # largest tappable activation: the residual stream / the 6-channel image / the supervised logits
            max_elems = max(self.B * self.L * self.cfg.llm.hidden, self.B * 6 * self.cfg.img ** 2,
                            max(self.num_supervised, 1) * self.cfg.llm.vocab,
                            self.B * max(self.cfg.dino.tokens * self.cfg.dino.dim, self.cfg.siglip.tokens * self.cfg.siglip.dim))
        buf = torch.empty(max_elems, dtype=dtype, device=self.device)
        n = self._lib.vla_engine_debug_tap(self._h, what.encode(), _lib.ptr(buf), buf.numel() * buf.element_size(),
                                           _lib.cur_stream())
        if n < 0:
            raise _lib.VLAError(self._lib.vla_last_error().decode())
        return buf[:n].clone()
