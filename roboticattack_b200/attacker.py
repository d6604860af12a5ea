"""``OpenVLAAttacker`` for UADA / UPA / TMA / UADA-DDP with the reference's constructor and
``patchattack_unconstrained`` / ``attack`` signatures, driving the CUDA engine instead of PyTorch autograd.

Reference: VLAAttacker/white_patch/UADA.py:33-418, UPA.py:30-390, TMA.py:28-482, UADA_ddp.py:36-344.
What stays Python (as in the reference): the outer loop over batches, label masking, the LR schedule, logging,
validation cadence and checkpoint files.  What moved into the engine: everything inside the inner loop.
Differences from the reference that are deliberate and documented in DESIGN.md: the model weights are frozen in
every variant (UADA.py never freezes them and so also computes 7.5 B unused weight gradients); per-step scalars are
read back once per outer iteration instead of four ``.item()`` syncs per inner step; clean observations are uploaded
once per outer iteration instead of once per inner step.
"""
from __future__ import annotations

import math
import os
import pickle
import queue
import random
import threading
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib, labels as lab
from .config import NORM_MEAN, NORM_STD, OpenVLAConfig, openvla_7b
from .engine import LossSpec, VLAEngine
from .frontend import RandomPatchTransform
from .synthetic import draw_placements

IGNORE_INDEX = -100


def cosine_with_warmup(step: int, warmup: int, total: int, num_cycles: float = 0.5) -> float:
    """``transformers.get_cosine_schedule_with_warmup`` multiplier after ``step`` scheduler steps."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def observations_to_uint8(pixel_values, img: int) -> torch.Tensor:
    """The collator's ``pixel_values`` (list of PIL images, data_utils.py:203-204) -> pinned uint8 [B,H,W,3]."""
    if isinstance(pixel_values, torch.Tensor):
        t = pixel_values
    else:
        t = torch.from_numpy(np.stack([np.asarray(im.convert("RGB") if hasattr(im, "convert") else im, dtype=np.uint8)
                                       for im in pixel_values]))
    assert t.dtype == torch.uint8 and t.shape[1:] == (img, img, 3), f"expected uint8 [B,{img},{img},3], got {t.shape}"
    if not t.is_cuda and torch.cuda.is_available() and not t.is_pinned():
        t = t.pin_memory()
    return t


class LookaheadLoader:
    """Batch k + 1 is pulled from the loader and converted (PIL -> pinned uint8) while the device runs the ``innerLoop``
    steps of batch k -- on the CALLING thread, between the asynchronous launch of the inner steps and the
    read-back of their scalars.  The reference fetches and re-uploads synchronously (UADA.py:119-126,
    appply_random_transform.py:108).  Everything that consumes an RNG stream (the loader's ``iter`` / ``next``, label
    helpers, placement draws) therefore runs on one thread in the reference's order: next(k), placements(k), next(k + 1),
    placements(k + 1) ... so a seeded run is reproducible with and without the lookahead (``VLA_PREFETCH=0`` disables it).
    ``next()`` restarts an exhausted loader like the reference's ``while``-loop (``restart=False``: StopIteration)."""

    def __init__(self, loader, img: int, restart: bool = True, lookahead: bool = True):
        self._loader, self._img, self._restart, self._lookahead = loader, img, restart, lookahead
        self._it = None
        self._slot = None          # (batch, exception) fetched ahead of time

    def _fetch(self):
        if self._it is None:
            self._it = iter(self._loader)
        try:
            batch = next(self._it)
        except StopIteration:
            if not self._restart:
                raise
            self._it = iter(self._loader)
            batch = next(self._it)
        batch = dict(batch)
        batch["pixel_values"] = observations_to_uint8(batch["pixel_values"], self._img)
        return batch

    def prefetch(self):
        """Fetch the next batch now (called while the device is busy); an exception is kept for the ``next()`` that needs it."""
        if not self._lookahead or self._slot is not None:
            return
        try:
            self._slot = (self._fetch(), None)
        except BaseException as e:   # noqa: BLE001 -- re-raised by next()
            self._slot = (None, e)

    def next(self):
        if self._slot is not None:
            (batch, err), self._slot = self._slot, None
            if err is not None:
                raise err
            return batch
        return self._fetch()

    def close(self):
        self._slot = None


class InnerLoopJob:
    """Results of one outer iteration on their way to the host.  On a CUDA device the device tensors are copied to pinned
    memory on a read-back stream that waits only for the iteration's own last kernel (an event on the compute stream), so
    ``result()`` returns when THIS iteration is done even if the next one is already queued behind it."""

    def __init__(self, host, scalars_dev, pred_dev):
        self.ctx = None                               # the caller's bookkeeping (iteration index, batch, learning rate ...)
        if scalars_dev.is_cuda:
            if host._rb_stream is None:
                host._rb_stream = torch.cuda.Stream(device=host.device)
            rb = host._rb_stream
            ready = torch.cuda.Event()
            ready.record()
            self._sc = torch.empty(scalars_dev.shape, dtype=scalars_dev.dtype, pin_memory=True)
            self._pr = torch.empty(pred_dev.shape, dtype=pred_dev.dtype, pin_memory=True)
            with torch.cuda.stream(rb):
                rb.wait_event(ready)
                self._sc.copy_(scalars_dev, non_blocking=True)
                self._pr.copy_(pred_dev, non_blocking=True)
                scalars_dev.record_stream(rb)
                pred_dev.record_stream(rb)
                self._done = torch.cuda.Event()
                self._done.record(rb)
        else:
            self._sc, self._pr, self._done = scalars_dev.clone(), pred_dev.clone(), None

    def result(self):
        if self._done is not None:
            self._done.synchronize()
            self._done = None
        return self._sc, self._pr


class AttackEngineHost:
    """Shared machinery: engine creation from ``vla`` (an HF-style module, a state dict, or a loaded VLAEngine),
    patch / optimiser state, one outer iteration = set_batch + placements + ``innerLoop`` engine steps."""

    def __init__(self, vla, cfg: Optional[OpenVLAConfig] = None, device="cuda:0", max_text_len: int = 64,
                 engine_factory=None):
        if cfg is None and getattr(vla, "cfg", None) is None and hasattr(getattr(vla, "config", None), "text_config"):
            from .config import config_from_hf
            cfg = config_from_hf(vla.config)          # the HF OpenVLAForActionPrediction module the wrappers pass in
        self.cfg = cfg or getattr(vla, "cfg", None) or openvla_7b()
        self.engine_factory = engine_factory or VLAEngine
        self.device = torch.device(device)
        self._vla = vla
        self.engine: Optional[VLAEngine] = vla if isinstance(vla, VLAEngine) else None
        self.max_text_len = max_text_len
        self.patch = self.m = self.v = self.grad = None
        self.opt_step = 0
        self.world_size, self.rank = 1, 0
        self.comm = None
        self._slots, self._slot = [{}, {}], 0
        self._rb_stream = None

    # -- engine ---------------------------------------------------------------------------------------------
    def ensure_engine(self, B: int, T: int):
        if self.engine is None:
            self.engine = self.engine_factory(self.cfg, B, T, device=self.device)
            src = self._vla
            if src is None:
                raise _lib.VLAError("no weights: pass an HF model, a state dict or a loaded VLAEngine")
            if isinstance(src, (str, os.PathLike)):   # the reference's `vla_path` (UADA_ddp.py:37-50)
                from .weights import resolve_vla
                sd, _ = resolve_vla(src)
            else:
                sd = src if isinstance(src, dict) else src.state_dict()
            self.engine.load_state_dict(sd, strict=False)
        self.engine.ensure_plan(B, T)
        return self.engine

    # -- patch / optimiser state -----------------------------------------------------------------------------
    def init_patch(self, patch_size, patch: Optional[torch.Tensor] = None):
        """``patch = torch.rand(patch_size)`` (UADA.py:104); DDP: rank 0's value is broadcast (UADA_ddp.py:140-144)."""
        if patch is None:
            patch = torch.rand(list(patch_size))
        self.patch = patch.to(self.device, torch.float32).contiguous()
        if self.world_size > 1:
            torch.distributed.broadcast(self.patch, src=0)
        self.m = torch.zeros_like(self.patch)
        self.v = torch.zeros_like(self.patch)
        self.grad = torch.zeros_like(self.patch)
        self.opt_step = 0

    def state_dict(self, outer_iter: int, sched_step: int, accumulate: Optional[torch.Tensor] = None) -> dict:
        """Everything a restart needs beyond ``patch.pt`` (the reference cannot resume): the optimiser moments
        and step counter, the gradient-accumulation buffer of TMA / UPA (``accumulate_steps`` > 1), the outer / scheduler
        position and the three host RNG streams that drive placements and init.  Only tensors, numbers, strings and lists, so
        that the file loads with ``torch.load(weights_only=True)`` (no pickle execution on a path taken from the environment).
        In the data-parallel attack rank 0 writes the file and every rank restores from it: the reference seeds all ranks
        identically (UADA_wrapper_ddp.py:53), i.e. the placement streams of the ranks are identical by design."""
        ver, mt, gauss = random.getstate()
        np_name, np_keys, np_pos, np_has_gauss, np_cached = np.random.get_state()
        st = {"patch": self.patch.detach().cpu(), "m": self.m.detach().cpu(), "v": self.v.detach().cpu(),
              "opt_step": int(self.opt_step), "outer_iter": int(outer_iter), "sched_step": int(sched_step),
              "py_random": {"version": int(ver), "state": torch.tensor(list(mt), dtype=torch.int64),
                            "gauss": None if gauss is None else float(gauss)},
              "np_random": {"name": str(np_name), "keys": torch.from_numpy(np_keys.astype(np.int64)), "pos": int(np_pos),
                            "has_gauss": int(np_has_gauss), "cached": float(np_cached)},
              "torch_rng": torch.get_rng_state()}
        if accumulate is not None:
            st["accumulate"] = accumulate.detach().cpu()
        return st

    def load_state_dict(self, st: dict, accumulate: Optional[torch.Tensor] = None):
        """Restores patch / moments / counters / accumulation buffer and the RNG streams; returns (next outer iteration,
        scheduler step)."""
        if self.patch is not None and tuple(st["patch"].shape) != tuple(self.patch.shape):
            raise ValueError(f"resume: patch shape {tuple(st['patch'].shape)} != configured {tuple(self.patch.shape)}")
        self.patch = st["patch"].to(self.device, torch.float32).contiguous()
        self.m = st["m"].to(self.device, torch.float32).contiguous()
        self.v = st["v"].to(self.device, torch.float32).contiguous()
        self.grad = torch.zeros_like(self.patch)
        self.opt_step = int(st["opt_step"])
        if accumulate is not None:
            if "accumulate" not in st:
                raise ValueError("resume: the run uses accumulate_steps > 1 but the state file has no accumulation buffer")
            accumulate.copy_(st["accumulate"].to(accumulate.device))
        pr, nr = st["py_random"], st["np_random"]
        random.setstate((int(pr["version"]), tuple(int(x) for x in pr["state"].tolist()), pr["gauss"]))
        np.random.set_state((nr["name"], nr["keys"].numpy().astype(np.uint32), int(nr["pos"]), int(nr["has_gauss"]), float(nr["cached"])))
        torch.set_rng_state(st["torch_rng"])
        return int(st["outer_iter"]) + 1, int(st["sched_step"])

    def ensure_comm(self):
        """Communicator of the patch-gradient all-reduce (created once, after ``world_size`` is known)."""
        if self.world_size > 1 and self.comm is None:
            self.comm = self.engine.make_comm(self.rank, self.world_size)
        return self.comm

    def submit_inner_loop(self, batch, n_inner, fe_mode, loss: LossSpec, lr, opt_kind, clip_l1=0.0, do_step=True,
                          accumulate=None, after_launch=None):
        """One outer iteration, asynchronously: upload the batch and the placements of all inner steps, then ``n_inner`` calls of
        ``vla_attack_step`` (one CUDA graph launch each), nothing read back in between.  ``after_launch`` runs on the host while
        the device works (the loaders' lookahead).  Returns an ``InnerLoopJob``: the per-step scalar records and the predicted ids
        of the last step travel to pinned host memory on a side stream as soon as the last step finishes, so that the caller
        can launch the NEXT outer iteration before it looks at this one's numbers (``job.result()``)."""
        obs = observations_to_uint8(batch["pixel_values"], self.cfg.img)
        B, T = batch["input_ids"].shape
        eng = self.ensure_engine(B, T)
        comm = self.ensure_comm()
        R = eng.set_batch(obs, batch["input_ids"], batch["attention_mask"], batch["labels"])
        geometry = fe_mode == _lib.FE_WARP
        xy, theta = draw_placements(B, (self.cfg.img, self.cfg.img), tuple(self.patch.shape[1:]), geometry, steps=n_inner)
        eng.set_placements(xy, theta)
        eng.set_step_state(0, self.opt_step)
        # two sets of result buffers, used alternately: the previous iteration's are still being read back
        self._slot ^= 1
        sl = self._slots[self._slot]
        if sl.get("scalars") is None or sl["scalars"].shape[0] < n_inner:
            sl["scalars"] = torch.zeros(n_inner, _lib.NUM_SCALARS, device=self.device)
        if sl.get("pred") is None or sl["pred"].numel() < R:
            sl["pred"] = torch.full((max(R, B * 8),), -1, dtype=torch.int32, device=self.device)
        scalars, pred = sl["scalars"], sl["pred"]
        for s in range(n_inner):
            eng.attack_step(self.patch, self.m, self.v, self.grad, scalars, pred, fe_mode, loss, lr, opt_kind=opt_kind,
                            clip_l1=clip_l1, accumulate=accumulate, comm=comm, do_update=do_step)
            if do_step:
                self.opt_step += 1
        full = self._full_vocab_pred(eng, R, pred[:R])
        job = InnerLoopJob(self, scalars[:n_inner], full)
        if after_launch is not None:
            after_launch()
        return job

    def run_inner_loop(self, batch, n_inner, fe_mode, loss: LossSpec, lr, opt_kind, clip_l1=0.0, do_step=True,
                       accumulate=None, after_launch=None):
        """``submit_inner_loop(...).result()``: (scalars [n_inner, 8] on the host, pred_ids [R] of the last inner step)."""
        return self.submit_inner_loop(batch, n_inner, fe_mode, loss, lr, opt_kind, clip_l1, do_step, accumulate, after_launch).result()

    def _full_vocab_pred(self, eng, R, pred):
        """The reference's metrics take ``action_preds = logits.argmax(dim=2)`` over the FULL vocabulary (UADA.py:168,229;
        TMA.py:150,274) and let ``decode_token_ids_to_actions`` clip whatever id comes out, whereas the loss-head kernel reports
        the argmax inside the 256 action classes (what ``weighted_loss`` / UAD use).  The two differ whenever a non-action token
        wins -- which an attack can provoke -- and TMA selects its best patch by an L1 built on these predictions, so the ids of
        the action rows are taken from the engine's fp32 logits rows here (one [R, V] argmax per validation batch / outer
        iteration; not on the per-inner-step path).  The engine computes that argmax on the device (``vla_engine_full_vocab_pred``)."""
        if hasattr(eng, "full_vocab_pred"):          # computed on the device next to the loss head (one [R] copy)
            return torch.where(pred >= 0, eng.full_vocab_pred().to(pred.device), pred)
        if not hasattr(eng, "tap"):
            return pred
        V = self.cfg.llm.vocab
        logits = eng.tap("logits", dtype=torch.float32, max_elems=R * V).view(R, V)
        full = logits.argmax(dim=1).to(torch.int32)
        return torch.where(pred >= 0, full.to(pred.device), pred)

    def evaluate(self, batch, fe_mode, loss: LossSpec):
        """Forward-only pass (validation): scalars [8] and pred_ids [R] on the host."""
        obs = observations_to_uint8(batch["pixel_values"], self.cfg.img)
        B, T = batch["input_ids"].shape
        eng = self.ensure_engine(B, T)
        R = eng.set_batch(obs, batch["input_ids"], batch["attention_mask"], batch["labels"])
        if fe_mode != _lib.FE_NONE:   # im_process() pastes nothing and draws nothing (appply_random_transform.py:190-197)
            xy, theta = draw_placements(B, (self.cfg.img, self.cfg.img), tuple(self.patch.shape[1:]), fe_mode == _lib.FE_WARP, 1)
            eng.set_placements(xy, theta)
        scalars = torch.zeros(_lib.NUM_SCALARS, device=self.device)
        pred = torch.full((R,), -1, dtype=torch.int32, device=self.device)
        eng.fwd_bwd(self.patch, 0, fe_mode, loss, self.grad, scalars, pred, forward_only=True)
        return scalars.cpu(), self._full_vocab_pred(eng, R, pred).cpu()


def _decoded_pairs(pred_ids: torch.Tensor, labels: torch.Tensor):
    """(pred, gt) continuous actions of the supervised ACTION tokens, in (sample, position) order."""
    sup = labels[:, 1:][labels[:, 1:] != IGNORE_INDEX]
    act = sup > lab.ACTION_TOKEN_BEGIN_IDX
    gt = torch.tensor(lab.decode_token_ids_to_actions(sup[act].numpy()))
    pr = torch.tensor(lab.decode_token_ids_to_actions(pred_ids.long()[act].numpy()))
    return pr, gt


class _AttackerBase(object):
    KIND = "UADA"

    STATE_FILE = "attack_state.pt"

    def __init__(self, vla, processor=None, save_dir="", optimizer="pgd", resize_patch=False, cfg=None, device=None,
                 engine_factory=None, resume=None):
        if device is None:   # the reference works on the device the caller moved the model to (``self.vla.device``, UADA.py:55)
            device = getattr(vla, "device", None) or "cuda:0"
        self.vla = vla
        self.processor = processor
        self.save_dir = save_dir
        # resume: a run sub-directory written by this class (e.g. <save_dir>/last) or its attack_state.pt; env VLA_ATTACK_RESUME
        self.resume = resume or os.environ.get("VLA_ATTACK_RESUME") or None
        self.optimizer = optimizer
        if resize_patch:
            raise NotImplementedError("resize_patch=True is dead code in the reference (appply_random_transform.py:113-118)")
        self.host = AttackEngineHost(vla, cfg=cfg, device=device, engine_factory=engine_factory)
        self.action_tokenizer = lab.ActionTokenizer(getattr(processor, "tokenizer", None))   # UADA.py:60
        self.mean = [torch.tensor(NORM_MEAN[0]), torch.tensor(NORM_MEAN[1])]
        self.std = [torch.tensor(NORM_STD[0]), torch.tensor(NORM_STD[1])]
        self.randomPatchTransform = RandomPatchTransform(device, resize_patch)   # reference attribute; used by eval scripts
        self.loss_buffer = []
        self.val_every = 100
        self.val_batches = 1000

    # reference helpers kept under their names
    def mask_labels(self, labels, maskidx):
        return lab.mask_labels_uada(labels, maskidx)

    def _save_patch(self, patch, sub, outer_iter=None, sched_step=0):
        d = os.path.join(self.save_dir, sub)
        os.makedirs(d, exist_ok=True)
        torch.save(patch.detach().cpu(), os.path.join(d, "patch.pt"))     # fp32 [3,h,w] CPU tensor, as the reference
        if outer_iter is not None:   # restart state next to it (not in the reference's run-dir format; ignored by its consumers)
            torch.save(self.host.state_dict(outer_iter, sched_step, getattr(self, "_acc", None)), os.path.join(d, self.STATE_FILE))
        return d

    def _dump_val_images(self, d, extra=None):
        """``<dir>/val_related_data/<o>.png``: the last validation batch with the patch applied, de-normalised from the DINOv2
        channels (UADA.py:262-270, UPA.py:252-260, TMA.py:351-361), plus optional tensors (TMA: decoded actions).  The
        reference converts with ``ToPILImage`` (``mul(255).byte()``, which wraps values a rounding error above 1.0); the
        values are clamped to [0, 1] here."""
        eng = self.host.engine
        if eng is None or not hasattr(eng, "tap"):
            return
        from PIL import Image
        vd = os.path.join(d, "val_related_data")
        os.makedirs(vd, exist_ok=True)
        img = self.host.cfg.img
        px = eng.tap("px").view(-1, 6, img, img)[:, 0:3].float().cpu()
        px = self.randomPatchTransform.denormalize(px, self.mean[0], self.std[0]).clamp_(0, 1)
        for o in range(px.shape[0]):
            Image.fromarray(px[o].mul(255).byte().permute(1, 2, 0).numpy()).save(os.path.join(vd, f"{o}.png"))
        for name, t in (extra or {}).items():
            torch.save(t, os.path.join(vd, name))

    def _maybe_resume(self):
        """-> (first outer iteration to run, scheduler step).  (0, 0) without ``resume``."""
        if not self.resume:
            return 0, 0
        path = self.resume
        if os.path.isdir(path):
            path = os.path.join(path, self.STATE_FILE)
        st = torch.load(path, map_location="cpu", weights_only=True)
        return self.host.load_state_dict(st, getattr(self, "_acc", None))

    def _dump(self, **lists):
        os.makedirs(self.save_dir, exist_ok=True)
        for name, values in lists.items():
            with open(os.path.join(self.save_dir, f"{name}.pkl"), "wb") as f:
                pickle.dump(values, f)

    # ---- the reference's helper methods, same names and argument meaning (they are not on the engine's hot path: inside
    # patchattack_unconstrained the engine computes all of this on the device) -----------------------------------------
    SAVE_INFO_LISTS: tuple = ()

    def save_info(self, path):
        """``<path>/<list>.pkl`` for every metric list of the attack (UADA.py:341-353, UPA.py:309-325, TMA.py:454-468)."""
        os.makedirs(path, exist_ok=True)
        for name in self.SAVE_INFO_LISTS:
            with open(os.path.join(path, f"{name}.pkl"), "wb") as f:
                pickle.dump(getattr(self, name, []), f)

    def plot_loss(self):
        """Loss curves as ``<save_dir>/loss.png`` (UADA.py:76-91); skipped with a note when matplotlib is not installed."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            print("plot_loss: matplotlib is not installed; the .pkl lists written by save_info() hold the same data")
            return None
        names = [n for n in self.SAVE_INFO_LISTS if len(getattr(self, n, []))]
        fig, axes = plt.subplots(1, max(len(names), 1), figsize=(5 * max(len(names), 1), 4))
        for ax, n in zip(np.atleast_1d(axes), names):
            ax.plot(getattr(self, n))
            ax.set_title(n)
        out = os.path.join(self.save_dir or ".", "loss.png")
        fig.savefig(out)
        plt.close(fig)
        return out

    def modifiy_labels(self, labels, target_action={"0": 0, "1": 1, "2": 2, "3": 3, "4": 4, "5": 5, "6": 6, "7": 7, "8": 8}):
        """Write ``target_action[k]`` at the k-th supervised position of every row, entries equal to -100 left alone
        (UADA.py:295-307, TMA.py:385-396; the reference's spelling of the name is kept).  In place, returns ``labels``."""
        first = (labels != IGNORE_INDEX).int().argmax(dim=1)
        for key, value in target_action.items():
            if value != IGNORE_INDEX:
                labels[torch.arange(labels.shape[0]), first + int(key)] = value
        return labels

    def calculate_relative_distance(self, pred, gt, maskidx, relative_distance):
        """Append ``|pred - gt| / max(1 - gt, gt + 1)`` of every (sample, DoF) to ``relative_distance[str(dof)]``
        (UADA.py:354-369, UPA.py:327-342); ``pred`` / ``gt`` are flat continuous actions, ``len(maskidx)`` per sample."""
        n = len(maskidx)
        rd = lab.relative_distance(torch.as_tensor(pred, dtype=torch.float64).view(-1, n), torch.as_tensor(gt, dtype=torch.float64).view(-1, n))
        for j, dof in enumerate(maskidx):
            relative_distance[str(dof)].extend(rd[:, j].tolist())
        return relative_distance

    def cal_UAD(self, pred, gt):
        """Untargeted action discrepancy of predicted vs ground-truth action TOKEN ids: mean of |a_pred - a_gt| over the distance
        from a_gt to the far end of [-1, 1] (UADA.py:408-418)."""
        a_gt = torch.tensor(lab.decode_token_ids_to_actions(torch.as_tensor(gt).detach().cpu().numpy()))
        a_pr = torch.tensor(lab.decode_token_ids_to_actions(torch.as_tensor(pred).detach().cpu().numpy()))
        far = torch.where(a_gt > 0, (a_gt + 1).abs(), (a_gt - 1).abs())
        return ((a_pr - a_gt).abs() / far).mean()

    @staticmethod
    def _log(args, data, step):
        if args is not None and getattr(args, "wandb_project", "false") != "false":
            try:
                import wandb
                wandb.log(data, step=step)
            except ImportError:
                pass

    @staticmethod
    def filter_train(data):
        """``filter_train`` of the reference (UADA.py:311-341, UPA.py / TMA.py alike), used for the gripper-only attack with
        ``filterGripTrainTo1``: keep the samples whose gripper label is 31744 ("1") when there are 2..7 of them, a random
        8 of them (``random.sample``: consumes the Python RNG stream, as in the reference) when there are more than 8, and
        the batch unchanged otherwise (0, 1 or exactly 8 hits fall through every branch of the reference)."""
        labels = data["labels"]
        sup = labels[labels > lab.ACTION_TOKEN_BEGIN_IDX].view(-1, 7)
        one_index = [i for i in range(sup.shape[0]) if int(sup[i, 6]) == 31744]
        if 1 < len(one_index) < 8:
            chosen = one_index
        elif len(one_index) > 8:
            chosen = random.sample(one_index, k=8)
        else:
            return data
        out = dict(data)
        for k in ("labels", "attention_mask", "input_ids"):
            out[k] = data[k][chosen, :]
        pv = data["pixel_values"]
        out["pixel_values"] = [pv[i] for i in chosen] if isinstance(pv, (list, tuple)) else pv[chosen]
        return out

    def _open(self, loader, restart=True):
        """``LookaheadLoader`` over ``loader`` (``VLA_PREFETCH=0``: no lookahead, batches are fetched when needed)."""
        if loader is None:
            return None
        return LookaheadLoader(loader, self.host.cfg.img, restart=restart, lookahead=os.environ.get("VLA_PREFETCH", "1") != "0")

    @staticmethod
    def _close(*iterators):
        for it in iterators:
            if isinstance(it, LookaheadLoader):
                it.close()

    @staticmethod
    def _next(iterator, loader):
        return iterator.next(), iterator

    def _drain(self, pending, process):
        """Process the read-back of the previous outer iteration (if any).  The loops launch iteration i + 1 BEFORE they look at
        iteration i's numbers (logging, metric lists): the host-side bookkeeping then overlaps the device's work, and the
        order of everything that is logged or appended is unchanged."""
        if pending is not None:
            process(pending)
        return None

    def _lookahead(self, i, train_it, has_val=True):
        """Host work to overlap with the device's inner loop: fetch the next training batch -- except on iterations that end
        with a validation pass, whose ``next()`` calls on the validation loader come first in the reference's RNG order."""
        if i % self.val_every == 0 and has_val:
            return None
        return train_it.prefetch


class UADAAttacker(_AttackerBase):
    """Untargeted action-discrepancy attack (UADA.py).  loss = mean((5e - 5t)^2) + 1/CE."""
    KIND = "UADA"
    SAVE_INFO_LISTS = ("train_CE_loss", "train_MSE_distance_loss", "train_UAD", "val_CE_loss", "val_MSE_Distance", "val_UAD")

    def weighted_loss(self, logits, labels, maskid=None):
        """``(distance_loss, UAD)`` of UADA.py:381-406 on caller-supplied CUDA logits ``[B, L, V]`` and text labels ``[B, T]``
        (CUDA loss-head kernel, differentiable w.r.t. the logits); ``maskid`` is unused, as in the reference."""
        from .loss_heads import uada_weighted_loss
        return uada_weighted_loss(logits, labels, 5.0)

    def patchattack_unconstrained(self, train_dataloader, val_dataloader, num_iter=5000, target_action=np.zeros(7),
                                  patch_size=[3, 50, 50], lr=1 / 255, accumulate_steps=1, maskidx=[], warmup=20,
                                  filterGripTrainTo1=False, geometry=False, innerLoop=1, args=None):
        h = self.host
        self.val_CE_loss, self.val_MSE_Distance, self.val_UAD = [], [], []
        self.train_CE_loss, self.train_MSE_distance_loss, self.train_UAD = [], [], []
        self.MSE_Distance_best = 10000
        h.init_patch(patch_size)
        loss = LossSpec(_lib.LOSS_UADA, mse_weight=5.0)
        fe_mode = _lib.FE_WARP if geometry else _lib.FE_PASTE20
        opt_kind = _lib.OPT_ADAMW if self.optimizer == "adamW" else _lib.OPT_PGD
        total = int(num_iter / accumulate_steps)
        start_iter, sched_step = self._maybe_resume()
        self._sched_step = sched_step
        train_it = self._open(train_dataloader)
        val_it = self._open(val_dataloader)

        def process(job):      # logging of one outer iteration (UADA.py:165-186)
            i, labels, cur_lr = job.ctx
            scalars, pred = job.result()
            self.train_CE_loss += scalars[:, _lib.S_CE].tolist()
            self.train_MSE_distance_loss += scalars[:, _lib.S_LOSS].tolist()
            self.train_UAD += scalars[:, _lib.S_UAD].tolist()
            pr, gt = _decoded_pairs(pred, labels)
            rd = lab.relative_distance(pr, gt).view(-1, max(1, len(maskidx)))
            log = {"TRAIN_attack_loss(CE)": scalars[-1, _lib.S_CE].item(),
                   "TRAIN_patch_gradient": scalars[-1, _lib.S_GRAD_MEAN].item(),
                   "TRAIN_LR": cur_lr,
                   "TRAIN_attack_loss (MSE_Distance)": scalars[-1, _lib.S_LOSS].item(),
                   "TRAIN_UAD": scalars[-1, _lib.S_UAD].item()}
            for k, idx in enumerate(maskidx):
                log[f"train_rd_{idx}"] = rd[:, k].mean().item()
            self.loss_buffer.append(log["TRAIN_attack_loss (MSE_Distance)"])
            self._log(args, log, i)

        pending = None
        for i in range(start_iter, num_iter):
            data, train_it = self._next(train_it, train_dataloader)
            data = dict(data)
            if filterGripTrainTo1 and len(maskidx) == 1 and maskidx[0] == 6:
                data = self.filter_train(data)
            data["labels"] = self.mask_labels(data["labels"].clone(), maskidx)
            cur_lr = lr * cosine_with_warmup(sched_step, warmup, total) if self.optimizer == "adamW" else lr
            job = h.submit_inner_loop(data, innerLoop, fe_mode, loss, cur_lr, opt_kind,
                                      after_launch=self._lookahead(i, train_it, val_dataloader is not None))
            job.ctx = (i, data["labels"], cur_lr)
            if self.optimizer == "adamW" and ((i + 1) % accumulate_steps == 0):
                sched_step += 1
            self._sched_step = sched_step
            pending = self._drain(pending, process)          # iteration i - 1, while the device runs iteration i
            pending = job
            if i % self.val_every == 0:
                pending = self._drain(pending, process)
                if val_dataloader is not None:
                    val_it = self._validate(i, val_it, val_dataloader, maskidx, fe_mode, loss, args)
                elif self.save_dir:
                    self._save_patch(h.patch, "last", outer_iter=i, sched_step=sched_step)
        self._drain(pending, process)
        self._close(train_it, val_it)
        return h.patch.detach().cpu()

    def _validate(self, i, val_it, val_dataloader, maskidx, fe_mode, loss, args):
        h = self.host
        n, s_mse, s_uad, s_ce = 0, 0.0, 0.0, 0.0
        rds = []
        for _ in range(self.val_batches):
            data, val_it = self._next(val_it, val_dataloader)
            data = dict(data)
            data["labels"] = self.mask_labels(data["labels"].clone(), maskidx)
            n += data["labels"].shape[0]
            sc, pred = h.evaluate(data, fe_mode, loss)
            s_mse += sc[_lib.S_AUX0].item()
            s_uad += sc[_lib.S_UAD].item()
            s_ce += sc[_lib.S_CE].item()
            pr, gt = _decoded_pairs(pred, data["labels"])
            rds.append(lab.relative_distance(pr, gt).view(-1, max(1, len(maskidx))))
        avg_mse, avg_uad, avg_ce = s_mse / n, s_uad / n, s_ce / n     # the reference divides by the sample count
        log = {"VAL_MSE_Distance": avg_mse, "VAL_UAD": avg_uad}
        rd = torch.cat(rds)
        for k, idx in enumerate(maskidx):
            log[f"val_rd_{idx}"] = rd[:, k].mean().item()
        self._log(args, log, i)
        if avg_mse < self.MSE_Distance_best:
            self.MSE_Distance_best = avg_mse
            self._dump_val_images(self._save_patch(h.patch, str(i)))
        self._dump_val_images(self._save_patch(h.patch, "last", outer_iter=i, sched_step=getattr(self, "_sched_step", 0)))
        self.val_CE_loss.append(avg_ce)
        self.val_MSE_Distance.append(avg_mse)
        self.val_UAD.append(avg_uad)
        self._dump(train_CE_loss=self.train_CE_loss, train_MSE_distance_loss=self.train_MSE_distance_loss,
                   train_UAD=self.train_UAD, val_CE_loss=self.val_CE_loss, val_MSE_Distance=self.val_MSE_Distance,
                   val_UAD=self.val_UAD)
        return val_it


class UPAAttacker(_AttackerBase):
    """Untargeted position-aware attack (UPA.py)."""
    KIND = "UPA"

    def __init__(self, vla, processor=None, save_dir="", optimizer="pgd", resize_patch=False, alpha=0.5, belta=0.5, **kw):
        super().__init__(vla, processor, save_dir, optimizer, resize_patch, **kw)
        self.alpha, self.belta = alpha, belta
        self.val_batches = 100

    SAVE_INFO_LISTS = ("val_CE_loss", "train_CE_loss", "avg_angle_loss", "avg_distance_loss", "avg_reserve_loss")

    def mask_labels(self, labels, maskidx):
        return lab.mask_labels_upa(labels, maskidx)

    def change_target(self, gt):
        """UPA.py:358-364 (``guide`` labels), in place."""
        return lab.change_target(gt)

    def weighted_loss(self, logits, labels):
        """``(total_loss, angle_loss, distance_loss)`` of UPA.py:367-387 on caller-supplied CUDA logits (CUDA loss-head kernel,
        differentiable w.r.t. the logits) with this attacker's ``alpha`` / ``belta``."""
        from .loss_heads import upa_weighted_loss
        return upa_weighted_loss(logits, labels, self.alpha, self.belta)

    def patchattack_unconstrained(self, train_dataloader, val_dataloader, num_iter=5000, target_action=np.zeros(7),
                                  patch_size=[3, 50, 50], lr=1 / 255, accumulate_steps=1, maskidx=[], warmup=20,
                                  filterGripTrainTo1=False, geometry=False, guide=False, innerLoop=1,
                                  reverse_direction=True, args=None):
        h = self.host
        self.train_CE_loss, self.val_CE_loss = [], []
        self.avg_reserve_loss, self.avg_angle_loss, self.avg_distance_loss = [], [], []
        self.reverse_direction_loss = 1e8
        val_it = self._open(val_dataloader)
        h.init_patch(patch_size)
        if guide:
            loss = LossSpec(_lib.LOSS_CE, ce_scale=1.0)
        elif reverse_direction:
            loss = LossSpec(_lib.LOSS_UPA, alpha=self.alpha, belta=self.belta)
        else:
            loss = LossSpec(_lib.LOSS_NEG_CE)
        fe_mode = _lib.FE_WARP if geometry else _lib.FE_PASTE20
        opt_kind = _lib.OPT_ADAMW if self.optimizer == "adamW" else _lib.OPT_PGD
        total = int(num_iter / accumulate_steps)
        acc = self._acc = torch.zeros_like(h.patch) if accumulate_steps > 1 else None
        start_iter, sched_step = self._maybe_resume()
        train_it = self._open(train_dataloader)

        def process(job):      # logging of one outer iteration (UPA.py:170-191)
            i, cur_lr = job.ctx
            scalars, _ = job.result()
            log = {"TRAIN_attack_loss(CE)": scalars[-1, _lib.S_LOSS].item(),
                   "TRAIN_patch_gradient": scalars[-1, _lib.S_GRAD_MEAN].item(), "TRAIN_LR": cur_lr,
                   "TRAIN_ANGLE_LOSS": scalars[-1, _lib.S_AUX0].item(), "TRAIN_DISTANCE_LOSS": scalars[-1, _lib.S_AUX1].item()}
            self.train_CE_loss.append(log["TRAIN_attack_loss(CE)"])
            self.loss_buffer.append(log["TRAIN_attack_loss(CE)"])
            self._log(args, log, i)

        pending = None
        for i in range(start_iter, num_iter):
            data, train_it = self._next(train_it, train_dataloader)
            data = dict(data)
            if filterGripTrainTo1 and len(maskidx) == 1 and maskidx[0] == 6:
                data = self.filter_train(data)
            labels = data["labels"].clone()
            if not reverse_direction:
                labels = self.mask_labels(labels, maskidx)
            if guide:
                labels = lab.change_target(labels)
            data["labels"] = labels
            stepping = (i + 1) % accumulate_steps == 0
            cur_lr = lr * cosine_with_warmup(sched_step, warmup, total) if self.optimizer == "adamW" else lr
            job = h.submit_inner_loop(data, innerLoop, fe_mode, loss, cur_lr, opt_kind,
                                      clip_l1=1e-3 if self.optimizer == "adamW" else 0.0, do_step=stepping, accumulate=acc,
                                      after_launch=self._lookahead(i, train_it, val_dataloader is not None and reverse_direction and not guide))
            job.ctx = (i, cur_lr)
            if self.optimizer == "adamW" and stepping:
                sched_step += 1
            pending = self._drain(pending, process)          # iteration i - 1, while the device runs iteration i
            pending = job
            if i % self.val_every == 0:
                pending = self._drain(pending, process)
                if val_dataloader is not None and reverse_direction and not guide:
                    val_it = self._validate(i, val_it, val_dataloader, fe_mode, loss, sched_step, args)
                else:
                    self._save_patch(h.patch, "last", outer_iter=i, sched_step=sched_step)
                self._dump(train_CE_loss=self.train_CE_loss)
        self._drain(pending, process)
        self._close(train_it, val_it)
        return h.patch.detach().cpu()

    def _validate(self, i, val_it, val_dataloader, fe_mode, loss, sched_step, args):
        """UPA.py:193-275: 100 forward-only batches; sums of the per-batch reverse-direction loss / angle term / distance term
        divided by the SAMPLE count (as the reference does); the patch with the lowest reverse-direction loss is kept."""
        h = self.host
        n, s_loss, s_ang, s_dist = 0, 0.0, 0.0, 0.0
        for _ in range(self.val_batches):
            data, val_it = self._next(val_it, val_dataloader)
            data = dict(data)
            n += data["labels"].shape[0]
            sc, _ = h.evaluate(data, fe_mode, loss)
            s_loss += sc[_lib.S_LOSS].item()
            s_ang += sc[_lib.S_AUX0].item()
            s_dist += sc[_lib.S_AUX1].item()
        avg_loss, avg_ang, avg_dist = s_loss / n, s_ang / n, s_dist / n
        self._log(args, {"reverse_direction_loss": avg_loss, "avg_angle_loss": avg_ang, "avg_distance_loss": avg_dist}, i)
        if avg_loss < self.reverse_direction_loss:
            self.reverse_direction_loss = avg_loss
            self._dump_val_images(self._save_patch(h.patch, str(i)))
        self._dump_val_images(self._save_patch(h.patch, "last", outer_iter=i, sched_step=sched_step))
        # The reference appends the already averaged values divided by the sample count a second time (UPA.py:269-271); the
        # .pkl lists reproduce that so that they hold the numbers the reference writes (the logged values and the best-patch
        # selection above use the plain averages, as the reference's do).
        self.avg_reserve_loss.append(avg_loss / n)
        self.avg_angle_loss.append(avg_ang / n)
        self.avg_distance_loss.append(avg_dist / n)
        self._dump(val_CE_loss=self.val_CE_loss, avg_angle_loss=self.avg_angle_loss, avg_distance_loss=self.avg_distance_loss,
                   avg_reserve_loss=self.avg_reserve_loss)
        return val_it


class TMAAttacker(_AttackerBase):
    """Targeted manipulation attack (TMA.py): CE towards a target action on the DoF in ``maskidx``."""
    KIND = "TMA"
    SAVE_INFO_LISTS = ("val_CE_loss", "val_L1_loss", "val_ASR", "val_inner_relatived_distance", "train_CE_loss", "train_inner_avg_loss",
                       "train_inner_relatived_distance")

    def calculate_01_ASR(self, pred, gt):
        """Gripper flip counts on token ids (TMA.py:398-420): (0 -> other, #gt 0, 1 -> other, #gt 1, other -> 0, #gt other) with
        31872 = the zero action and 31744 = +1."""
        pred, gt = torch.as_tensor(pred).view(-1), torch.as_tensor(gt).view(-1)
        is0, is1 = gt == 31872, gt == 31744
        other = ~is0 & ~is1
        return (int((is0 & (pred != 31872)).sum()), int(is0.sum()), int((is1 & (pred != 31744)).sum()), int(is1.sum()),
                int((other & (pred == 31872)).sum()), int(other.sum()))

    def calculate_relative_distance_target(self, pred, gt):
        """Mean of ``|pred - gt| / max(1 - gt, gt + 1)`` over the entries (TMA.py:470-483)."""
        return lab.relative_distance(torch.as_tensor(pred, dtype=torch.float64), torch.as_tensor(gt, dtype=torch.float64)).mean()

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.val_batches = 100

    def patchattack_unconstrained(self, train_dataloader, val_dataloader, num_iter=5000, target_action=np.zeros(7),
                                  patch_size=[3, 50, 50], alpha=1 / 255, accumulate_steps=1, maskidx=[], warmup=20,
                                  filterGripTrainTo1=False, geometry=False, colorjitter=False, innerLoop=1, args=None):
        h = self.host
        self.train_CE_loss, self.train_inner_avg_loss, self.train_inner_relatived_distance = [], [], []
        self.val_CE_loss, self.val_L1_loss, self.val_ASR, self.val_inner_relatived_distance = [], [], [], []
        self.min_val_avg_L1_loss = 1e8
        val_it = self._open(val_dataloader)
        h.init_patch(patch_size)
        target = lab.tma_target(target_action, maskidx)
        loss = LossSpec(_lib.LOSS_CE, ce_scale=1.0 / accumulate_steps)
        fe_mode = _lib.FE_WARP if geometry else _lib.FE_FIX          # paste_patch_fix when no geometry (TMA.py:133-135)
        opt_kind = _lib.OPT_ADAMW if self.optimizer == "adamW" else _lib.OPT_PGD
        total = int(num_iter / accumulate_steps)
        acc = self._acc = torch.zeros_like(h.patch) if accumulate_steps > 1 else None
        start_iter, sched_step = self._maybe_resume()
        train_it = self._open(train_dataloader)

        def process(job):      # logging of one outer iteration (TMA.py:176-200)
            i, labels, cur_lr = job.ctx
            scalars, pred = job.result()
            # logging only: the reference averages this metric over the inner steps (TMA.py:161,177); reading the predictions of
            # every inner step would put a device->host sync into the inner loop, so the last inner step's are used
            pr, gt = _decoded_pairs(pred, labels)
            rd = lab.relative_distance(pr, gt).mean().item() if pr.numel() else 0.0
            log = {"TRAIN_attack_loss(CE)": scalars[-1, _lib.S_LOSS].item(),
                   "TRAIN_patch_gradient": scalars[-1, _lib.S_GRAD_MEAN].item(), "TRAIN_LR": cur_lr,
                   "TRAIN_inner_avg_loss": scalars[:, _lib.S_LOSS].mean().item(), "TRAIN_inner_relatived_distance": rd}
            self.train_CE_loss.append(log["TRAIN_attack_loss(CE)"])
            self.train_inner_avg_loss.append(log["TRAIN_inner_avg_loss"])
            self.train_inner_relatived_distance.append(rd)
            self.loss_buffer.append(log["TRAIN_attack_loss(CE)"])
            self._log(args, log, i)

        pending = None
        for i in range(start_iter, num_iter):
            data, train_it = self._next(train_it, train_dataloader)
            data = dict(data)
            if filterGripTrainTo1 and len(maskidx) == 1 and maskidx[0] == 6:
                data = self.filter_train(data)
            data["labels"] = lab.tma_labels(data["labels"], target)
            stepping = (i + 1) % accumulate_steps == 0
            cur_lr = alpha * cosine_with_warmup(sched_step, warmup, total) if self.optimizer == "adamW" else alpha
            job = h.submit_inner_loop(data, innerLoop, fe_mode, loss, cur_lr, opt_kind, do_step=stepping, accumulate=acc,
                                      after_launch=self._lookahead(i, train_it, val_dataloader is not None))
            job.ctx = (i, data["labels"], cur_lr)
            if self.optimizer == "adamW" and stepping:
                sched_step += 1
            pending = self._drain(pending, process)          # iteration i - 1, while the device runs iteration i
            pending = job
            if i % self.val_every == 0:
                pending = self._drain(pending, process)
                if val_dataloader is not None:
                    val_it = self._validate(i, val_it, val_dataloader, target, maskidx, fe_mode, sched_step, args)
                else:
                    self._save_patch(h.patch, "last", outer_iter=i, sched_step=sched_step)
                self._dump(train_CE_loss=self.train_CE_loss, train_inner_avg_loss=self.train_inner_avg_loss,
                           train_inner_relatived_distance=self.train_inner_relatived_distance)
        self._drain(pending, process)
        self._close(train_it, val_it)
        return h.patch.detach().cpu()

    def _validate(self, i, val_it, val_dataloader, target, maskidx, fe_mode, sched_step, args):
        """TMA.py:202-383: 100 forward-only batches with the labels replaced by the target; CE, L1 between the decoded
        greedy prediction and the target, ASR (all attacked DoF hit exactly), relative distance to the target; sums divided
        by the sample count as in the reference; the patch with the lowest L1 is kept.  For the gripper-only attack
        (maskidx == [6]) the reference first drops the samples whose CLEAN prediction of the gripper token is wrong and
        reports the 0->other / 1->other / other->0 rates (calculate_01_ASR); both are reproduced."""
        h = self.host
        grip = len(maskidx) == 1 and maskidx[0] == 6
        ce_loss = LossSpec(_lib.LOSS_CE, ce_scale=1.0)
        n, s_ce, s_l1, s_rd, hits = 0, 0.0, 0.0, 0.0, 0
        c01 = [0, 0, 0, 0, 0, 0]
        pr = gt = torch.zeros(0)
        for _ in range(self.val_batches):
            data, val_it = self._next(val_it, val_dataloader)
            data = dict(data)
            real = data["labels"]
            if grip:
                _, clean_pred = h.evaluate(data, _lib.FE_NONE, ce_loss)                 # im_process: no patch
                sup = real[:, 1:][real[:, 1:] != IGNORE_INDEX].view(real.shape[0], -1)  # [B, 8] = 7 actions + EOS
                keep = [b for b in range(real.shape[0]) if int(clean_pred.view(real.shape[0], -1)[b, 6]) == int(sup[b, 6])]
                if not keep:
                    continue
                data = {k: ([v[b] for b in keep] if isinstance(v, list) else v[keep]) for k, v in data.items()
                        if k in ("pixel_values", "input_ids", "attention_mask", "labels")}
                real = data["labels"]
            B = real.shape[0]
            n += B
            data["labels"] = lab.tma_labels(real, target)
            sc, pred = h.evaluate(data, fe_mode, ce_loss)
            pr, gt = _decoded_pairs(pred, data["labels"])
            s_ce += sc[_lib.S_CE].item()
            s_l1 += torch.nn.functional.l1_loss(pr, gt).item() if pr.numel() else 0.0
            s_rd += float(lab.relative_distance(pr, gt).mean()) if pr.numel() else 0.0
            k = max(1, len(maskidx) - (1 if 7 in maskidx else 0))
            hits += int((pr.view(B, k) == gt.view(B, k)).all(dim=1).sum()) if pr.numel() else 0
            if grip:
                p_ids = pred.view(B, -1)[:, 0]
                g_ids = real[:, 1:][real[:, 1:] != IGNORE_INDEX].view(B, -1)[:, 6]
                for b in range(B):
                    g_, p_ = int(g_ids[b]), int(p_ids[b])
                    if g_ == 31872:
                        c01[1] += 1
                        c01[0] += p_ != 31872
                    elif g_ == 31744:
                        c01[3] += 1
                        c01[2] += p_ != 31744
                    else:
                        c01[5] += 1
                        c01[4] += p_ == 31872
        n = max(n, 1)
        avg_ce, avg_l1, asr, avg_rd = s_ce / n, s_l1 / n, hits / n, s_rd / n
        log = {"VAL_avg_CE_loss": avg_ce, "VAL_avg_L1_loss": avg_l1}
        if grip:
            log.update({"VAL_ASR(pred0-AllCorrect)": asr, "ASR_02other": c01[0] / c01[1] if c01[1] else 0,
                        "ASR_12other": c01[2] / c01[3] if c01[3] else 0, "ASR_other20": c01[4] / c01[5] if c01[5] else 0,
                        "ALL_ASR_6": (c01[0] + c01[2]) / (c01[1] + c01[3]) if (c01[1] + c01[3]) else 0})
        else:
            log.update({"VAL_ASR": asr, "VAL_inner_relatived_distance": avg_rd})
        self._log(args, log, i)
        extra = {"continuous_actions_pred.pt": pr, "continuous_actions_gt.pt": gt}
        if avg_l1 < self.min_val_avg_L1_loss:
            self.min_val_avg_L1_loss = avg_l1
            self._dump_val_images(self._save_patch(h.patch, str(i)), extra)
        self._dump_val_images(self._save_patch(h.patch, "last", outer_iter=i, sched_step=sched_step), extra)
        self.val_CE_loss.append(avg_ce)
        self.val_L1_loss.append(avg_l1)
        self.val_ASR.append(asr)
        self.val_inner_relatived_distance.append(avg_rd)
        self._dump(val_CE_loss=self.val_CE_loss, val_L1_loss=self.val_L1_loss, val_ASR=self.val_ASR,
                   val_inner_relatived_distance=self.val_inner_relatived_distance)
        return val_it


class UADADDPAttacker(_AttackerBase):
    """UADA with the batch sharded over ranks (UADA_ddp.py): one process per GPU, each rank runs the engine on its
    shard, the patch gradient is all-reduced (mean) every inner step, the replicated update keeps patches identical.
    Constructor mirrors UADA_ddp.py:37 except that the model / dataset are passed in instead of loaded by path."""
    KIND = "UADA_DDP"

    def __init__(self, vla_path, dataset_name=None, save_dir="", resize_patch=False, patch_size=[3, 50, 50], lr=0.01, bs=1,
                 warmup=20, num_iter=10000, maskidx=[], innerLoop=1, geometry=True, use_wandb=True, MSE_weights=1,
                 cfg=None, device=None, engine_factory=None, backend="nccl", resume=None, dataloaders=None):
        """Positional / keyword compatible with UADA_ddp.py:37.  ``vla_path``: what the reference passes (a hub id or checkpoint
        directory, loaded like the reference does) or a ``torch.save``d state dict file, a state dict, an HF module or a loaded
        ``VLAEngine``.  ``dataset_name``: the RLDS dataset name (resolved through the caller's ``white_patch.openvla_dataloader
        .get_dataset`` and sharded over ranks as UADA_ddp.py:157-160 does), or ``(train_loader, val_loader)``, or a callable
        ``(rank, world_size) -> (train_loader, val_loader)``."""
        rank = int(os.environ.get("LOCAL_RANK", 0))
        device = device or f"cuda:{rank}"
        if cfg is None and isinstance(vla_path, (str, os.PathLike)) and not os.path.isfile(vla_path):
            from .weights import resolve_vla
            vla_path, cfg = resolve_vla(vla_path)        # -> state dict + engine configuration: loaded once, here, like the reference
        super().__init__(vla_path, None, save_dir, "adamW", resize_patch, cfg=cfg, device=device, engine_factory=engine_factory, resume=resume)
        self.backend = backend
        self.dataloaders = dataloaders if dataloaders is not None else dataset_name
        self.patch_size, self.lr, self.bs, self.warmup, self.num_iter = patch_size, lr, bs, warmup, num_iter
        self.maskidx, self.innerLoop, self.geometry, self.use_wandb, self.MSE_weights = maskidx, innerLoop, geometry, use_wandb, MSE_weights
        self.val_every, self.val_batches = 200, 100
        self.val_CE_loss, self.val_MSE_Distance, self.val_UAD = [], [], []
        self.MSE_Distance_best = 10000

    def _resolve_dataloaders(self, rank, world_size):
        """-> (train_loader, val_loader | None) for this rank."""
        d = self.dataloaders
        if d is None:
            raise ValueError("no data: pass dataset_name (an RLDS name, a (train, val) pair or a callable (rank, world) -> pair)")
        if callable(d):
            d = d(rank, world_size)
        if isinstance(d, str):
            # the reference's own data path (UADA_ddp.py:52,157-160): get_dataset(name) -> RLDS datasets, sharded by rank, batched
            # by a torch DataLoader with the padded collator; all of it is the caller's code (out of this repo's scope)
            from white_patch.openvla_dataloader import get_dataset
            train_ds, val_ds = get_dataset(dataset=d)
            try:
                from prismatic.util.data_utils import PaddedCollatorForActionPrediction
                from transformers import AutoProcessor
                tok = AutoProcessor.from_pretrained(self.vla if isinstance(self.vla, str) else "openvla/openvla-7b", trust_remote_code=True).tokenizer
                collator = PaddedCollatorForActionPrediction(tok.model_max_length, tok.pad_token_id, padding_side="right")
            except ImportError:
                collator = None
            mk = lambda ds: torch.utils.data.DataLoader(ds.shard(num_shards=world_size, index=rank) if hasattr(ds, "shard") else ds,  # noqa: E731
                                                        batch_size=self.bs, collate_fn=collator)
            d = (mk(train_ds), mk(val_ds) if val_ds is not None else None)
        d = tuple(d)
        return d[0], (d[1] if len(d) > 1 else None)

    def setup(self, rank, world_size):
        import torch.distributed as dist
        if world_size > 1 and not dist.is_initialized():
            dist.init_process_group(self.backend, rank=rank, world_size=world_size)
        if self.host.device.type == "cuda":
            torch.cuda.set_device(self.host.device)
        self.host.rank, self.host.world_size = rank, world_size

    def weighted_loss(self, logits, labels, maskid=None):
        """``(distance_loss, UAD)`` of UADA_ddp.py:99-124 with this attacker's ``MSE_weights`` (CUDA loss-head kernel)."""
        from .loss_heads import uada_weighted_loss
        return uada_weighted_loss(logits, labels, float(self.MSE_weights))

    def cleanup(self):
        """UADA_ddp.py:134-136."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()

    @classmethod
    def run(cls, vla_path, dataset_name, save_dir, resize_patch, patch_size, lr, bs, warmup, num_iter, maskidx, innerLoop, geometry,
            use_wandb, MSE_weights, world_size=None, **extra):
        """UADA_ddp.py:327-336: one process per visible GPU, each constructing its own attacker.  ``vla_path`` is a file holding
        the state dict (every rank loads it itself); ``dataset_name`` is a picklable callable ``(rank, world_size) ->
        (train_loader, val_loader | None)`` standing in for the reference's RLDS dataset name (its TF loader is out of scope).
        ``extra`` (cfg, backend, device, engine_factory, resume) goes to the constructor."""
        import torch.multiprocessing as mp
        world_size = world_size or torch.cuda.device_count()
        instance_params = dict(vla_path=vla_path, dataset_name=dataset_name, save_dir=save_dir, resize_patch=resize_patch, patch_size=patch_size,
                               lr=lr, bs=bs, warmup=warmup, num_iter=num_iter, maskidx=maskidx, innerLoop=innerLoop, geometry=geometry,
                               use_wandb=use_wandb, MSE_weights=MSE_weights, **extra)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        mp.spawn(cls._attack_entry, args=(instance_params, world_size), nprocs=world_size)

    @staticmethod
    def _attack_entry(rank, instance_params, world_size):
        """UADA_ddp.py:338-344."""
        import torch.distributed as dist
        os.environ["LOCAL_RANK"] = str(rank)
        params = dict(instance_params)
        if not dist.is_initialized():
            dist.init_process_group(params.get("backend", "nccl"), rank=rank, world_size=world_size)
        instance = UADADDPAttacker(**params)
        instance.attack(rank, world_size)
        instance.cleanup()

    def attack(self, rank, world_size, train_dataloader=None, val_dataloader=None):
        import torch.distributed as dist
        self.setup(rank, world_size)
        h = self.host
        if train_dataloader is None:
            train_dataloader, val_from_name = self._resolve_dataloaders(rank, world_size)
            val_dataloader = val_dataloader or val_from_name
        h.init_patch(self.patch_size)
        start_iter, _ = self._maybe_resume()   # every rank restores the same replicated state (scheduler step = outer index)
        loss = LossSpec(_lib.LOSS_UADA_DDP, mse_weight=float(self.MSE_weights))
        fe_mode = _lib.FE_WARP if self.geometry else _lib.FE_PASTE20
        logs = []
        train_it = self._open(train_dataloader, restart=False)

        def process(job):
            """The four scalar all-reduces of the reference (UADA_ddp.py:214-221), packed into one exchange.  On a GPU it runs on
            the read-back stream: a collective issued on the compute stream would queue behind the NEXT iteration's kernels."""
            import contextlib
            i, cur_lr = job.ctx
            scalars, _ = job.result()
            last = scalars[-1]
            side = torch.cuda.stream(h._rb_stream) if h._rb_stream is not None else contextlib.nullcontext()
            with side:
                pack = torch.tensor([last[_lib.S_CE], last[_lib.S_LOSS], last[_lib.S_UAD], last[_lib.S_GRAD_MEAN]], device=h.device)
                if world_size > 1:
                    gathered = [torch.zeros_like(pack) for _ in range(world_size)]
                    dist.all_gather(gathered, pack)
                    g = torch.stack(gathered)
                    pack = torch.stack([g[:, 0].mean(), g[:, 1].mean(), g[:, 2].mean(), g[:, 3].max()])
                vals = pack.tolist()
            logs.append({"TRAIN_attack_loss(CE)": vals[0], "TRAIN_attack_loss (MSE_Distance)": vals[1], "TRAIN_UAD": vals[2],
                         "TRAIN_patch_gradient": vals[3], "TRAIN_LR": cur_lr})

        pending = None
        for i in range(start_iter, int(self.num_iter)):
            try:                                   # ``for i, data in enumerate(loader)`` of UADA_ddp.py:176: ends with the loader
                data = train_it.next()
            except StopIteration:
                break
            data = dict(data)
            data["labels"] = self.mask_labels(data["labels"].clone(), self.maskidx)
            cur_lr = self.lr * cosine_with_warmup(i, self.warmup, int(self.num_iter))
            job = h.submit_inner_loop(data, self.innerLoop, fe_mode, loss, cur_lr, _lib.OPT_ADAMW,
                                      after_launch=self._lookahead(i, train_it, val_dataloader is not None))
            job.ctx = (i, cur_lr)
            pending = self._drain(pending, process)          # iteration i - 1, while the device runs iteration i
            pending = job
            if i % self.val_every == 0:
                pending = self._drain(pending, process)
                if val_dataloader is not None:
                    self._validate(i, rank, world_size, val_dataloader, fe_mode, loss)
                elif rank == 0 and self.save_dir:
                    self._save_patch(h.patch, "last", outer_iter=i, sched_step=i + 1)
        self._drain(pending, process)
        self.train_logs = logs
        self._close(train_it)
        return h.patch.detach().cpu()

    def _validate(self, i, rank, world_size, val_dataloader, fe_mode, loss):
        """UADA_ddp.py:232-325: every rank runs ``val_batches`` forward-only batches of its validation shard, the per-batch
        MSE distance / UAD / CE are averaged over BATCHES (not samples, unlike UADA.py), averaged over ranks, and rank 0 keeps
        the patch with the lowest distance (+ the images of its last validation batch) and refreshes ``last``."""
        import torch.distributed as dist
        h = self.host
        s = torch.zeros(3, dtype=torch.float64)
        nb = 0
        for j, data in enumerate(val_dataloader):
            if j == self.val_batches:
                break
            data = dict(data)
            data["labels"] = self.mask_labels(data["labels"].clone(), self.maskidx)
            sc, _ = h.evaluate(data, fe_mode, loss)
            s += torch.tensor([sc[_lib.S_LOSS].item(), sc[_lib.S_UAD].item(), sc[_lib.S_CE].item()], dtype=torch.float64)
            nb += 1
        avg = (s / max(nb, 1)).to(torch.float32).to(h.device)
        if world_size > 1:
            dist.all_reduce(avg, op=dist.ReduceOp.SUM)
            avg /= world_size
        mse, uad, ce = (float(v) for v in avg.cpu())
        if rank == 0:
            if self.save_dir:
                if mse < self.MSE_Distance_best:
                    self.MSE_Distance_best = mse
                    self._dump_val_images(self._save_patch(h.patch, str(i)))
                self._save_patch(h.patch, "last", outer_iter=i, sched_step=i + 1)
            self.val_CE_loss.append(ce)
            self.val_MSE_Distance.append(mse)
            self.val_UAD.append(uad)
            self.val_logs = getattr(self, "val_logs", []) + [{"VAL_MSE_Distance": mse, "VAL_UAD": uad, "step": i}]
