#!/usr/bin/env python
"""bench.py -- attack-iterations/sec of the UADA inner loop on OpenVLA-7B shapes (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # engine arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the reference's own PyTorch path on host cores

Workload (config[1] of BASELINE.json): UADA, bridge_orig-shaped synthetic batch, patch 3x50x50, per-GPU bs=8,
geometry=True, random-init OpenVLA-7B shapes.  One "step" = one attack iteration = one pass of the inner-loop body
(UADA.py:134-158): front end -> model forward -> loss -> input-gradient backward -> [patch-grad all-reduce] -> AdamW
update -> clamp = ONE call of vla_attack_step = one CUDA graph launch.  `value` is measured with inputs resident in HBM;
`e2e` goes through the reference-facing plugin class (OpenVLAAttacker.patchattack_unconstrained / .attack) fed by a
collator-style loader of PIL images with innerLoop=1, so that EVERY step uploads its batch from host memory and reads its
scalars and predictions back inside the timed region; `e2e_innerloop50` is the same call at the reference's own setting
(one upload per 50 steps).  Extra keys: `strong` (global batch 64 split over the ranks, config #5), `per_rank` (step time
and all-reduce wait per rank), `ref_gpu_path` (the oracle in the reference's eager style on this GPU; N=1 only),
`predict_action` (greedy 7-token action decode with the KV cache at batch 1, against the HBM roofline; N=1 only).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attack_iters_per_sec"
UNIT = "attack-iterations/s (one iteration = fwd + input-grad + update on a per-GPU batch)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (bs)")
    ap.add_argument("--patch", type=int, default=50)
    ap.add_argument("--text-len", type=int, default=33)
    ap.add_argument("--model", default="openvla-7b", choices=["openvla-7b", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--watchdog", type=int, default=900, help="seconds before a hung run dumps stacks and exits")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: bracket ONE steady-state step with cudaProfilerStart/Stop and exit (no JSON line)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="budget of the cpu_baseline sample of the engine arm")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of measured CPU iterations in --impl reference")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling / reference-GPU-path / innerLoop=50 legs")
    ap.add_argument("--strong-global-batch", type=int, default=64)
    return ap.parse_args()


def get_cfg(name):
    from roboticattack_b200.config import openvla_7b, tiny
    return openvla_7b() if name == "openvla-7b" else tiny(img=56)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU reference
def fast_reference_state_dict(cfg, seed=0):
    """Random-init weights with the reference's init rule (modeling_prismatic.py:185-205: N(0, 0.02), zero biases, unit norm
    weights, LayerScale 1e-5) for the CPU timing leg.  torch.randn makes ~0.1 G values/s on the host, so the 7.5 G weights
    are cut from one 64 Mi-value normal pool at per-tensor offsets (memcpy speed); only the timing depends on them."""
    from roboticattack_b200.weights import param_shapes
    g = torch.Generator().manual_seed(seed)
    pool = (torch.randn(1 << 26, generator=g) * 0.02).to(torch.bfloat16)
    sd, off = {}, 0
    for name, shape in param_shapes(cfg, all_blocks=True).items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = "norm" in name.rsplit(".", 2)[-2] if name.count(".") >= 2 else False
        n = int(np.prod(shape))
        if "scale_factor" in name:
            t = torch.full(shape, 1e-5, dtype=torch.bfloat16)
        elif is_norm and leaf == "weight":
            t = torch.ones(shape, dtype=torch.bfloat16)
        elif leaf == "bias":
            t = torch.zeros(shape, dtype=torch.bfloat16)
        else:
            t = torch.empty(n, dtype=torch.bfloat16)
            done = 0
            while done < n:
                off = (off * 1103515245 + 12345 + n) % (pool.numel() // 2)
                k = min(n - done, pool.numel() - off)
                t[done:done + k] = pool[off:off + k]
                done += k
            t = t.view(shape)
        sd[name] = t
    return sd


class CpuReference:
    """The reference's own path (oracle restatement: eager PyTorch, bf16 weights, per-image front-end loop, ALL blocks of both
    towers as timm executes them, full-sequence lm_head + fp32 logits, shifted CE, weighted_loss, autograd backward to the patch
    with frozen weights (UADA_ddp.py:50-51), HF-AdamW + clamp) on the host cores.  One sample = ONE attack iteration at full
    depth and full widths with micro-batch 1 -- measured, not extrapolated over depth; the only scaling applied is the
    stated per-GPU batch factor (CPU time is linear in the batch to first order; a bs-8 iteration would take minutes)."""

    def __init__(self, cfg, patch_hw, text_len, threads):
        from oracle import losses as ol, optim as oo
        from roboticattack_b200.synthetic import synthetic_batch
        torch.set_num_threads(threads)
        self.cfg, self.p, self.threads = cfg, patch_hw, threads
        t0 = time.perf_counter()
        self.sd = fast_reference_state_dict(cfg)
        self.setup_s = time.perf_counter() - t0
        self.b = synthetic_batch(cfg, 1, text_len, seed=1234)
        self.labels = ol.mask_labels_uada(self.b["labels"].clone(), [0])     # scripts/run_UADA.sh: --maskidx 0
        torch.manual_seed(42)
        self.patch = torch.rand(3, patch_hw, patch_hw)
        self.opt = oo.HFAdamW(self.patch.shape, 2e-3)
        random.seed(42)
        np.random.seed(42)

    def step(self):
        from oracle import frontend as ofe, losses as ol, model as om
        from roboticattack_b200.config import NORM_MEAN, NORM_STD
        from roboticattack_b200.synthetic import draw_placements
        c = self.cfg
        t0 = time.perf_counter()
        xy, th = draw_placements(1, (c.img, c.img), (self.p, self.p), True)
        p = self.patch.clone().requires_grad_(True)
        px = ofe.apply_patch_batch(self.b["obs"], p, xy[0], th[0], ofe.MODE_WARP, NORM_MEAN, NORM_STD)
        out = om.forward(self.sd, c, self.b["input_ids"], self.b["attention_mask"], px.to(torch.bfloat16), self.labels,
                         run_unused_last_block=True)
        mse, _ = ol.weighted_loss_uada(out.logits, self.labels, 5)
        (mse + 1 / out.loss).backward()
        self.opt.step(self.patch, p.grad)
        self.patch.clamp_(0, 1)
        return time.perf_counter() - t0


def cpu_reference_rate(cfg, batch, patch_hw, text_len, threads, steps=1, warmup=0, budget_s=30.0):
    """-> (it/s at the per-GPU batch, description, samples taken, seconds per sample)."""
    ref = CpuReference(cfg, patch_hw, text_len, threads)
    t_start = time.perf_counter()
    ts = []
    for i in range(warmup + steps):
        dt = ref.step()
        if i >= warmup:
            ts.append(dt)
        if time.perf_counter() - t_start > budget_s and ts:
            break
    t = float(np.median(ts))
    sample = (f"oracle (PyTorch restatement of the reference path, frozen weights) on {threads} host threads: {len(ts)} measured attack "
              f"iteration(s) at FULL depth ({cfg.llm.layers} Llama layers, {cfg.dino.depth}+{cfg.siglip.depth} ViT blocks) and full widths "
              f"with micro-batch 1, median {t:.2f} s (all: {[round(x, 2) for x in ts]}), x per-GPU batch {batch} (linear in the batch); "
              f"weights from a 64 Mi-value normal pool in {ref.setup_s:.0f} s")
    return 1.0 / (t * batch), sample, len(ts), t


def reference_arm(args):
    """A step = the reference's CPU path over ONE per-GPU batch, executed as `batch` micro-batch-1 attack iterations back to
    back (the same arithmetic as one bs-`batch` iteration at 1/`batch` of the activation memory; a single bs-8 iteration
    needs ~100 GB of fp32/bf16 autograd state on the host).  Every step is measured at full depth; the number of steps is
    what fits into --ref-budget seconds and is reported as `steps`."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cfg = get_cfg(args.model)
    threads = os.cpu_count() or 1
    ref = CpuReference(cfg, args.patch, args.text_len, threads)
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        ref.step()                                  # one micro-iteration: allocator / thread-pool warm-up
    t_start = time.perf_counter()
    step_s = []
    while len(step_s) < max(1, args.steps):
        step_s.append(sum(ref.step() for _ in range(args.batch)))
        if time.perf_counter() - t_start > args.ref_budget:
            break
    t = float(np.median(step_s))
    value = 1.0 / t
    sample = (f"oracle (PyTorch restatement of the reference path, frozen weights) on {threads} host threads: {len(step_s)} measured step(s), "
              f"each = {args.batch} micro-batch-1 attack iterations at FULL depth ({cfg.llm.layers} Llama layers, {cfg.dino.depth}+{cfg.siglip.depth} "
              f"ViT blocks) and full widths = one per-GPU batch of {args.batch}; median {t:.1f} s per step (all: {[round(x, 1) for x in step_s]}); "
              f"weights from a 64 Mi-value normal pool in {ref.setup_s:.0f} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(step_s),
            "warmup": warm, "ms_per_step": 1000.0 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args),
                       "note": "reference arm = the reference's CPU PyTorch path on this box's host cores, measured at full depth; "
                               f"steps = the steps that fit into the {args.ref_budget:.0f} s budget (asked: {args.steps}); warm-up = {warm} micro-iteration"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def workload_name(args):
    return (f"UADA bridge_orig-shaped synthetic, patch 3x{args.patch}x{args.patch}, per-GPU bs={args.batch}, T={args.text_len} "
            f"(L={args.text_len + 256 if args.model == 'openvla-7b' else 'tiny'}), geometry=True, maskidx 0, {args.model} random-init")


# ----------------------------------------------------------------------------------------------------- engine arm
class TimedPILLoader:
    """Collator-style loader (prismatic/util/data_utils.py:183-217): dicts with `pixel_values` = list of PIL images and
    right-padded id / mask / label tensors.  A CUDA event + wall time is recorded at every fetch; the attack loops fetch batch
    k + 1 right after launching the steps of batch k, so event[a] -> event[a + n] spans exactly n outer iterations."""

    def __init__(self, cfg, B, T, n_distinct, seed):
        from PIL import Image
        from roboticattack_b200.synthetic import synthetic_batch
        self.batches = []
        for i in range(n_distinct):
            b = synthetic_batch(cfg, B, T, seed=seed + i)
            self.batches.append({"pixel_values": [Image.fromarray(im.numpy()) for im in b["obs"]], "input_ids": b["input_ids"],
                                 "attention_mask": b["attention_mask"], "labels": b["labels"]})
        self.events, self.walls = [], []

    def __iter__(self):
        while True:
            for b in self.batches:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.events.append(ev)
                self.walls.append(time.perf_counter())
                yield b


def plugin_e2e(cfg, eng, args, world, rank, dev, inner, outer, warm_outer):
    """attack-iterations/s through the reference-facing class: UADAAttacker.patchattack_unconstrained (UADA.py:93-292) on one
    GPU, UADADDPAttacker.attack (UADA_ddp.py:138-232) under torchrun.  Returns (ms per attack iteration, h2d, d2h bytes per
    iteration)."""
    import torch.distributed as dist
    from roboticattack_b200.attacker import UADADDPAttacker, UADAAttacker
    B, T, p = args.batch, args.text_len, args.patch
    loader = TimedPILLoader(cfg, B, T, 4, seed=4321 + 100 * rank)
    n_outer = warm_outer + outer + 1            # the event of fetch a + n is recorded while iteration a + n - 1 runs
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42)
    ns = argparse.Namespace(wandb_project="false")
    if world == 1:
        a = UADAAttacker(eng, None, save_dir="", optimizer="adamW", cfg=cfg, device=dev)
        a.val_every = 10 ** 9
        a.patchattack_unconstrained(loader, None, num_iter=n_outer, patch_size=[3, p, p], lr=2e-3, maskidx=[0], warmup=0,
                                    geometry=True, innerLoop=inner, args=ns)
    else:
        a = UADADDPAttacker(eng, None, save_dir="", patch_size=[3, p, p], lr=2e-3, bs=B, warmup=0, num_iter=n_outer, maskidx=[0],
                            innerLoop=inner, geometry=True, use_wandb=False, MSE_weights=5, cfg=cfg, device=dev)
        a.val_every = 10 ** 9
        a.attack(rank, world, train_dataloader=loader)
    torch.cuda.synchronize()
    e0, e1 = loader.events[warm_outer + 1], loader.events[warm_outer + 1 + outer]
    ms = e0.elapsed_time(e1)
    wall_ms = (loader.walls[warm_outer + 1 + outer] - loader.walls[warm_outer + 1]) * 1e3
    t = torch.tensor([max(ms, wall_ms)], device=dev, dtype=torch.float64)    # device time and host clock agree in steady state
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    R = eng.num_supervised
    h2d = (B * cfg.img * cfg.img * 3 + B * T * 8 + B * 4 + R * 16) / inner + B * 2 * 4 + B * 6 * 4 + 8 / inner
    d2h = _SC * 4 + R * 4 / inner
    return t.item() / (outer * inner), int(h2d), int(d2h)


_SC = 8


def ref_gpu_path(cfg, args, dev, iters=5):
    """The reference-STYLE GPU path on this box (last row): the oracle restatement of the reference's eager
    PyTorch attack iteration (per-image front-end loop, autograd, all ViT blocks, full-sequence lm_head + fp32 logits,
    weighted_loss, HF-AdamW, clamp, one .item()) in bf16 with torch's own kernels, weights frozen (UADA_ddp.py:50-51) and
    requiring grad (UADA.py never freezes them).  Baseline measurement only; outside every timed region of the engine."""
    from oracle import frontend as ofe, losses as ol, model as om, optim as oo
    from roboticattack_b200.config import NORM_MEAN, NORM_STD
    from roboticattack_b200.synthetic import draw_placements, synthetic_batch
    from roboticattack_b200.weights import param_shapes, random_tensor
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {n: random_tensor(n, sh, g, dev, "reference").to(torch.bfloat16) for n, sh in param_shapes(cfg, all_blocks=True).items()}
    b = synthetic_batch(cfg, args.batch, args.text_len, seed=1234)
    labels = ol.mask_labels_uada(b["labels"].clone(), [0]).to(dev)
    out = {}
    with torch.device(dev):
        ids, mask, obs = b["input_ids"].to(dev), b["attention_mask"].to(dev), b["obs"].to(dev)
        for frozen in (True, False):
            for t in sd.values():
                t.requires_grad_(not frozen)
            torch.manual_seed(42)
            patch = torch.rand(3, args.patch, args.patch, device=dev)
            opt = oo.HFAdamW(patch.shape, 2e-3)
            random.seed(42)
            np.random.seed(42)
            ts = []
            for _ in range(iters):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                xy, th = draw_placements(args.batch, (cfg.img, cfg.img), (args.patch, args.patch), True)
                p = patch.clone().requires_grad_(True)
                px = ofe.apply_patch_batch(obs, p, xy[0], th[0], ofe.MODE_WARP, NORM_MEAN, NORM_STD)
                o = om.forward(sd, cfg, ids, mask, px.to(torch.bfloat16), labels, run_unused_last_block=True)
                mse, _ = ol.weighted_loss_uada(o.logits, labels, 5)
                loss = mse + 1 / o.loss
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
            out["frozen_weights_UADA_ddp" if frozen else "weights_require_grad_UADA"] = {"ms_per_iter": round(best * 1e3, 2), "iters_per_sec": round(1 / best, 3)}
    del sd
    torch.cuda.empty_cache()
    out["what"] = ("oracle restatement of the reference's eager PyTorch iteration on this GPU (cuBLAS / SDPA / ATen, bf16, bs "
                   f"{args.batch}), best of {iters - 2} after 2 warm-up iterations; measured by bench.py outside the engine's timed regions")
    return out


def predict_action_leg(cfg, eng, text_len, n_tokens=7, iters=5):
    """The consumer of ``patch.pt``: ``predict_action`` (modeling_prismatic.py:506-536) at batch 1 on the engine -- one prefill over
    image + prompt, then ``n_tokens - 1`` single-position steps on the KV cache (five fused HBM-bound kernels per layer, one
    recorded step replayed per token).  Wall clock around synchronised calls; the decode streams every Llama weight once per
    token, so the per-token figure is reported against the measured HBM bandwidth.  Outside every timed region of the attack step."""
    from roboticattack_b200.policy import ActionPolicy
    from roboticattack_b200.synthetic import synthetic_batch
    b = synthetic_batch(cfg, 1, text_len, seed=3)
    prompt = b["input_ids"][:, :text_len - n_tokens - 1]
    pol = ActionPolicy(eng)

    def timed(n):
        for _ in range(2):
            pol.generate_action_tokens(b["obs"], prompt, n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            pol.generate_action_tokens(b["obs"], prompt, n)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3
    t_all, t_first = timed(n_tokens), timed(1)
    per_tok = (t_all - t_first) / (n_tokens - 1)
    h, f, V, nl = cfg.llm.hidden, cfg.llm.ffn, cfg.llm.vocab, cfg.llm.layers
    wbytes = 2 * (nl * (3 * h * h + h * h + 2 * f * h + h * f) + V * h)     # bf16 weights a decode step streams once
    pk = peaks()
    return {"ms_per_action": round(t_all, 3), "actions_per_sec": round(1e3 / t_all, 2), "prefill_plus_first_token_ms": round(t_first, 3),
            "ms_per_token": round(per_tok, 4), "weight_bytes_per_token": wbytes,
            "roofline": {"bound": "hbm", "achieved": round(wbytes / per_tok / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                         "frac": round(wbytes / per_tok / 1e6 / pk["hbm"], 4), "peak_source": pk["src"]},
            "note": f"ActionPolicy.generate_action_tokens, batch 1, prompt of {prompt.shape[1]} tokens, {n_tokens} greedy tokens, "
                    f"{iters} calls after 2 warm-up calls; per token = (action - (prefill + first token)) / {n_tokens - 1}"}


def engine_arm(args):
    import ctypes
    import torch.distributed as dist
    from roboticattack_b200 import _lib
    from roboticattack_b200 import labels as lab
    from roboticattack_b200.config import flops_per_sample
    from roboticattack_b200.engine import LossSpec, VLAEngine
    from roboticattack_b200.synthetic import draw_placements, synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py engine arm needs a CUDA device; there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = get_cfg(args.model)
    B, T, p = args.batch, args.text_len, args.patch
    K, W = args.steps, max(args.warmup, 3)
    lib = _lib.lib()
    MASKIDX = [0]                                                   # scripts/run_UADA.sh: --maskidx 0

    eng = VLAEngine(cfg, B, T, device=dev)
    eng.load_random_weights(seed=0, init="reference")
    comm = eng.make_comm(rank, world)                               # vla_comm (NCCL) for the patch-gradient all-reduce
    # the SAME loss at every N (the DDP variant's, UADA_ddp.py:203-206): identical work per GPU in the scaling run
    loss = LossSpec(_lib.LOSS_UADA_DDP, mse_weight=5.0)
    lr = 2e-3

    def make_state(b_, seed, n_place):
        batch = synthetic_batch(cfg, b_, T, seed=seed)              # each rank its own shard of the global batch
        labels = lab.mask_labels_uada(batch["labels"].clone(), MASKIDX)
        random.seed(42)
        np.random.seed(42)                                         # identical placement stream on every rank (UADA_wrapper_ddp.py:53)
        xy, theta = draw_placements(b_, (cfg.img, cfg.img), (p, p), True, steps=n_place)
        torch.manual_seed(42)
        patch = torch.rand(3, p, p).to(dev)
        if world > 1:
            dist.broadcast(patch, src=0)
        st = {"patch": patch, "m": torch.zeros_like(patch), "v": torch.zeros_like(patch), "grad": torch.zeros_like(patch),
              "hist": torch.zeros(n_place, _lib.NUM_SCALARS, device=dev), "xy": xy, "theta": theta, "batch": batch, "labels": labels}
        eng.ensure_plan(b_, T)
        eng.set_batch(batch["obs"].pin_memory(), batch["input_ids"], batch["attention_mask"], labels)
        eng.set_placements(xy, theta)
        st["pred"] = torch.full((eng.num_supervised,), -1, dtype=torch.int32, device=dev)
        eng.set_step_state(0, 0)
        return st

    def step(st, graph=True):
        eng.attack_step(st["patch"], st["m"], st["v"], st["grad"], st["hist"], st["pred"], _lib.FE_WARP, loss, lr, comm=comm, graph=graph)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def gather(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world == 1:
            return [float(x)]
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.item() for o in outs]

    # ---- value: K attack iterations (one graph launch each), inputs resident in HBM --------------------------------------
    st = make_state(B, 1234 + rank, W + K)
    for s in range(W):
        step(st)                                                    # eager (autotune), recording, replays
    sync()
    if args.ncu_step:
        torch.cuda.profiler.start()
        step(st, graph=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    sampler = ClockSampler(local)             # every rank samples its own board (the slowest one paces a data-parallel step)
    launches0, replays0 = lib.vla_launch_count(), lib.vla_graph_replays()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for s in range(K):
        step(st)
    e1.record()
    sync()
    wall1 = time.time()
    ms_rank = e0.elapsed_time(e1)
    launches, replays = lib.vla_launch_count() - launches0, lib.vla_graph_replays() - replays0
    clocks = sampler.stop(wall0, wall1)
    ms_max = max_over_ranks(ms_rank)
    per_rank_ms = gather(ms_rank / K)
    per_rank_mhz = gather(clocks["sm_mhz"] or 0.0)
    losses = st["hist"][W:W + K, _lib.S_LOSS].cpu()
    assert torch.isfinite(losses).all(), "non-finite loss in the timed region"
    # all ranks must hold bit-identical patches (same reduced gradient, same replicated update; UADA_ddp.py:140-166,206)
    patches_identical = None
    if world > 1:
        allp = [torch.zeros_like(st["patch"]) for _ in range(world)]
        dist.all_gather(allp, st["patch"])
        patches_identical = all(torch.equal(allp[0], q) for q in allp[1:])
        assert patches_identical, "patches diverged across ranks"

    # ---- roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of eager steps; the same
    # steps time the all-reduce on every rank (wait for the slowest rank + the collective itself) ------------------------
    roof = None
    eng.set_single_stream(True)               # per-kernel event timing needs the kernels of a stream back to back
    NPROF = 3                                 # three steps: one step's sum of ~660 event pairs moves by +-5 % from run to run
    st_p = make_state(B, 1234 + rank, NPROF + 4)
    step(st_p, graph=False)
    sync()
    if rank == 0:
        lib.vla_profile_gemm_begin()
    def composed_steps(first, n, timed):      # composed from the ABI's parts so that the collective can be bracketed by events
        evs = []
        for i in range(first, first + n):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.fwd_bwd(st_p["patch"], i, _lib.FE_WARP, loss, st_p["grad"], st_p["hist"][i], st_p["pred"])
            a0.record()
            if comm is not None:
                comm.all_reduce_(st_p["grad"])
            a1.record()
            eng.patch_update(st_p["patch"], st_p["grad"], st_p["m"], st_p["v"], 1 + i, lr, grad_scale=1.0 / world, scalars=st_p["hist"][i])
            evs.append((a0, a1))
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b_) for a, b_ in evs])) if timed else 0.0
    composed_steps(1, NPROF, False)           # rank 0: per-GEMM events (these steps are slower on rank 0, so they are not the ones
    eng.set_single_stream(False)              # whose all-reduce wait is reported)
    if rank == 0:
        tm, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        _lib.check(lib.vla_profile_gemm_end(ctypes.byref(tm), ctypes.byref(fl), ctypes.byref(n)), "profile end")
        pk = peaks()
        ach = fl.value / (tm.value * 1e-3) / 1e12
        f = flops_per_sample(cfg, T, supervised_rows=len(MASKIDX) + 1)
        from roboticattack_b200.config import gemm_algorithmic_bytes
        traffic = None   # DRAM bytes of the GEMM launches of one step, from the committed ncu launch list of this command
        tpath = os.path.join(ROOT, "profiles", "r02_gemm_dram_traffic.json")
        if os.path.exists(tpath) and world == 1:
            with open(tpath) as fh:
                traffic = json.load(fh)["dram_bytes_per_step"]
        alg_bytes = gemm_algorithmic_bytes(cfg, B, T, len(MASKIDX) + 1)
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05, all launches of one step)", "achieved": round(ach, 1),
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["bf16_sustained"], 4), "traffic": traffic,
                "traffic_note": "dram__bytes_read+write summed over the GEMM launches of one step (ncu, caches flushed between launches; "
                                "profiles/r02_gemm_dram_traffic.json); algorithmic_bytes = every GEMM operand and result once, "
                                "summed over the step's GEMM shape table (roboticattack_b200.config.gemm_algorithmic_bytes)",
                "algorithmic_bytes": alg_bytes,
                "peak_source": pk["src"] + ", sustained figure (kernel timed inside a long step)",
                "gemm_launches_per_step": n.value // NPROF, "gemm_ms_per_step": round(tm.value / NPROF, 3),
                "gemm_flops_per_step": fl.value / NPROF, "profiled_steps": NPROF, "algorithmic_flops_per_step": f["iter"] * B,
                "executed_flops_per_step": f["iter_executed"] * B,   # last decoder layer pruned to the supervised rows (exact)
                "step_tflops": round(f["iter_executed"] * B / (ms_max / K * 1e-3) / 1e12, 1)}
    sync()
    ar_all = gather(composed_steps(1 + NPROF, 3, True))   # un-profiled eager steps on every rank: wait for the slowest rank + collective
    sync()
    # the same graph-replayed step WITHOUT the collective: each rank's own pace (attributes the weak-scaling loss: the step of a
    # data-parallel run is the slowest rank's compute + the collective, and the boards run power-capped at different clocks)
    uncoupled = None
    if world > 1:
        keep, comm = comm, None
        st_u = make_state(B, 1234 + rank, 14)
        for _ in range(4):
            step(st_u)
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(10):
            step(st_u)
        u1.record()
        torch.cuda.synchronize()
        comm = keep
        uncoupled = gather(u0.elapsed_time(u1) / 10)
        sync()

    # ---- e2e: the reference-facing plugin class with a PIL loader; every step uploads its batch and reads its results --------
    Ke = min(K, 20)
    e2e_ms, h2d, d2h = plugin_e2e(cfg, eng, args, world, rank, dev, inner=1, outer=Ke, warm_outer=3)
    e2e50 = None
    if not args.no_extras:
        ms50, h50, d50 = plugin_e2e(cfg, eng, args, world, rank, dev, inner=50, outer=1, warm_outer=1)
        e2e50 = {"value": world / (ms50 * 1e-3), "unit": UNIT, "ms_per_step": ms50, "h2d_bytes_per_step": h50, "d2h_bytes_per_step": d50,
                 "note": "same call with the reference's innerLoop=50 (scripts/run_UADA.sh): one batch upload per 50 attack iterations"}

    # ---- strong scaling of config #5: global batch 64 split over the ranks; this box's own 1-GPU time next to it ---------------
    strong = None
    GB = args.strong_global_batch
    if not args.no_extras and args.model == "openvla-7b" and GB % world == 0:
        def time_bs(b_, use_comm, n=3):
            nonlocal comm
            keep, comm = comm, (comm if use_comm else None)
            s_ = make_state(b_, 777 + (rank if use_comm else 0), n + 2)
            step(s_)
            step(s_)
            sync()
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record()
            for _ in range(n):
                step(s_)
            t1_.record()
            sync()
            comm = keep
            return max_over_ranks(t0_.elapsed_time(t1_) / n)
        t_n = time_bs(GB // world, True)
        t_1 = time_bs(GB, False) if world > 1 else t_n      # every rank times the whole global batch alone (no collective)
        strong = {"global_batch": GB, "per_gpu_batch": GB // world, "ms_per_step": round(t_n, 3),
                  "single_gpu_ms_per_step": round(t_1, 3), "speedup_vs_1gpu": round(t_1 / t_n, 3), "iters_per_sec": round(1e3 / t_n, 3),
                  "note": "UADA_ddp global bs 64 (BASELINE config #5): the same global batch on N GPUs vs on one GPU of this box "
                          "(max over ranks, 3 steps after 2 warm-up steps, CUDA events)"}

    decode = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            decode = predict_action_leg(cfg, eng, T)
        except Exception as ex:      # noqa: BLE001 -- an extra leg must never cost the headline line
            decode = {"unavailable": f"{type(ex).__name__}: {str(ex)[:120]}"}

    refgpu = None
    if rank == 0 and world == 1 and not args.no_extras and args.model == "openvla-7b":
        eng.shrink_plan(1, T)                 # hand the activation arena (82 GB after the bs-64 leg) back before the eager path allocates
        try:
            refgpu = ref_gpu_path(cfg, args, dev)
        except torch.cuda.OutOfMemoryError as ex:     # noqa: PERF203
            refgpu = {"unavailable": f"out of memory next to the engine's arenas: {str(ex)[:80]}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v_cpu, sample, _, _ = cpu_reference_rate(cfg, B, p, T, os.cpu_count() or 1, steps=2, warmup=0, budget_s=args.cpu_seconds)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    if rank == 0:
        value = world * K / (ms_max * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": workload_name(args), "per_gpu_batch": B, "global_batch": B * world, "loss": "UADA_ddp (mean((w e - w t)^2), MSE_weights 5)",
                           "maskidx": MASKIDX, "cache": "inputs larger than L2: 30 GB of weights streamed per step",
                           "parallelism": f"dp{world} (batch sharded over ranks, NCCL all-reduce of the patch gradient inside the step's CUDA graph)" if world > 1 else "single GPU"},
                "e2e": {"value": world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms,
                        "path": ("UADAAttacker.patchattack_unconstrained" if world == 1 else "UADADDPAttacker.attack") +
                                " with a PIL-list loader, innerLoop=1: every attack iteration uploads its batch (PIL -> pinned uint8 -> HBM), "
                                "its placements and label tables, and reads back its scalar record and predicted ids"},
                "e2e_innerloop50": e2e50,
                "gpu_launches": int(launches), "launches_per_step": launches / K, "graph_launches_per_step": replays / K,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "ref_gpu_path": refgpu, "strong": strong, "predict_action": decode,
                "per_rank": {"ms_per_step": [round(x, 3) for x in per_rank_ms], "sm_mhz": [round(x) for x in per_rank_mhz],
                             "uncoupled_ms_per_step": None if uncoupled is None else [round(x, 3) for x in uncoupled],
                             "allreduce_wait_ms_eager": [round(x, 4) for x in ar_all],
                             "patches_bit_identical": patches_identical,
                             "note": "ms_per_step: each rank's own CUDA-event time of the timed region / K (coupled by the all-reduce inside the "
                                     "graph); sm_mhz: each board's median SM clock in that region; uncoupled_ms_per_step: the same graph-replayed "
                                     "step without the collective, every rank at its own pace (10 steps) -- the coupled step tracks the slowest "
                                     "rank; allreduce_wait_ms_eager: events around vla_allreduce_patch_grad in eager, Python-driven steps = wait "
                                     "for the slowest rank + the 30 KB collective"},
                "loss_first_last": [losses[0].item(), losses[-1].item()]}
        emit(line)
    if world > 1:
        # the vla_comm is left to process exit: ncclCommDestroy is a collective that also waits for the recorded graphs, and a
        # benchmark has nothing to gain from an orderly teardown after its line is printed
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was sent to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # fd 1 -> stderr for the whole run (NCCL prints its version banner on stdout)
    args = parse()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)   # a hung collective dumps stacks and exits instead of burning the box
    if args.impl == "reference":
        return reference_arm(args)
    return engine_arm(args)


if __name__ == "__main__":
    sys.exit(main())
