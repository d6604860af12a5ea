#!/usr/bin/env python
"""bench.py -- attack-iterations/sec of the UADA inner loop on OpenVLA-7B shapes (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # engine arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the reference's own PyTorch path on host cores

Workload (config[1] of BASELINE.json): UADA, bridge_orig-shaped synthetic batch, patch 3x50x50, per-GPU bs=8,
geometry=True, random-init OpenVLA-7B shapes.  One "step" = one attack iteration = one pass of the inner-loop body
(UADA.py:134-158): front end -> model forward -> loss -> input-gradient backward -> [patch-grad all-reduce] -> AdamW
update -> clamp.  `value` is measured with inputs resident in HBM; `e2e` goes through the public API with host
buffers (per-step H2D of the step's inputs from pinned memory, D2H of the step's loss) inside the timed region.
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
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline sample")
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
def cpu_reference_rate(cfg, batch, patch_hw, text_len, budget_s, threads):
    """The reference's own path (oracle restatement: eager PyTorch, bf16 weights, per-image front-end loop, full-sequence
    lm_head + fp32 logits, shifted CE, weighted_loss, autograd backward to the patch, HF-AdamW + clamp) timed on the host
    cores on a BOUNDED sample: micro-batch 1, full widths, reduced depth (dl Llama layers, dv blocks per tower), timed at
    two depths so that the per-layer cost is separated from the depth-independent cost, then extrapolated linearly to the
    full depth (all blocks of both towers, as the reference executes them) and to the per-GPU batch."""
    import dataclasses
    from oracle import frontend as ofe, losses as ol, model as om, optim as oo
    from roboticattack_b200.config import NORM_MEAN, NORM_STD, OpenVLAConfig
    from roboticattack_b200.synthetic import draw_placements, synthetic_batch
    from roboticattack_b200.weights import random_state_dict
    torch.set_num_threads(threads)

    def run(dl, dv, iters):
        c = OpenVLAConfig(dino=dataclasses.replace(cfg.dino, depth=dv + 1), siglip=dataclasses.replace(cfg.siglip, depth=dv + 1),
                          llm=dataclasses.replace(cfg.llm, layers=dl), name="slice")
        sd = random_state_dict(c, seed=0, dtype=torch.bfloat16, init="reference")
        b = synthetic_batch(c, 1, text_len, seed=1234)
        labels = ol.mask_labels_uada(b["labels"].clone(), [0, 1, 2])
        torch.manual_seed(42)
        patch = torch.rand(3, patch_hw, patch_hw)
        opt = oo.HFAdamW(patch.shape, 2e-3)
        random.seed(42)
        np.random.seed(42)
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            xy, th = draw_placements(1, (c.img, c.img), (patch_hw, patch_hw), True)
            p = patch.clone().requires_grad_(True)
            px = ofe.apply_patch_batch(b["obs"], p, xy[0], th[0], ofe.MODE_WARP, NORM_MEAN, NORM_STD)
            out = om.forward(sd, c, b["input_ids"], b["attention_mask"], px.to(torch.bfloat16), labels)
            mse, _ = ol.weighted_loss_uada(out.logits, labels, 5)
            (mse + 1 / out.loss).backward()
            opt.step(patch, p.grad)
            patch.clamp_(0, 1)
            ts.append(time.perf_counter() - t0)
        return min(ts)

    t_start = time.perf_counter()
    t1 = run(1, 1, 3)
    t2 = run(3, 3, 3)
    per_layer_all = max((t2 - t1) / 2.0, 1e-6)   # one Llama layer + one block of each tower
    base = max(t1 - per_layer_all, 0.0)          # front end, patch embed, projector, lm_head, loss, update
    # split the per-depth cost between LLM and towers by their FLOP shares
    from roboticattack_b200.config import flops_per_sample
    L = text_len + cfg.num_patches
    f_llm = 2 * L * (4 * cfg.llm.hidden ** 2 + 3 * cfg.llm.hidden * cfg.llm.ffn)
    f_d = 2 * cfg.dino.tokens * (4 * cfg.dino.dim ** 2 + 2 * cfg.dino.dim * cfg.dino.mlp_hidden)
    f_s = 2 * cfg.siglip.tokens * (4 * cfg.siglip.dim ** 2 + 2 * cfg.siglip.dim * cfg.siglip.mlp_hidden)
    tot = f_llm + f_d + f_s
    t_full = base + per_layer_all * (cfg.llm.layers * f_llm + cfg.dino.depth * f_d + cfg.siglip.depth * f_s) / tot
    t_iter = t_full * batch
    sample = (f"oracle (PyTorch restatement of the reference path, frozen weights), bs=1, full widths, depth 1 and 3 of "
              f"{cfg.llm.layers} Llama layers / {cfg.dino.depth}+{cfg.siglip.depth} ViT blocks, fwd + backward-to-patch + AdamW, "
              f"best of 3 each ({t1:.2f}s, {t2:.2f}s); linear extrapolation to full depth ({t_full:.1f}s/sample) x bs={batch}; "
              f"sample took {time.perf_counter() - t_start:.0f}s of CPU work")
    return 1.0 / t_iter, sample


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cfg = get_cfg(args.model)
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 2))):          # each "step" is one bounded sample; keep the run to minutes
        v, sample = cpu_reference_rate(cfg, args.batch, args.patch, args.text_len, args.cpu_seconds, threads)
        vals.append(v)
        if time.perf_counter() - t0 > 120:
            break
    value = float(np.median(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "reference arm = the reference's CPU PyTorch path on this box's host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def workload_name(args):
    return (f"UADA bridge_orig-shaped synthetic, patch 3x{args.patch}x{args.patch}, per-GPU bs={args.batch}, T={args.text_len} "
            f"(L={args.text_len + 256 if args.model == 'openvla-7b' else 'tiny'}), geometry=True, {args.model} random-init")


# ----------------------------------------------------------------------------------------------------- engine arm
def engine_arm(args):
    import torch.distributed as dist
    from roboticattack_b200 import _lib
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

    eng = VLAEngine(cfg, B, T, device=dev)
    eng.load_random_weights(seed=0, init="reference")
    batch = synthetic_batch(cfg, B, T, seed=1234 + rank)            # each rank its own shard of the global batch
    from roboticattack_b200 import labels as lab
    labels = lab.mask_labels_uada(batch["labels"].clone(), [0, 1, 2])   # scripts/run_UADA.sh: --maskidx 0,1,2
    obs_pinned = batch["obs"].pin_memory()
    random.seed(42)
    np.random.seed(42)                                             # identical placement stream on every rank (UADA_wrapper_ddp.py:53)
    n_place = W + K
    xy, theta = draw_placements(B, (cfg.img, cfg.img), (p, p), True, steps=n_place)
    torch.manual_seed(42)
    patch = torch.rand(3, p, p).to(dev)
    if world > 1:
        dist.broadcast(patch, src=0)
    m, v, grad = torch.zeros_like(patch), torch.zeros_like(patch), torch.zeros_like(patch)
    loss = LossSpec(_lib.LOSS_UADA if world == 1 else _lib.LOSS_UADA_DDP, mse_weight=5.0)
    scal = torch.zeros(n_place, _lib.NUM_SCALARS, device=dev)
    eng.set_batch(obs_pinned, batch["input_ids"], batch["attention_mask"], labels)
    eng.set_placements(xy, theta)
    pred = torch.full((eng.num_supervised,), -1, dtype=torch.int32, device=dev)
    lr = 2e-3
    state = {"t": 0}

    def step(s):
        eng.fwd_bwd(patch, s, _lib.FE_WARP, loss, grad, scal[s], pred)
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        state["t"] += 1
        eng.patch_update(patch, grad, m, v, state["t"], lr, grad_scale=1.0 / world, scalars=scal[s])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        step(s)
    sync()
    if args.ncu_step:
        torch.cuda.profiler.start()
        step(W)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.vla_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for s in range(W, W + K):
        step(s)
    e1.record()
    sync()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.vla_launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    losses = scal[W:W + K, _lib.S_LOSS].cpu()
    assert torch.isfinite(losses).all(), "non-finite loss in the timed region"

    # ---- e2e: public API with HOST buffers, H2D of the step's inputs + D2H of the step's loss inside the timed region
    ids_h, mask_h = batch["input_ids"], batch["attention_mask"]

    def e2e_step(s):
        eng.set_batch(obs_pinned, ids_h, mask_h, labels)            # H2D: obs + ids (+ row tables) from pinned / host memory
        eng.set_placements(xy[s:s + 1], theta[s:s + 1])             # H2D: this step's placements
        eng.fwd_bwd(patch, 0, _lib.FE_WARP, loss, grad, scal[s], pred)
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        state["t"] += 1
        eng.patch_update(patch, grad, m, v, state["t"], lr, grad_scale=1.0 / world, scalars=scal[s])
        return scal[s].cpu()                                        # D2H: loss / CE / UAD / grad mean of this step

    Ke = min(K, 10)
    for s in range(2):
        e2e_step(s)
    sync()
    e0.record()
    for s in range(W, W + Ke):
        e2e_step(s)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / Ke
    h2d = obs_pinned.numel() + ids_h.numel() * 8 + B * 4 + eng.num_supervised * 16 + B * 2 * 4 + B * 6 * 4
    d2h = _lib.NUM_SCALARS * 4

    # ---- roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of one more step
    roof = None
    eng.set_placements(xy, theta)
    import ctypes
    eng.set_single_stream(True)               # per-kernel event timing needs the kernels of a stream back to back
    if rank == 0:
        lib.vla_profile_gemm_begin()
    NPROF = 3                                 # three steps: one step's sum of ~660 event pairs moves by +-5 % from run to run
    for _ in range(NPROF):
        step(W)                               # every rank runs the steps (they contain the all-reduce); rank 0 times its GEMMs
    eng.set_single_stream(False)
    if rank == 0:
        tm, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        _lib.check(lib.vla_profile_gemm_end(ctypes.byref(tm), ctypes.byref(fl), ctypes.byref(n)), "profile end")
        pk = peaks()
        traffic = None   # DRAM bytes of the GEMM launches of one step, from the committed ncu launch list of this command
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01o_gemm_dram_traffic.json")
        if os.path.exists(tpath) and world == 1:
            with open(tpath) as fh:
                traffic = json.load(fh)["dram_bytes_per_step"]
        ach = fl.value / (tm.value * 1e-3) / 1e12
        f = flops_per_sample(cfg, T, supervised_rows=4)             # maskidx 0,1,2 + EOS = 4 supervised rows per sample
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05, all launches of one step)", "achieved": round(ach, 1),
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["bf16_sustained"], 4), "traffic": traffic,
                "traffic_note": "dram__bytes_read+write summed over the GEMM launches of one step (ncu, caches flushed between launches; profiles/r01o_gemm_dram_traffic.json); algorithmic: 30 GB of weights + ~12 GB of activations",
                "peak_source": pk["src"] + ", sustained figure (kernel timed inside a long step)",
                "gemm_launches_per_step": n.value // NPROF, "gemm_ms_per_step": round(tm.value / NPROF, 3),
                "gemm_flops_per_step": fl.value / NPROF, "profiled_steps": NPROF, "algorithmic_flops_per_step": f["iter"] * B,
                "executed_flops_per_step": f["iter_executed"] * B,   # last decoder layer pruned to the supervised rows (exact)
                "step_tflops": round(f["iter_executed"] * B / (ms_max / K * 1e-3) / 1e12, 1)}
    sync()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v_cpu, sample = cpu_reference_rate(cfg, B, p, T, args.cpu_seconds, os.cpu_count() or 1)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    if rank == 0:
        value = world * K / (ms_max * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": workload_name(args), "per_gpu_batch": B, "global_batch": B * world, "loss": "UADA" if world == 1 else "UADA_ddp",
                           "cache": "inputs larger than L2: 30 GB of weights streamed per step",
                           "parallelism": f"dp{world} (batch sharded over ranks, patch-grad all-reduce per step)" if world > 1 else "single GPU"},
                "e2e": {"value": world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms},
                "gpu_launches": int(launches), "launches_per_step": launches / K,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "loss_first_last": [losses[0].item(), losses[-1].item()]}
        emit(line)
    if world > 1:
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
