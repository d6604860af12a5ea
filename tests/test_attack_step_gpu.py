"""`vla_attack_step` (include/vla_b200.h): the whole inner-loop body of the reference (UADA.py:134-158, UADA_ddp.py:192-209,
TMA.py:133-175) as one C-ABI call that is recorded into a CUDA graph on its second use.

  * the graph replays, the eager sequence and the round-1 composition (vla_fwd_bwd + vla_patch_update driven from Python)
    produce the same trajectory (first-step scalars bit-identical; later steps to fp32 round-off: the front-end backward
    accumulates a patch pixel's taps with fp32 atomics, whose order varies);
  * device-side counters: placement index, AdamW step count, learning rate changes between outer iterations;
  * accumulate-only iterations (accumulate_steps > 1 of TMA / UPA) and the buffer reset after a stepping iteration;
  * two NCCL ranks x bs 2 reproduce one GPU x bs 4 on the concatenated batch, with bit-identical patches across ranks
    (UADA_ddp.py:140-166,206) -- skipped with fewer than two GPUs.
"""
import os
import random
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200 import _lib, labels as lab  # noqa: E402
from roboticattack_b200.config import tiny  # noqa: E402
from roboticattack_b200.engine import LossSpec, VLAEngine  # noqa: E402
from roboticattack_b200.synthetic import draw_placements, synthetic_batch  # noqa: E402
from roboticattack_b200.weights import random_state_dict  # noqa: E402

CFG = dict(img=56, llm_layers=2, vit_depth=3)
B, T, P, STEPS, LR = 3, 16, 12, 6, 2e-3


@pytest.fixture(scope="module")
def setup():
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test")
    eng = VLAEngine(cfg, B, T)
    eng.load_state_dict(sd)
    batch = synthetic_batch(cfg, B, T, seed=11, ragged=True)
    batch["labels"] = lab.mask_labels_uada(batch["labels"].clone(), [0, 1, 2])
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(B, (cfg.img, cfg.img), (P, P), True, steps=STEPS)
    eng.set_batch(batch["obs"], batch["input_ids"], batch["attention_mask"], batch["labels"])
    eng.set_placements(xy, theta)
    torch.manual_seed(42)
    patch0 = torch.rand(3, P, P).cuda()
    return cfg, eng, patch0


def run_steps(eng, patch0, mode, loss, opt=_lib.OPT_ADAMW, steps=STEPS, lrs=None):
    p = patch0.clone()
    m, v, g = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    hist = torch.zeros(steps, _lib.NUM_SCALARS, device="cuda")
    pred = torch.zeros(max(eng.num_supervised, 1), dtype=torch.int32, device="cuda")
    lrs = lrs or [LR] * steps
    if mode == "composed":        # round-1 path: two C-ABI calls per step, counters on the host
        for s in range(steps):
            eng.fwd_bwd(p, s, _lib.FE_WARP, loss, g, hist[s], pred)
            eng.patch_update(p, g, m, v, s + 1, lrs[s], kind=opt, scalars=hist[s])
    else:
        eng.set_step_state(0, 0)
        for s in range(steps):
            eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, lrs[s], opt_kind=opt, graph=(mode == "graph"))
    torch.cuda.synchronize()
    return p.cpu(), hist.cpu(), pred.cpu(), m.cpu(), v.cpu()


def test_graph_replay_equals_eager_equals_composed(setup):
    cfg, eng, patch0 = setup
    loss = LossSpec(_lib.LOSS_UADA, 5.0)
    L = _lib.lib()
    ref = run_steps(eng, patch0, "composed", loss)
    eag = run_steps(eng, patch0, "eager", loss)
    r0 = L.vla_graph_replays()
    gra = run_steps(eng, patch0, "graph", loss)          # call 1 eager (autotune), call 2 records, calls 2.. replay
    assert L.vla_graph_replays() - r0 == STEPS - 1
    assert eng.graph_kernel_nodes > 50, "the recorded graph must hold the step's kernels"
    r1 = L.vla_graph_replays()
    gra2 = run_steps(eng, patch0, "graph", loss)         # fresh buffers: a new recording, or -- when the allocator hands back the
    assert L.vla_graph_replays() - r1 >= STEPS - 1       # same addresses -- the recorded graph replayed from the first step on
    assert eng.get_step_state() == (STEPS, STEPS)
    for name, out in (("eager", eag), ("graph", gra), ("graph again", gra2)):
        assert torch.equal(out[1][0, :_lib.S_GRAD_MEAN], ref[1][0, :_lib.S_GRAD_MEAN]), f"{name}: first-step scalars must be bit-identical"
        torch.testing.assert_close(out[1], ref[1], rtol=2e-4, atol=1e-6, msg=lambda m: f"{name}: scalar history\n{m}")
        assert (out[0] - ref[0]).abs().max().item() <= 2 * LR * 1.001, name      # at most a sign flip of a ~zero gradient
        assert (out[0] - ref[0]).abs().mean().item() < 1e-5, name
        assert torch.equal(out[2], ref[2]), f"{name}: predicted ids"
    assert (ref[0] - patch0.cpu()).abs().max() > 0


def test_single_sample_step_runs_the_towers_on_two_streams_inside_the_graph():
    """Per-GPU batch 1 (BASELINE config #1): the two vision towers run as two chains on two streams (a tower's GEMMs have a
    handful of tiles there), fork / join recorded into the step's CUDA graph.  Same arithmetic as the lock-step form: the
    forward is bit-identical, graph replay == eager, and forcing the lock-step form gives the same scalars."""
    import os
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test")
    eng = VLAEngine(cfg, 1, T)
    eng.load_state_dict(sd)
    batch = synthetic_batch(cfg, 1, T, seed=12)
    batch["labels"] = lab.mask_labels_uada(batch["labels"].clone(), [0, 1, 2])
    random.seed(7)
    np.random.seed(7)
    xy, theta = draw_placements(1, (cfg.img, cfg.img), (P, P), True, steps=STEPS)
    eng.set_batch(batch["obs"], batch["input_ids"], batch["attention_mask"], batch["labels"])
    eng.set_placements(xy, theta)
    torch.manual_seed(1)
    patch0 = torch.rand(3, P, P).cuda()
    loss = LossSpec(_lib.LOSS_UADA, 5.0)
    assert "VLA_TOWERS" not in os.environ
    eag = run_steps(eng, patch0, "eager", loss)
    r0 = _lib.lib().vla_graph_replays()
    gra = run_steps(eng, patch0, "graph", loss)
    assert _lib.lib().vla_graph_replays() - r0 == STEPS - 1
    os.environ["VLA_TOWERS"] = "lockstep"
    try:
        lock = run_steps(eng, patch0, "eager", loss)
    finally:
        del os.environ["VLA_TOWERS"]
    for name, out in (("graph", gra), ("lock-step towers", lock)):
        assert torch.equal(out[1][0, :_lib.S_GRAD_MEAN], eag[1][0, :_lib.S_GRAD_MEAN]), f"{name}: first-step scalars must be bit-identical"
        torch.testing.assert_close(out[1], eag[1], rtol=2e-4, atol=1e-6, msg=lambda m: f"{name}: scalar history\n{m}")
        assert (out[0] - eag[0]).abs().max().item() <= 2 * LR * 1.001, name
        assert torch.equal(out[2], eag[2]), f"{name}: predicted ids"
    assert (eag[0] - patch0.cpu()).abs().max() > 0


def test_learning_rate_and_counters_live_on_the_device(setup):
    """sign-PGD moves every pixel with a non-zero gradient by exactly lr: the per-outer-iteration learning rate reaches the
    replayed graph through device memory."""
    cfg, eng, patch0 = setup
    loss = LossSpec(_lib.LOSS_CE)
    mid = torch.full_like(patch0, 0.5)
    lrs = [1e-3, 1e-3, 1e-3, 4e-3, 4e-3, 4e-3]
    p = mid.clone()
    m, v, g = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    hist = torch.zeros(STEPS, _lib.NUM_SCALARS, device="cuda")
    pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
    eng.set_step_state(0, 0)
    prev = p.clone()
    for s in range(STEPS):
        eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, lrs[s], opt_kind=_lib.OPT_PGD)
        torch.cuda.synchronize()
        d = (p - prev).abs()
        moved = d[g != 0]
        assert moved.numel() > 0 and torch.allclose(moved, torch.full_like(moved, lrs[s]), rtol=0, atol=1e-7), (s, moved.min(), moved.max())
        prev = p.clone()
    assert eng.get_step_state() == (STEPS, STEPS)
    # running past the uploaded placements is refused, not silently clamped
    with pytest.raises(_lib.VLAError, match="placement"):
        eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, LR, opt_kind=_lib.OPT_PGD)
    # set_step_state rewinds the placement index and keeps / restores the optimiser step (resume)
    eng.set_step_state(2, 40)
    eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, LR, opt_kind=_lib.OPT_ADAMW)
    assert eng.get_step_state() == (3, 41)


def test_accumulate_only_iterations(setup):
    """accumulate_steps = 2 (TMA.py:162-170): the first iteration only adds its gradient to the buffer, the second one
    steps on the sum and clears the buffer."""
    cfg, eng, patch0 = setup
    loss = LossSpec(_lib.LOSS_CE, ce_scale=0.5)
    p = patch0.clone()
    m, v, g, acc = (torch.zeros_like(p) for _ in range(4))
    hist = torch.zeros(STEPS, _lib.NUM_SCALARS, device="cuda")
    pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
    for graph in (False, True, True):            # eager, recording, replay
        p.copy_(patch0)
        for t in (m, v, acc):
            t.zero_()
        eng.set_step_state(0, 0)
        eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, LR, accumulate=acc, do_update=False, graph=graph)
        torch.cuda.synchronize()
        g0 = g.clone()
        assert torch.equal(p, patch0) and torch.equal(acc, g0) and eng.get_step_state() == (1, 0)
        eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, LR, accumulate=acc, do_update=True, graph=graph)
        torch.cuda.synchronize()
        assert acc.abs().max().item() == 0 and eng.get_step_state() == (2, 1)
        # reference: one AdamW step on g0 + g1 from zero moments
        gs = g0 + g
        expect = (patch0 - LR * (0.001 ** 0.5) / 0.1 * (0.1 * gs) / ((0.001 * gs * gs).sqrt() + 1e-6)).clamp(0, 1)
        torch.testing.assert_close(p, expect, rtol=0, atol=2e-6)
        assert hist[0, _lib.S_GRAD_MEAN].item() == 0 and abs(hist[1, _lib.S_GRAD_MEAN].item() - gs.mean().item()) < 1e-6 + 1e-4 * abs(gs.mean().item())


# ------------------------------------------------------------------------------------------------ two NCCL ranks
WORLD, B_RANK = 2, 2


def _rank_main(rank, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device(f"cuda:{rank}"))
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test")
    eng = VLAEngine(cfg, B_RANK, T, device=f"cuda:{rank}")
    eng.load_state_dict(sd)
    comm = eng.make_comm(rank, WORLD)
    full = synthetic_batch(cfg, WORLD * B_RANK, T, seed=21)
    full["labels"] = lab.mask_labels_uada(full["labels"].clone(), [0, 1, 2])
    sl = slice(rank * B_RANK, (rank + 1) * B_RANK)
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(WORLD * B_RANK, (cfg.img, cfg.img), (P, P), True, steps=STEPS)
    eng.set_batch(full["obs"][sl], full["input_ids"][sl], full["attention_mask"][sl], full["labels"][sl])
    eng.set_placements(xy[:, sl], theta[:, sl])
    torch.manual_seed(42 + rank)
    p = torch.rand(3, P, P).cuda()
    dist.broadcast(p, src=0)
    m, v, g = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    hist = torch.zeros(STEPS, _lib.NUM_SCALARS, device="cuda")
    pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
    loss = LossSpec(_lib.LOSS_UADA_DDP, 5.0)
    eng.set_step_state(0, 0)
    for s in range(STEPS):
        eng.attack_step(p, m, v, g, hist, pred, _lib.FE_WARP, loss, LR, comm=comm)
    torch.cuda.synchronize()
    torch.save({"patch": p.cpu(), "hist": hist.cpu(), "replays": _lib.lib().vla_graph_replays()}, os.path.join(outdir, f"rank{rank}.pt"))
    comm.close()
    dist.destroy_process_group()


def test_two_nccl_ranks_equal_one_gpu_on_the_concatenated_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_rank_main, args=(port, d), nprocs=WORLD, join=True)
        outs = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(WORLD)]
    assert torch.equal(outs[0]["patch"], outs[1]["patch"]), "patches must stay bit-identical across ranks"
    assert outs[0]["replays"] == STEPS - 1, "the all-reduce must be part of the replayed graph"
    # one GPU, the concatenated batch, same placements, same loss (mean over all tokens == mean of the per-rank means)
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test")
    eng = VLAEngine(cfg, WORLD * B_RANK, T)
    eng.load_state_dict(sd)
    full = synthetic_batch(cfg, WORLD * B_RANK, T, seed=21)
    full["labels"] = lab.mask_labels_uada(full["labels"].clone(), [0, 1, 2])
    random.seed(42)
    np.random.seed(42)
    xy, theta = draw_placements(WORLD * B_RANK, (cfg.img, cfg.img), (P, P), True, steps=STEPS)
    eng.set_batch(full["obs"], full["input_ids"], full["attention_mask"], full["labels"])
    eng.set_placements(xy, theta)
    torch.manual_seed(42)
    p0 = torch.rand(3, P, P).cuda()
    one = run_steps(eng, p0, "graph", LossSpec(_lib.LOSS_UADA_DDP, 5.0))
    mean_hist = (outs[0]["hist"] + outs[1]["hist"]) / 2
    torch.testing.assert_close(mean_hist[:, _lib.S_LOSS], one[1][:, _lib.S_LOSS], rtol=2e-2, atol=1e-4)
    dev = (outs[0]["patch"] - one[0]).abs()
    assert dev.mean().item() < LR / 4 and (dev < LR / 2).float().mean().item() > 0.9, (dev.mean().item(), dev.max().item())
