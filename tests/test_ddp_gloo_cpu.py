"""World-size-2 run (gloo, CPU) of the data-parallel attack loop: the batch is sharded over ranks, the patch gradient is
all-reduced every inner step, the replicated update keeps the patches bit-identical across ranks and equal to the
serially computed DDP semantics (mean of per-rank gradients; UADA_ddp.py:140-209)."""
import os
import random
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from roboticattack_b200.config import tiny
from roboticattack_b200.synthetic import draw_placements, synthetic_batch
from roboticattack_b200.weights import random_state_dict

CFG = dict(img=28, llm_layers=1, vit_depth=2)
B_PER_RANK, T, P_HW, INNER, OUTER, LR, WARMUP = 2, 14, 8, 2, 3, 2e-3, 1


def shard_batches(rank):
    cfg = tiny(**CFG)
    out = []
    for i in range(OUTER):
        b = synthetic_batch(cfg, B_PER_RANK, T, seed=100 + 10 * i + rank)
        out.append({"pixel_values": b["obs"], "input_ids": b["input_ids"], "attention_mask": b["attention_mask"], "labels": b["labels"]})
    return out


def _worker_val(rank, world, port, outdir):
    """Same loop with the per-rank validation pass (UADA_ddp.py:232-325) every outer iteration."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.attacker import UADADDPAttacker
    torch.set_num_threads(2)
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42 + rank)
    att = UADADDPAttacker(sd, save_dir=os.path.join(outdir, "run"), patch_size=[3, P_HW, P_HW], lr=LR, bs=B_PER_RANK, warmup=WARMUP,
                          num_iter=2, maskidx=[0, 1, 2], innerLoop=1, geometry=True, use_wandb=False, MSE_weights=5, cfg=cfg,
                          device="cpu", engine_factory=OracleEngine, backend="gloo")
    att.val_every, att.val_batches = 1, 2
    patch = att.attack(rank, world, train_dataloader=shard_batches(rank), val_dataloader=shard_batches(rank + 10))
    torch.save({"patch": patch, "val": (att.val_CE_loss, att.val_MSE_Distance, att.val_UAD)}, os.path.join(outdir, f"rank{rank}.pt"))
    torch.distributed.destroy_process_group()


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.attacker import UADADDPAttacker
    torch.set_num_threads(2)
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    random.seed(42)
    np.random.seed(42)
    torch.manual_seed(42 + rank)            # rank 0's torch.rand patch must win through the broadcast
    att = UADADDPAttacker(sd, save_dir="", patch_size=[3, P_HW, P_HW], lr=LR, bs=B_PER_RANK, warmup=WARMUP, num_iter=OUTER,
                          maskidx=[0, 1, 2], innerLoop=INNER, geometry=True, use_wandb=False, MSE_weights=5, cfg=cfg,
                          device="cpu", engine_factory=OracleEngine, backend="gloo")
    patch = att.attack(rank, world, train_dataloader=shard_batches(rank))
    torch.save({"patch": patch, "logs": att.train_logs}, os.path.join(outdir, f"rank{rank}.pt"))
    torch.distributed.destroy_process_group()


def serial_reference(world):
    """The same semantics computed in one process: per-rank oracle gradients averaged, one replicated AdamW update."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200 import _lib, labels as lab
    from roboticattack_b200.attacker import cosine_with_warmup
    from roboticattack_b200.engine import LossSpec
    cfg = tiny(**CFG)
    sd = random_state_dict(cfg, seed=0, dtype=torch.float32, init="test")
    engines = []
    for r in range(world):
        e = OracleEngine(cfg, B_PER_RANK, T)
        e.load_state_dict(sd)
        engines.append(e)
    torch.manual_seed(42)
    patch = torch.rand(3, P_HW, P_HW)
    m, v = torch.zeros_like(patch), torch.zeros_like(patch)
    shards = [shard_batches(r) for r in range(world)]
    loss = LossSpec(_lib.LOSS_UADA_DDP, mse_weight=5.0)
    random.seed(42)
    np.random.seed(42)
    t = 0
    for i in range(OUTER):
        xy, th = draw_placements(B_PER_RANK, (cfg.img, cfg.img), (P_HW, P_HW), True, steps=INNER)   # same stream on every rank
        for r in range(world):
            d = shards[r][i]
            engines[r].set_batch(d["pixel_values"], d["input_ids"], d["attention_mask"], lab.mask_labels_uada(d["labels"].clone(), [0, 1, 2]))
            engines[r].set_placements(xy, th)
        lr = LR * cosine_with_warmup(i, WARMUP, OUTER)
        for s in range(INNER):
            gs = []
            for r in range(world):
                g = torch.zeros_like(patch)
                sc = torch.zeros(_lib.NUM_SCALARS)
                pr = torch.zeros(engines[r].num_supervised, dtype=torch.int32)
                engines[r].fwd_bwd(patch, s, _lib.FE_WARP, loss, g, sc, pr)
                gs.append(g)
            t += 1
            engines[0].patch_update(patch, torch.stack(gs).sum(0), m, v, t, lr, grad_scale=1.0 / world)
    return patch


def test_ddp_world2_matches_serial_semantics():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, d), nprocs=world, join=True)
        outs = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(world)]
    assert torch.equal(outs[0]["patch"], outs[1]["patch"]), "patches diverged across ranks"
    ref = serial_reference(world)
    torch.testing.assert_close(outs[0]["patch"], ref, rtol=0, atol=1e-6)
    assert (ref - torch.rand(3, P_HW, P_HW, generator=torch.Generator().manual_seed(42))).abs().max() > 0
    # the packed metric reduction: CE / loss / UAD are means over ranks, identical on both
    assert outs[0]["logs"] == outs[1]["logs"] and len(outs[0]["logs"]) == OUTER


def test_ddp_world2_validation():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker_val, args=(world, port, d), nprocs=world, join=True)
        outs = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(world)]
        assert torch.equal(outs[0]["patch"], outs[1]["patch"]), "validation must not desynchronise the ranks' RNG streams"
        ce, mse, uad = outs[0]["val"]
        assert len(ce) == len(mse) == len(uad) == 2 and all(np.isfinite(v) for v in ce + mse + uad)
        assert outs[1]["val"] == ([], [], []), "only rank 0 records the reduced validation metrics"
        assert os.path.exists(os.path.join(d, "run", "last", "patch.pt")) and os.path.exists(os.path.join(d, "run", "0", "patch.pt"))
        assert os.path.exists(os.path.join(d, "run", "last", "attack_state.pt"))


def _loaders(rank, world):
    return shard_batches(rank), None


def test_ddp_run_classmethod_spawns_ranks():
    """``OpenVLAAttacker.run(...)`` of UADA_ddp.py:327-344: spawns one process per rank, each loads the weights from
    ``vla_path`` and its own data shard, runs ``attack`` and tears the process group down; rank 0 leaves ``last/patch.pt``."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from roboticattack_b200.white_patch.UADA_ddp import OpenVLAAttacker
    cfg = tiny(**CFG)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    old = {k: os.environ.get(k) for k in ("MASTER_ADDR", "MASTER_PORT")}
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "vla.pt")
            torch.save(random_state_dict(cfg, seed=0, dtype=torch.float32, init="test"), path)
            OpenVLAAttacker.run(path, _loaders, os.path.join(d, "run"), False, [3, P_HW, P_HW], LR, B_PER_RANK, WARMUP, 2, [0, 1, 2], 1, True,
                                False, 5, world_size=2, cfg=cfg, device="cpu", engine_factory=OracleEngine, backend="gloo")
            patch = torch.load(os.path.join(d, "run", "last", "patch.pt"), weights_only=True)
            assert patch.shape == (3, P_HW, P_HW) and patch.dtype == torch.float32 and 0 <= patch.min() and patch.max() <= 1
            assert os.path.exists(os.path.join(d, "run", "last", "attack_state.pt"))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
