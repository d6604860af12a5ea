"""Stage-by-stage probe of the NCCL path of vla_attack_step on 2+ GPUs (prints a line per stage; a stall shows where).
usage: torchrun --nproc-per-node 2 tools/nccl_probe.py"""
import faulthandler
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(int(os.environ.get("PROBE_TIMEOUT", "75")), exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
T0 = time.time()


def say(msg):
    print(f"[{time.time() - T0:6.1f}s rank {rank}] {msg}", flush=True)


dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
say("process group up")
from roboticattack_b200 import _lib, labels as lab
from roboticattack_b200.config import tiny
from roboticattack_b200.engine import LossSpec, VLAEngine
from roboticattack_b200.synthetic import draw_placements, synthetic_batch
from roboticattack_b200.weights import random_state_dict

cfg = tiny(img=56, llm_layers=2, vit_depth=3)
eng = VLAEngine(cfg, 2, 16, device=f"cuda:{rank}")
eng.load_state_dict(random_state_dict(cfg, seed=0, dtype=torch.bfloat16, init="test"))
say("engine loaded")
x = torch.ones(4, device="cuda")
dist.all_reduce(x)
torch.cuda.synchronize()
say(f"torch all_reduce ok {x[0].item()}")
comm = eng.make_comm(rank, world)
say("vla_comm created")
g = torch.full((7500,), float(rank + 1), device="cuda")
comm.all_reduce_(g)
torch.cuda.synchronize()
say(f"vla_allreduce_patch_grad eager ok: {g[0].item()} (expect {world * (world + 1) / 2})")
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    comm.all_reduce_(g)
torch.cuda.synchronize()
say("vla_allreduce_patch_grad on a non-default stream ok")
b = synthetic_batch(cfg, 2, 16, seed=5 + rank)
b["labels"] = lab.mask_labels_uada(b["labels"].clone(), [0, 1, 2])
random.seed(1)
np.random.seed(1)
xy, th = draw_placements(2, (cfg.img, cfg.img), (12, 12), True, steps=8)
eng.set_batch(b["obs"], b["input_ids"], b["attention_mask"], b["labels"])
eng.set_placements(xy, th)
p = torch.rand(3, 12, 12, device="cuda")
dist.broadcast(p, src=0)
m, v, gr = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
hist = torch.zeros(8, _lib.NUM_SCALARS, device="cuda")
pred = torch.zeros(eng.num_supervised, dtype=torch.int32, device="cuda")
loss = LossSpec(_lib.LOSS_UADA_DDP, 5.0)
eng.set_step_state(0, 0)
for s in range(3):
    eng.attack_step(p, m, v, gr, hist, pred, _lib.FE_WARP, loss, 2e-3, comm=comm, graph=False)
    torch.cuda.synchronize()
    say(f"eager attack_step {s} ok loss {hist[s, 0].item():.4f}")
for s in range(3, 8):
    eng.attack_step(p, m, v, gr, hist, pred, _lib.FE_WARP, loss, 2e-3, comm=comm, graph=True)
    say(f"graph-mode attack_step {s} issued (replays so far {_lib.lib().vla_graph_replays()})")
    torch.cuda.synchronize()
    say(f"graph-mode attack_step {s} done loss {hist[s, 0].item():.4f}")
ps = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(ps, p)
say(f"patches identical across ranks: {all(torch.equal(ps[0], q) for q in ps)}")
comm.close()
dist.destroy_process_group()
say("done")
