"""ms per action of ActionPolicy.predict_action at OpenVLA-7B shapes, batch 1: KV-cache decode vs one full forward per token.
usage (GPU box): python tools/decode_bench.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from roboticattack_b200.config import openvla_7b
from roboticattack_b200.engine import VLAEngine
from roboticattack_b200.policy import ActionPolicy
from roboticattack_b200.synthetic import synthetic_batch

cfg = openvla_7b()
eng = VLAEngine(cfg, 1, 33)
eng.load_random_weights(seed=0, init="reference")
b = synthetic_batch(cfg, 1, 33, seed=3)
prompt = b["input_ids"][:, :25]
pol = ActionPolicy(eng)
if "--profile" in sys.argv:   # under ncu: one recording call, one replayed call, nothing else
    for _ in range(2):
        pol.generate_action_tokens(b["obs"], prompt, 7, kv_cache=True)
    torch.cuda.synchronize()
    sys.exit(0)
for kv in (True, False):
    for _ in range(2):
        pol.generate_action_tokens(b["obs"], prompt, 7, kv_cache=kv)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        toks = pol.generate_action_tokens(b["obs"], prompt, 7, kv_cache=kv)
    torch.cuda.synchronize()
    t7 = (time.perf_counter() - t0) / n * 1e3
    line = f"predict_action tokens, bs 1, prompt 25, 7 tokens, kv_cache={kv}: {t7:.1f} ms per action"
    if kv:   # prefill + first token alone, to split the action into prefill and per-token decode
        pol.generate_action_tokens(b["obs"], prompt, 1, kv_cache=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            pol.generate_action_tokens(b["obs"], prompt, 1, kv_cache=True)
        torch.cuda.synchronize()
        t1 = (time.perf_counter() - t0) / n * 1e3
        line += f" (prefill + token 0: {t1:.1f} ms, then {(t7 - t1) / 6:.2f} ms per token = {13.22 / ((t7 - t1) / 6):.2f} TB/s of weights)"
    print(line + f"  tokens {toks[0].tolist()}")
