"""Drop-in proof at the wrapper level: the reference's OWN, unmodified launch scripts -- VLAAttacker/UADA_wrapper.py,
UPA_wrapper.py, TMA_wrapper.py and UADA_wrapper_ddp.py (``main(args)``, the function scripts/run_{UADA,UPA,TMA}.sh and torchrun
execute) -- run against this package's ``white_patch`` modules and write the reference's run directory.

What is stubbed is everything the tier puts out of scope and that is not installed here: the HF / prismatic model
loading (``AutoModelForVision2Seq.from_pretrained`` hands back a toy OpenVLA with an HF-style ``config`` and ``state_dict``), the
RLDS dataloader module (``openvla_dataloader.get_dataloader`` / ``get_dataset`` hand back collator-style batches with PIL
images) and ``wandb``.  The attack classes are resolved exactly as INTEGRATION.md's "PYTHONPATH" option does it: ``white_patch``
(and the bare ``TMA`` / ``UPA`` names the wrappers import after their ``sys.path`` hack) resolve to ``roboticattack_b200/white_patch``.
On this CPU-only host the engine behind the classes is the oracle engine (tests/oracle_engine.py); on a GPU box the same
call path constructs ``VLAEngine``.

The wrappers are read from /root/reference at test time (never copied); the test is skipped where the reference is absent.
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/VLAAttacker"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not mounted here")

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from roboticattack_b200.config import tiny  # noqa: E402
from roboticattack_b200.synthetic import synthetic_batch  # noqa: E402
from roboticattack_b200.weights import random_state_dict  # noqa: E402

CFG = tiny(img=28, llm_layers=1, vit_depth=2)
T = 14


class ToyHFConfig:
    """The fields of the HF OpenVLAConfig the engine configuration is derived from."""
    timm_model_ids = ["vit_large_patch14_reg4_dinov2.lvd142m", "vit_so400m_patch14_siglip_224"]
    pad_to_multiple_of = 64
    text_config = {"hidden_size": CFG.llm.hidden, "num_hidden_layers": CFG.llm.layers, "num_attention_heads": CFG.llm.heads,
                   "intermediate_size": CFG.llm.ffn, "vocab_size": 32000, "rms_norm_eps": CFG.llm.rms_eps, "rope_theta": CFG.llm.rope_theta}


class ToyVLA:
    """Stands in for OpenVLAForActionPrediction: ``.to()``, ``.state_dict()``, ``.config`` -- and ``.cfg`` carrying the toy tower
    widths (an HF config only names the timm towers; the real ones are fixed by those names)."""

    def __init__(self):
        self.sd = random_state_dict(CFG, seed=0, dtype=torch.float32, init="test")
        self.config, self.cfg = ToyHFConfig(), CFG
        self.device = torch.device("cpu")

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(self.sd.values())

    def state_dict(self):
        return self.sd


class _Auto:
    loaded = []

    @classmethod
    def register(cls, *a, **k):
        pass

    @classmethod
    def from_pretrained(cls, path, **kw):
        cls.loaded.append((cls.__name__, path))
        if cls.__name__ == "AutoModelForVision2Seq":
            return ToyVLA()
        return types.SimpleNamespace(tokenizer=types.SimpleNamespace(model_max_length=2048, pad_token_id=32000))


def collator_batches(n, B, seed):
    from PIL import Image
    out = []
    for i in range(n):
        b = synthetic_batch(CFG, B, T, seed=seed + i, ragged=(i % 2 == 1))
        out.append({"pixel_values": [Image.fromarray(im.numpy()) for im in b["obs"]], "input_ids": b["input_ids"],
                    "attention_mask": b["attention_mask"], "labels": b["labels"]})
    return out


class ShardableDataset(list):
    """What get_dataset returns in the reference is an RLDS IterableDataset with ``.shard``; batches here are pre-collated."""
    shards = []

    def shard(self, num_shards, index):
        ShardableDataset.shards.append((num_shards, index))
        return ShardableDataset(self[index::num_shards])


@pytest.fixture
def reference_env(monkeypatch, tmp_path):
    import transformers
    from oracle_engine import OracleEngine
    import roboticattack_b200.attacker as attacker
    import roboticattack_b200.white_patch as wp
    from roboticattack_b200.white_patch import TMA, UADA, UADA_ddp, UPA, appply_random_transform

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    autos = {n: type(n, (_Auto,), {}) for n in ("AutoConfig", "AutoProcessor", "AutoModelForVision2Seq", "AutoImageProcessor")}
    for n in autos:                       # resolve the lazy attributes first: transformers swaps its module object while doing so
        getattr(transformers, n, None)
    transformers = sys.modules["transformers"]
    for n, c in autos.items():
        monkeypatch.setattr(transformers, n, c, raising=False)
    mod("prismatic"), mod("prismatic.extern"), mod("prismatic.extern.hf")
    mod("prismatic.extern.hf.configuration_prismatic", OpenVLAConfig=ToyHFConfig)
    mod("prismatic.extern.hf.processing_prismatic", PrismaticProcessor=object, PrismaticImageProcessor=object)
    mod("prismatic.extern.hf.modeling_prismatic", OpenVLAForActionPrediction=ToyVLA)
    mod("wandb", init=lambda **k: None, log=lambda *a, **k: None, config={})
    calls = []

    def get_dataloader(batch_size, dataset, server=None, vla_path=None):
        calls.append(("get_dataloader", batch_size, dataset))
        return collator_batches(3, batch_size, 10), collator_batches(2, batch_size, 50)

    def get_dataset(dataset):
        calls.append(("get_dataset", dataset))
        return ShardableDataset(collator_batches(3, 2, 10)), ShardableDataset(collator_batches(2, 2, 50))

    loader_mod = mod("openvla_dataloader", get_dataloader=get_dataloader, get_bridge_dataloader=get_dataloader, get_dataset=get_dataset)
    # INTEGRATION.md, "PYTHONPATH" option: the wrappers' module names resolve to this package
    for name, m in (("white_patch", wp), ("white_patch.UADA", UADA), ("white_patch.UPA", UPA), ("white_patch.TMA", TMA),
                    ("white_patch.UADA_ddp", UADA_ddp), ("white_patch.appply_random_transform", appply_random_transform),
                    ("TMA", TMA), ("UPA", UPA), ("UADA", UADA), ("appply_random_transform", appply_random_transform),
                    ("white_patch.openvla_dataloader", loader_mod)):
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.setattr(attacker, "VLAEngine", OracleEngine)          # CPU-only host: no CUDA engine can be constructed
    monkeypatch.setattr(torch.utils.data, "DataLoader", lambda ds, batch_size=1, collate_fn=None: ds)   # batches are pre-collated
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("VLA_PREFETCH", "1")
    return calls, tmp_path


def load_wrapper(name):
    spec = importlib.util.spec_from_file_location(f"reference_{name}", os.path.join(REF, f"{name}.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def base_args(tmp_path, **over):
    a = dict(maskidx=[0, 1, 2], lr=2e-3, device=0, iter=2, accumulate=1, bs=2, warmup=0, tags=[""], filterGripTrainTo1=False,
             geometry=True, patch_size=[3, 8, 8], wandb_project="false", wandb_entity="x", innerLoop=2, dataset="bridge_orig",
             resize_patch=False, reverse_direction=True, server=str(tmp_path), targetAction=0, alpha=0.8, belta=0.2, MSE_weights=5,
             guide=False, colorjitter=False)
    a.update(over)
    return argparse.Namespace(**a)


def only_run_dir(root):
    runs = [os.path.join(root, d) for d in os.listdir(root)]
    assert len(runs) == 1, runs
    return runs[0]


def check_patch(run_dir, shape=(3, 8, 8)):
    p = torch.load(os.path.join(run_dir, "last", "patch.pt"), weights_only=True)
    assert p.dtype == torch.float32 and tuple(p.shape) == shape and 0 <= p.min() and p.max() <= 1
    torch.manual_seed(42)
    assert (p - torch.rand(*shape)).abs().max() > 0, "the saved patch is the initial one: no attack step ran"
    return p


def test_uada_wrapper_main_runs_unmodified(reference_env):
    calls, tmp = reference_env
    w = load_wrapper("UADA_wrapper")
    w.main(base_args(tmp))
    assert ("get_dataloader", 2, "bridge_orig") in calls
    run = only_run_dir(os.path.join(tmp, "run", "UADA"))
    check_patch(run)
    assert os.path.exists(os.path.join(run, "0", "patch.pt")), "iteration 0 validates and keeps the best patch (UADA.py:193-258)"
    assert os.path.exists(os.path.join(run, "train_CE_loss.pkl")) and os.path.isdir(os.path.join(run, "last", "val_related_data"))


def test_tma_wrapper_main_runs_unmodified(reference_env):
    calls, tmp = reference_env
    w = load_wrapper("TMA_wrapper")
    w.main(base_args(tmp, lr=1 / 255))
    run = only_run_dir(os.path.join(tmp, "run", "white_patch_attack"))
    check_patch(run)
    assert os.path.exists(os.path.join(run, "val_L1_loss.pkl"))


def test_upa_wrapper_main_runs_unmodified(reference_env):
    calls, tmp = reference_env
    w = load_wrapper("UPA_wrapper")
    w.main(base_args(tmp))
    roots = [d for d in os.listdir(os.path.join(tmp, "run"))]
    run = only_run_dir(os.path.join(tmp, "run", roots[0]))
    check_patch(run)


def test_uada_ddp_wrapper_main_runs_unmodified(reference_env, monkeypatch):
    """UADA_wrapper_ddp.main -> OpenVLAAttacker._attack_entry(rank, {"vla_path": ..., "dataset_name": ..., ...}, world): the
    constructor takes the reference's own keyword names; the model comes from AutoModelForVision2Seq.from_pretrained(vla_path),
    the data from white_patch.openvla_dataloader.get_dataset(dataset_name).shard(world, rank) (UADA_ddp.py:37-55,157-160).
    The wrapper calls dist.broadcast_object_list before any init_process_group (its own init is commented out, :38), so --
    as under a launcher that has initialised the group -- a one-rank gloo group exists before main() runs."""
    import socket
    import torch.distributed as dist
    calls, tmp = reference_env
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", str(port))
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "1")
    import roboticattack_b200.attacker as attacker
    real_init = attacker.UADADDPAttacker.__init__

    def cpu_init(self, *a, **k):          # the wrapper has no device flag; this host has no GPU
        k.setdefault("device", "cpu")
        k.setdefault("backend", "gloo")
        real_init(self, *a, **k)
        self.val_every, self.val_batches = 1, 1

    monkeypatch.setattr(attacker.UADADDPAttacker, "__init__", cpu_init)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        w = load_wrapper("UADA_wrapper_ddp")
        w.main(base_args(tmp))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
    assert ("get_dataset", "bridge_orig") in calls and (1, 0) in ShardableDataset.shards
    assert ("AutoModelForVision2Seq", "openvla/openvla-7b") in _Auto.loaded
    run = only_run_dir(os.path.join(tmp, "run", "UADA"))
    check_patch(run)
