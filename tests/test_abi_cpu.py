"""The C-ABI library loads on a machine without a GPU and exports every symbol include/vla_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from roboticattack_b200.build import build_extension
    return build_extension()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vla_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vla_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(str(built))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in vla_b200.h but not exported: {missing}"


def test_binding_covers_header(built):
    from roboticattack_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    h = _lib.lib()
    assert h.vla_abi_version() == 3
    assert h.vla_launch_count() >= 0


def test_errors_are_reported_not_swallowed(built):
    """Host-side validation works without a GPU: bad arguments -> non-zero rc + message (no compute call is made)."""
    from roboticattack_b200 import _lib
    h = _lib.lib()
    cfg = _lib.Config()
    cfg.img, cfg.patch = 225, 14            # not a multiple of the ViT patch
    out = ctypes.c_void_p()
    rc = h.vla_engine_create(ctypes.byref(cfg), ctypes.byref(out))
    assert rc != 0 and b"multiple" in h.vla_last_error()
    with pytest.raises(_lib.VLAError):
        _lib.check(rc, "vla_engine_create")
    rc = h.vla_patch_update(None, None, None, None, 0, 0, 0.0, 0.9, 0.999, 1e-6, 0, 1.0, 0.0, None, None)
    assert rc != 0


def test_engine_planning_without_gpu(built):
    """Arena sizes are pure host arithmetic: OpenVLA-7B at bs=8 needs ~30 GB of weights and ~10 GB of activations."""
    from roboticattack_b200 import _lib
    from roboticattack_b200.config import openvla_7b
    from roboticattack_b200.engine import _c_config
    h = _lib.lib()
    c = _c_config(openvla_7b())
    e = ctypes.c_void_p()
    _lib.check(h.vla_engine_create(ctypes.byref(c), ctypes.byref(e)))
    wb = h.vla_engine_weight_bytes(e)
    sb = h.vla_engine_workspace_bytes(e, 8, 33)
    h.vla_engine_destroy(e)
    assert 29e9 < wb < 32e9, wb
    assert 8e9 < sb < 16e9, sb


def test_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from roboticattack_b200 import _lib
    from roboticattack_b200.config import tiny
    from roboticattack_b200.engine import VLAEngine
    with pytest.raises(_lib.VLAError):
        VLAEngine(tiny(), 1, 16)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the product package (nor bench.py outside its cpu_baseline /
    --impl reference legs, nor the tools) may import it, directly or through another product module."""
    import ast
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent

    def imports(path):
        out = []
        for node in ast.walk(ast.parse(path.read_text())):
            if isinstance(node, ast.Import):
                out += [(a.name, node.lineno) for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                out.append((node.module, node.lineno))
        return out

    for path in sorted((root / "roboticattack_b200").rglob("*.py")):
        bad = [(m, ln) for m, ln in imports(path) if m == "oracle" or m.startswith("oracle.")]
        assert not bad, f"{path.relative_to(root)} imports the oracle: {bad}"
    # bench.py: only inside the two CPU legs
    src = (root / "bench.py").read_text()
    tree = ast.parse(src)
    # the baseline legs: the CPU reference (class CpuReference behind cpu_reference_rate / --impl reference) and the
    # reference-style eager path on the GPU (ref_gpu_path); none of them is on the engine's timed path
    allowed = {"cpu_reference_rate", "reference_arm", "__init__", "step", "ref_gpu_path"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            mod = node.module if isinstance(node, ast.ImportFrom) else None
            names = [a.name for a in node.names] if isinstance(node, ast.Import) else []
            if (mod and (mod == "oracle" or mod.startswith("oracle."))) or any(n == "oracle" or n.startswith("oracle.") for n in names):
                assert fn.name in allowed, f"bench.py:{node.lineno} imports the oracle inside {fn.name}()"
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    for node in top:
        mod = node.module if isinstance(node, ast.ImportFrom) else None
        names = [a.name for a in node.names] if isinstance(node, ast.Import) else []
        assert not (mod and mod.startswith("oracle")) and not any(n.startswith("oracle") for n in names), "bench.py imports the oracle at module level"


def test_bench_reference_arm_contract():
    """``bench.py --impl reference`` (the reference's CPU path on the host cores) prints ONE JSON line with the keys of the
    bench contract; run here on the toy model so that it takes seconds.  Under torchrun only rank 0 prints."""
    import json
    import os
    import pathlib
    import subprocess
    import sys
    root = pathlib.Path(__file__).resolve().parent.parent
    cmd = [sys.executable, str(root / "bench.py"), "--impl", "reference", "--model", "tiny", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "attack_iters_per_sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
