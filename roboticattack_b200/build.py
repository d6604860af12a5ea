"""Build recipe for the sm_100a C-ABI shared library (``libvla_b200.so``).

nvcc cross-compiles without a GPU; the built ``.so`` lives in-tree (git-ignored) so that it travels to the
GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libvla_b200.so"
STAMP = PKG_DIR / ".libvla_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
    "-shared",
] + os.environ.get("VLA_NVCC_EXTRA", "").split()


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                   + list((PKG_DIR.parent / "include").glob("*.h")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_extension(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` into ``libvla_b200.so`` (skipped when sources are unchanged)."""
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] nvcc failed for {src.name}:\n{out}\n")
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", str(LIB_PATH)] + [str(o) for o in objs] + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    subprocess.run(cmd, check=True)
    STAMP.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    p = build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
