"""Builds ``libsimlingo_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch dependency: the library links against the CUDA runtime only (NCCL for the data-parallel exchange is resolved with
dlsym at run time, csrc/comm.cu); Python reaches it through
``ctypes`` (``simlingo_b200/lib.py``).  ``nvcc`` cross-compiles without a GPU."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsimlingo_b200.so"
SOURCES = ["api.cu", "gemm.cu", "patch_embed.cu", "gemv.cu", "decode.cu", "attention.cu", "attention_gqa.cu", "attention_vit.cu", "elementwise.cu", "backward.cu", "lora.cu", "attention_bwd.cu", "optim.cu", "comm.cu", "preprocess.cu", "postprocess.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "simlingo_b200.h"]
    stamp = HERE / ".build_stamp"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    flags = list(NVCC_FLAGS)
    procs = []
    objs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        objs.append(str(o))
        cmd = [nvcc, *flags, "-c", str(s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "shared", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
