"""Build a problem library: generated header + csrc/dmb_api.cu -> sm_100a shared object (in-tree cache).

nvcc cross-compiles without a GPU, so MPC.setup() works on CPU-only hosts up to (and including)
the build; only creating a solver handle needs a B200.  Libraries are cached by the hash of the generated
source + kernel sources under do_mpc_b200/_kcache/ so that they travel with the repository
snapshot (they are git-ignored build products).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

from ..recipe import Recipe
from .codegen import generate

PKG = Path(__file__).resolve().parent.parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
CACHE = PKG / "_kcache"
NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _kernel_sources_digest() -> str:
    h = hashlib.sha256()
    for f in ("ipm_core.cuh", "ipm_solve.cuh", "dmb_api.cu", "dmb_host_util.h"):
        h.update((CSRC / f).read_bytes())
    h.update((INCLUDE / "dompc_b200.h").read_bytes())
    return h.hexdigest()


def launch_config(r: Recipe):
    """Threads per CTA and whether the iterate vectors fit in shared memory (227 KB per CTA on sm_100)."""
    n_vp = r.n_x + r.n_u + r.n_eps
    n_q = n_vp + r.n_y
    n_node, n_edge = len(r.node_k), len(r.edge_k)
    nwi = n_node * n_vp + n_edge * r.n_y
    nlam = r.n_x + n_edge * (r.n_y + r.n_x + r.nl_rows_per_edge)
    wv = 6 * nwi + 5 * nlam + 4 * n_edge * r.nl_rows_per_edge
    threads = 128
    tile = (threads // 32) * (r.n_y * ((r.n_y + 2 * n_vp + 1) | 1) + n_q * (n_q | 1))
    nxi, nv = r.n_x + r.n_u, r.n_x + 2 * r.n_u + r.n_eps
    qs = nv * nv + nv + 2 * nxi * nv
    fixed = 8 * (tile + qs) + 2048
    vec_smem = int(fixed + 8 * wv <= 100 * 1024)       # keep >= 2 CTAs per SM
    if fixed > 200 * 1024:
        threads = 64
    return threads, vec_smem


def build_library(r: Recipe, verbose: bool = False) -> Path:
    src = generate(r)
    threads, vec_smem = launch_config(r)
    key = hashlib.sha256((src + _kernel_sources_digest() + f"{threads}{vec_smem}").encode()).hexdigest()[:16]
    d = CACHE / key
    lib = d / f"libdmb_{key}.so"
    if lib.exists():
        return lib
    d.mkdir(parents=True, exist_ok=True)
    (d / "problem_gen.h").write_text(src)
    cmd = [NVCC, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           f"-DDMB_THREADS={threads}", f"-DDMB_VEC_SMEM={vec_smem}", f"-I{d}", f"-I{CSRC}", f"-I{INCLUDE}",
           str(CSRC / "dmb_api.cu"), "-o", str(lib) + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for the generated problem library:\n" + " ".join(cmd) + "\n" + res.stderr[-6000:])
    if verbose:
        print(res.stderr)
    os.replace(str(lib) + ".tmp", lib)
    (d / "build_cmd.txt").write_text(" ".join(cmd) + "\n")
    return lib
