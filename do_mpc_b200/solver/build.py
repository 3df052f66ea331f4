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


def smem_doubles(r: Recipe, threads: int):
    """Mirror of the shared-memory carve-up in csrc/dmb_api.cu / ipm_core.cuh (doubles): (fixed part, iterate vectors)."""
    n_vp = r.n_x + r.n_u + r.n_eps
    n_e = n_vp + 1
    n_node, n_edge = len(r.node_k), len(r.edge_k)
    nnl = r.nl_rows_per_edge
    nwi = n_node * n_vp + n_edge * r.n_y
    nlam = r.n_x + n_edge * (r.n_y + r.n_x + nnl)
    gsz = 16 if r.n_y <= 16 else 32
    tile = (threads // gsz) * (r.n_y + n_vp) * ((r.n_y + 2 * n_e) | 1)
    nxi, nut = r.n_x + r.n_u, r.n_u + r.n_eps
    nv = nxi + nut
    qs = nv * nv + nv + nxi * nv + 8
    er = n_edge * (r.n_x * n_vp + r.n_x + 2 * n_vp * n_vp + 2 * n_vp)
    node = n_node * (nxi * nxi + nxi + nut * nxi + nut + nxi)
    vec = 7 * nwi + 5 * nlam + 4 * n_edge * nnl
    return tile + qs + er + node + 64, vec


def launch_config(r: Recipe):
    """Threads per CTA, whether the iterate vectors live in shared memory, and the CTAs/SM the register budget is
    compiled for (227 KB of shared memory per SM on sm_100)."""
    threads = 128
    fixed, vec = smem_doubles(r, threads)
    if 8 * fixed > 200 * 1024:
        threads = 64
        fixed, vec = smem_doubles(r, threads)
    if 8 * fixed > 220 * 1024:
        raise NotImplementedError(f"problem too large for the shared-memory resident solver ({8 * fixed} bytes per CTA)")
    vec_smem = int(8 * (fixed + vec) <= 113 * 1024)
    if "DMB_FORCE_VEC_SMEM" in os.environ:                 # tuning experiments only
        vec_smem = int(os.environ["DMB_FORCE_VEC_SMEM"])
    total = 8 * (fixed + (vec if vec_smem else 0)) + 1024
    min_ctas = max(1, min(4, (227 * 1024) // total))
    if "DMB_FORCE_MIN_CTAS" in os.environ:
        min_ctas = int(os.environ["DMB_FORCE_MIN_CTAS"])
    return threads, vec_smem, min_ctas


def build_library(r: Recipe, verbose: bool = False) -> Path:
    src = generate(r)
    threads, vec_smem, min_ctas = launch_config(r)
    key = hashlib.sha256((src + _kernel_sources_digest() + f"{threads}{vec_smem}{min_ctas}").encode()).hexdigest()[:16]
    d = CACHE / key
    lib = d / f"libdmb_{key}.so"
    if lib.exists():
        return lib
    d.mkdir(parents=True, exist_ok=True)
    (d / "problem_gen.h").write_text(src)
    cmd = [NVCC, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           f"-DDMB_THREADS={threads}", f"-DDMB_VEC_SMEM={vec_smem}", f"-DDMB_MIN_CTAS={min_ctas}", f"-I{d}", f"-I{CSRC}", f"-I{INCLUDE}",
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
