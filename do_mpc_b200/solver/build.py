"""Build a problem library: generated header + csrc/dmb_api.cu -> sm_100a shared object (in-tree cache).

nvcc cross-compiles without a GPU, so MPC.setup() works on CPU-only hosts up to (and including)
the build; only creating a solver handle needs a B200.  Libraries are cached by the hash of the generated
source + kernel sources under do_mpc_b200/_kcache/ so that they travel with the repository
snapshot (they are git-ignored build products).
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
from pathlib import Path

from ..recipe import Recipe
from .codegen import generate

PKG = Path(__file__).resolve().parent.parent
CSRC = Path(os.environ.get("DMB_CSRC_DIR", PKG / "csrc"))              # override: A/B experiments against an older kernel source
INCLUDE = Path(os.environ.get("DMB_INCLUDE_DIR", PKG.parent / "include"))
CACHE = PKG / "_kcache"
NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _kernel_sources_digest() -> str:
    h = hashlib.sha256()
    for f in ("ipm_core.cuh", "ipm_solve.cuh", "ipm_refine.cuh", "ipm_sens.cuh", "dmb_api.cu", "dmb_host_util.h"):
        h.update((CSRC / f).read_bytes())
    h.update((INCLUDE / "dompc_b200.h").read_bytes())
    return h.hexdigest()


def _sizes(r: Recipe, src: str, threads: int):
    """Mirror of the layout macros in csrc/ipm_core.cuh / dmb_api.cu (in doubles)."""
    nnz_h = int(re.search(r"#define DMB_NNZ_H (\d+)", src).group(1))
    nnz_d = int(re.search(r"#define DMB_NNZ_D (\d+)", src).group(1))
    nx, nu, neps, ny = r.n_x, r.n_u, r.n_eps, r.n_y
    nnl = r.nl_rows_per_edge
    nvp = nx + nu + neps; nq = nvp + ny; ne = nvp + 1
    nxi, nut = nx + nu, nu + neps
    nv = nxi + nut
    n_node, n_edge = len(r.node_k), len(r.edge_k)
    gsz = 16 if ny <= 16 else 32
    tile = ny * ((ny + 2 * ne) | 1) + nvp * ne + nnz_h + nnz_d + nq + 4 * ny + 2 * nnl
    tiles = (threads // gsz) * tile
    qs = (threads // 32 if r.n_comb ** r.n_robust > 1 else 1) * (nxi * nvp + nxi + 8)      # Riccati scratch (per warp for scenario trees)
    ern = n_edge * (nx * nvp + nx + 2 * nvp * nvp + 2 * nvp) + n_node * (nv * nv + nv + nut * nxi + nut + nxi + 2 * nu + 4 * nu * nu)
    nwi = n_node * nvp + n_edge * ny
    nlam = nx + n_edge * (ny + nx + nnl)
    vec = 7 * nwi + 7 * nlam + 5 * n_edge * nnl
    return tiles + qs + 128, ern, vec


def _config_for(r: Recipe, src: str, threads: int):
    fixed, ern, vec = _sizes(r, src, threads)
    kb = lambda d: 8 * d / 1024.0
    ern_smem = int(kb(fixed + ern) <= 110)
    vec_smem = int(ern_smem and kb(fixed + ern + vec) <= 73)
    if "DMB_FORCE_VEC_SMEM" in os.environ:                 # tuning experiments only
        vec_smem = int(os.environ["DMB_FORCE_VEC_SMEM"])
    total = kb(fixed + (ern if ern_smem else 0) + (vec if vec_smem else 0)) + 1
    return fixed, ern_smem, vec_smem, total


def launch_config(r: Recipe, src: str, threads: int = None):
    """(threads per CTA, reduced blocks + node data in shared memory?, iterate vectors in shared memory?, CTAs/SM the
    register budget is compiled for).  227 KB of shared memory per SM on sm_100.  Measured on B200 (cfg #2):
      * 3 CTAs/SM with the iterate vectors in the L1/L2-backed slot beat 2 CTAs/SM with everything on chip (132.6k vs 110.6k),
      * the factor and recovery phases process the edges in rounds of (threads / GSZ) lane groups: 160 threads = 10 groups
        finish the 20 edges of cfg #2 in 2 full rounds instead of 3 (8 + 8 + 4) with 128 threads: 137.8k vs 127.5k solves/s
        at the same 3 CTAs/SM (registers capped at 128 instead of 168; the extra spills cost less than the idle round).
    Rule: 128 threads unless the edges fit in few rounds (<= 6 with 128 threads); then the warp-multiple CTA size with the
    fewest rounds among those that keep the CTAs/SM of the 128-thread build (shared memory and a 128-register budget).
    Problems with many edges (scenario trees: cfg #3 270 edges, cfg #5 180) measured faster with 128 threads and more
    CTAs (3301 vs 2794, 1248 vs 1216 solves/s) — the round granularity does not matter there."""
    n_edge = len(r.edge_k)
    gsz = 16 if r.n_y <= 16 else 32

    def cfg(threads):
        fixed, ern_smem, vec_smem, total = _config_for(r, src, threads)
        if 8 * fixed > 220 * 1024 or total > 226:
            return None
        ctas = max(1, min(4, int(227 // total), 65536 // (threads * 128)))
        return (-(-n_edge // (threads // gsz)), threads, ern_smem, vec_smem, ctas)

    if threads is not None:                                # explicit CTA size (latency build, see BatchedNlpSolver)
        best = cfg(int(threads))
    elif "DMB_FORCE_THREADS" in os.environ:                # tuning experiments only
        best = cfg(int(os.environ["DMB_FORCE_THREADS"]))
    else:
        best = cfg(128) or cfg(64)
        if best is not None and best[1] == 128 and best[0] <= 6:
            for threads in (160, 192, 224, 256):
                c = cfg(threads)
                if c is not None and c[4] >= min(best[4], 3) and c[0] < best[0]:
                    best = c
    if best is None:
        raise NotImplementedError("problem too large for the shared-memory tiles of the solver")
    _, threads, ern_smem, vec_smem, min_ctas = best
    if "DMB_FORCE_MIN_CTAS" in os.environ:
        min_ctas = int(os.environ["DMB_FORCE_MIN_CTAS"])
    return threads, ern_smem, vec_smem, min_ctas


MAX_LATENCY_THREADS = 256      # 8 warps: the CTA-wide reductions keep 7 values x warps in DmbShared::red[64] (csrc/ipm_solve.cuh)


def latency_threads(r: Recipe, src: str):
    """CTA size of the *latency build* of a problem, or None if a larger CTA would not shorten the edge loops.

    The throughput rule above sizes a CTA so that 3-4 instances share an SM (registers capped at 128); a single instance (the
    reference's real-time `make_step(x0)`) then runs on one CTA of 128-160 threads and walks through its edges in many rounds of
    lane groups.  The latency build gives one instance more of an SM: up to 256 threads, compiled for one CTA per SM without a
    register cap; it is used for small batches (BatchedNlpSolver.latency_max_batch).  Measured on B200, cold start, B = 1
    (profiles/r2_latency_threads.json, 256 threads): robust CSTR (180 edges) 34.3 -> 18.8 ms, batch reactor tree N=30 (270 edges)
    85.5 -> 45.4 ms (30.3 ms with 512 threads, which the reduction scratch does not support yet), nominal CSTR (20 edges, 2 rounds
    either way) 2.10 -> 1.82 ms; identical iteration counts.  Rule: the CTA size <= 256 with the fewest rounds (the smallest
    among equals), if that saves at least a quarter of the rounds of the throughput build."""
    gsz = 16 if r.n_y <= 16 else 32
    base = launch_config(r, src)[0]
    n_edge = len(r.edge_k)
    rounds = lambda t: -(-n_edge // (t // gsz))
    best = None
    for threads in range(base + 32, MAX_LATENCY_THREADS + 1, 32):
        fixed, ern_smem, vec_smem, total = _config_for(r, src, threads)
        if 8 * fixed > 220 * 1024 or total > 226:
            continue
        if best is None or rounds(threads) < rounds(best):
            best = threads
    if best is None or 4 * rounds(best) > 3 * rounds(base):
        return None
    return best


def build_library(r: Recipe, verbose: bool = False, src: str = None, threads: int = None) -> Path:
    """Generated problem library of a stage recipe (in-tree cache).  `src`: the already generated problem_gen.h (lets a caller
    generate serially — the symbolic engine is not thread safe — and compile several libraries concurrently).  `threads`: CTA
    size of the latency build (one CTA per SM, see latency_threads); None = the throughput rule."""
    src = generate(r) if src is None else src
    latency = threads is not None
    threads, ern_smem, vec_smem, min_ctas = launch_config(r, src, threads=threads)
    if latency:
        min_ctas = 1                                        # one CTA per SM: no register cap
    extra = os.environ.get("DMB_EXTRA_NVCC_FLAGS", "").split()          # tuning experiments only
    key = hashlib.sha256((src + _kernel_sources_digest() + f"{threads}{ern_smem}{vec_smem}{min_ctas}{extra}").encode()).hexdigest()[:16]
    d = CACHE / key
    lib = d / f"libdmb_{key}.so"
    if lib.exists():
        return lib
    d.mkdir(parents=True, exist_ok=True)
    (d / "problem_gen.h").write_text(src)
    cmd = [NVCC, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           f"-DDMB_THREADS={threads}", f"-DDMB_ERN_SMEM={ern_smem}", f"-DDMB_VEC_SMEM={vec_smem}", f"-DDMB_MIN_CTAS={min_ctas}", *extra, f"-I{d}", f"-I{CSRC}", f"-I{INCLUDE}",
           str(CSRC / "dmb_api.cu"), "-o", str(lib) + f".tmp{os.getpid()}"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for the generated problem library:\n" + " ".join(cmd) + "\n" + res.stderr[-6000:])
    if verbose:
        print(res.stderr)
    os.replace(str(lib) + f".tmp{os.getpid()}", lib)        # atomic: concurrent ranks may build the same library
    (d / "build_cmd.txt").write_text(" ".join(cmd) + "\n")
    return lib


def build_sim_library(model, deg: int = 3) -> Path:
    """Plant-simulator library for a model (csrc/sim_api.cu + generated sim_gen.h), cached like the solver libraries."""
    from .sim_codegen import generate_sim
    src = generate_sim(model, deg)
    h = hashlib.sha256(src.encode())
    for f in ("sim_core.cuh", "sim_api.cu"):
        h.update((CSRC / f).read_bytes())
    h.update((INCLUDE / "dompc_b200_sim.h").read_bytes())
    key = h.hexdigest()[:16]
    d = CACHE / f"sim_{key}"
    lib = d / f"libdms_{key}.so"
    if lib.exists():
        return lib
    d.mkdir(parents=True, exist_ok=True)
    (d / "sim_gen.h").write_text(src)
    tmp = str(lib) + f".tmp{os.getpid()}"
    cmd = [NVCC, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", f"-I{d}", f"-I{CSRC}",
           f"-I{INCLUDE}", str(CSRC / "sim_api.cu"), "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for the generated simulator library:\n" + " ".join(cmd) + "\n" + res.stderr[-6000:])
    os.replace(tmp, lib)
    return lib
