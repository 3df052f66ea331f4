"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib_path():
    import cstr
    from do_mpc_b200.solver.build import build_library
    mpc = cstr.cstr_nominal_mpc(setup=False)
    mpc.prepare_nlp()
    return build_library(mpc.recipe)


def test_exports_every_declared_symbol(lib_path):
    hdr = (ROOT / "include" / "dompc_b200.h").read_text()
    names = set(re.findall(r"\b(dmb_[a-z_]+)\s*\(", hdr))
    assert {"dmb_create", "dmb_destroy", "dmb_solve_host", "dmb_solve_device", "dmb_set_bounds", "dmb_set_options",
            "dmb_get_dims", "dmb_default_options", "dmb_aux_host", "dmb_get_stats", "dmb_last_error", "dmb_abi_version"} <= names
    lib = ctypes.CDLL(str(lib_path))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dompc_b200.h but not exported"
    assert lib.dmb_abi_version() == 1


def test_dims_and_option_struct(lib_path):
    from do_mpc_b200.solver import load_library, DIM_COUNT, DIM_NW, DIM_NG, DIM_NP, DIM_OPTSIZE
    from do_mpc_b200.solver.options import DmbOptions
    lib = load_library(lib_path)
    dims = (ctypes.c_int * DIM_COUNT)()
    lib.dmb_get_dims(dims)
    assert (dims[DIM_NW], dims[DIM_NG], dims[DIM_NP]) == (376, 324, 6)
    assert dims[DIM_OPTSIZE] == ctypes.sizeof(DmbOptions)
    o = DmbOptions()
    lib.dmb_default_options(ctypes.byref(o))
    assert o.tol == 1e-8 and o.max_iter == 3000 and o.mu_init == 0.1 and o.kappa_w_plus == 8.0


def test_image_is_sm100a_and_has_the_kernel(lib_path):
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(lib_path)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure_not_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cstr
    mpc = cstr.cstr_nominal_mpc(setup=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mpc.setup()


def test_simulator_library_exports_every_declared_symbol():
    import cstr
    from do_mpc_b200.solver.build import build_sim_library
    hdr = (ROOT / "include" / "dompc_b200_sim.h").read_text()
    names = set(re.findall(r"\b(dms_[a-z_]+)\s*\(", hdr))
    assert {"dms_create", "dms_destroy", "dms_step_host", "dms_step_device", "dms_get_dims", "dms_last_error"} <= names
    lib = ctypes.CDLL(str(build_sim_library(cstr.cstr_model(False))))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dompc_b200_sim.h but not exported"
    d = (ctypes.c_int * 8)()
    lib.dms_get_dims(d)
    assert list(d)[:3] == [4, 2, 0]
    lib.dms_create.restype = ctypes.c_void_p
    assert not lib.dms_create(0)          # no GPU here: loud failure, no fallback
