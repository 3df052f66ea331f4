"""Host logic: the symbolic engine that replaces CasADi SX for the stage atoms (expression DAG, AD, C emission,
structures with power indexing).  Derivatives are checked against sympy and finite differences."""
import math
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest
import sympy as sy

from do_mpc_b200.casadi_api import *  # noqa: F401,F403
from do_mpc_b200.casadi_api import SX, DM, Function, jacobian, hessian, vertcat, exp, sin, cos, sqrt, sum1, substitute
from do_mpc_b200.sym import expr as E
from do_mpc_b200.sym.struct import SymStruct, entry


def test_arithmetic_and_function_evaluation():
    x = SX.sym("x", 3); u = SX.sym("u")
    f = Function("f", [x, u], [vertcat(x[0] * exp(-x[1]) + u ** 2, sin(x[2]) / (1 + x[0] ** 2), sqrt(x[1] + 2) * cos(u))])
    v = f(np.array([0.3, 1.2, -0.7]), 0.5).full().ravel()
    ref = [0.3 * math.exp(-1.2) + 0.25, math.sin(-0.7) / (1 + 0.09), math.sqrt(3.2) * math.cos(0.5)]
    assert np.allclose(v, ref, rtol=1e-15)
    A = np.array([[1.0, 2.0, 0.0], [0.0, -1.0, 3.0]])
    g = Function("g", [x], [A @ x])                       # numpy matrix @ SX must defer to SX.__rmatmul__
    assert np.allclose(g(np.array([1.0, 2.0, 3.0])).full().ravel(), A @ [1, 2, 3])


def test_jacobian_and_hessian_against_sympy():
    xs = sy.symbols("a b c")
    expr_sy = xs[0] ** 2 * sy.exp(-xs[1] / (xs[2] + 3)) + sy.sin(xs[0] * xs[2]) - sy.sqrt(xs[1] + 1) / xs[0]
    x = SX.sym("x", 3)
    ex = x[0] ** 2 * exp(-x[1] / (x[2] + 3)) + sin(x[0] * x[2]) - sqrt(x[1] + 1) / x[0]
    H, g = hessian(ex, x)
    fg, fH = Function("g", [x], [g]), Function("H", [x], [H])
    pt = np.array([0.7, 1.3, -0.4])
    sub = dict(zip(xs, pt))
    g_ref = [float(sy.diff(expr_sy, v).subs(sub)) for v in xs]
    H_ref = [[float(sy.diff(expr_sy, a, b).subs(sub)) for b in xs] for a in xs]
    assert np.allclose(fg(pt).full().ravel(), g_ref, rtol=1e-12)
    assert np.allclose(fH(pt).full(), H_ref, rtol=1e-11, atol=1e-13)
    J = jacobian(vertcat(ex, x[0] * x[1]), x)
    assert J.shape == (2, 3)


def test_common_subexpressions_are_shared_and_constants_fold():
    x = SX.sym("x")
    a = exp(x) * exp(x) + exp(x)
    assert E.count_ops(a.ids()) <= 3                      # exp, sq, add
    assert float(SX(2.0) * 3 + 1) == 7.0
    z = x * 0 + 0 * x
    assert z.is_constant() and float(z) == 0.0
    assert substitute(x + 1, x, SX(2.0)).is_constant()


def test_emitted_c_matches_python_evaluation():
    x = SX.sym("x", 2)
    ex = vertcat(exp(x[0]) * x[1] ** 3 - x[0] / (x[1] ** 2 + 1), cos(x[0]) ** 2)
    code = E.emit_c([(f"out[{i}]", e) for i, e in enumerate(ex.ids())], {s: f"x[{i}]" for i, s in enumerate(x.ids())})
    src = f"#include <math.h>\n#include <stdio.h>\nint main(){{double x[2]={{0.3,-1.7}};double out[2];\n{code}\nprintf(\"%.17g %.17g\\n\",out[0],out[1]);return 0;}}\n"
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-O0", str(Path(d) / "t.c"), "-o", str(Path(d) / "t"), "-lm"], check=True)
        out = subprocess.run([str(Path(d) / "t")], capture_output=True, text=True, check=True).stdout.split()
    ref = E.evaluate(ex.ids(), dict(zip(x.ids(), [0.3, -1.7])))
    assert np.allclose([float(v) for v in out], ref, rtol=1e-15)


def test_struct_power_indexing_like_casadi_tools():
    inner = SymStruct([entry("a"), entry("th", shape=(2, 1))])
    s = SymStruct([entry("_x", repeat=[3, 2], struct=inner), entry("_u", repeat=[2], shape=(1, 1))])
    assert s.size == 3 * 2 * 3 + 2
    n = s(0.0)
    n["_x", 1, 0, "th"] = np.array([5.0, 6.0])
    n["_x", 2, :, "a"] = 9.0
    n["_u"] = 2.5
    v = n.full().ravel()
    assert v[(1 * 2 + 0) * 3 + 1] == 5.0 and v[(1 * 2 + 0) * 3 + 2] == 6.0
    assert v[(2 * 2 + 0) * 3] == 9.0 and v[(2 * 2 + 1) * 3] == 9.0
    assert np.all(v[-2:] == 2.5)
    assert n["_x", 1, 0, "th", 1].full().item() == 6.0 and n["_x", -1, -1, "a"].full().item() == 9.0
    assert len(n["_x", :, 0]) == 3 and n["_x", 0, 0].shape == (3, 1)
    assert s.labels()[0] == "[_x,0,0,a,0]" and s.keys() == ["_x", "_u"]
    with pytest.raises(KeyError):
        n["nope"]


def test_model_and_mpc_validation_errors():
    import do_mpc_b200 as do_mpc
    with pytest.raises(AssertionError):
        do_mpc.model.Model("continous")
    m = do_mpc.model.Model("continuous")
    x = m.set_variable("_x", "x"); m.set_variable("_u", "u")
    with pytest.raises(Exception):
        m.setup()                                              # rhs missing
    m.set_rhs("x", -x)
    m.setup()
    mpc = do_mpc.controller.MPC(m)
    mpc.settings.n_horizon, mpc.settings.t_step = 5, 0.1
    with pytest.raises(Exception, match="Objective is undefined"):
        mpc.prepare_nlp()
    mpc.set_objective(mterm=x ** 2, lterm=x ** 2)
    mpc.bounds["lower", "_x", "x"] = 1.0; mpc.bounds["upper", "_x", "x"] = -1.0
    with pytest.raises(Exception, match="bounds are inconsistent"):
        mpc.prepare_nlp()
