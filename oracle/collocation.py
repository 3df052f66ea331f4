"""Collocation constants — CPU restatement (TEST INFRASTRUCTURE, see oracle/README.md).

Follows reference do_mpc/optimizer.py:843-888: tau_root = [0] + collocation_points(deg, scheme),
Lagrange basis L_j through tau_root, D[j] = L_j(1), C[j][r] = dL_j/dtau (tau_root[r]).
The reference obtains the points from casadi.collocation_points (third-party, absent here);
they are the roots of the shifted Legendre polynomial P_d(2t-1) ('legendre') or of the shifted
Radau polynomial P_d(2t-1) - P_{d-1}(2t-1) ('radau', right end point included), which is the
published definition (Biegler 2010, ch. 10) and reproduces SURVEY.md App. C to 1e-15.
"""
import numpy as np
from numpy.polynomial import legendre as npleg
from numpy.polynomial import polynomial as nppoly


def collocation_points(deg, scheme="radau"):
    one = [0.0] * deg + [1.0]
    if scheme == "legendre":
        c = one
    elif scheme == "radau":
        c = np.array(one) - np.array([0.0] * (deg - 1) + [1.0, 0.0])
    else:
        raise ValueError("Unknown collocation scheme")
    roots = np.sort(npleg.legroots(c).real)
    return list((roots + 1.0) / 2.0)


def collocation_matrices(deg, scheme="radau"):
    """Return tau_root (deg+1), C ((deg+1)x(deg+1)), D (deg+1) exactly as optimizer.py:855-888."""
    tau = np.array([0.0] + collocation_points(deg, scheme))
    C = np.zeros((deg + 1, deg + 1))
    D = np.zeros(deg + 1)
    for j in range(deg + 1):
        p = np.array([1.0])
        for r in range(deg + 1):
            if r != j:
                p = nppoly.polymul(p, np.array([-tau[r], 1.0]) / (tau[j] - tau[r]))
        D[j] = nppoly.polyval(1.0, p)
        dp = nppoly.polyder(p)
        for r in range(deg + 1):
            C[j, r] = nppoly.polyval(tau[r], dp)
    return tau, C, D
