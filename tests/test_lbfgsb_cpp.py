"""CPU gate for the host optimiser used by fgp_fit_batch (fungp_b200/csrc/lbfgsb.hpp): same bounded problems, same
R-style finite-difference gradient, against scipy's L-BFGS-B with R's optim() controls."""
import json
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin", "lbfgsb_test")


@pytest.fixture(scope="module")
def binary():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", BIN, os.path.join(ROOT, "tools", "lbfgsb_test.cpp")], check=True)
    return BIN


def _fun(prob, x):
    n = x.size
    i = np.arange(n)
    if prob == 0:
        return float(np.sum((i + 1) * (x - (0.5 * i - 0.3)) ** 2) + 0.3 * np.sum(x[:-1] * x[1:]))
    if prob == 1:
        return float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
    sgn = np.where(i % 2 == 1, 1.0, -0.4)
    return float(np.sum(np.log(x + 0.1) * sgn + 0.5 / (x + 0.05) + 0.02 * (i + 1) * x * x) + 0.1 * np.sin(3.0 * x[0] + x[-1]))


def _setup(prob, n):
    i = np.arange(n)
    if prob == 2:
        return np.full(n, 1e-10), 2.0 + 0.5 * i, np.ones(n)
    return np.full(n, -1.0), np.full(n, 1.5), np.where(i % 2 == 1, 0.2, -0.5)


@pytest.mark.parametrize("prob,n", [(0, 2), (0, 6), (0, 13), (1, 2), (1, 6), (2, 2), (2, 6), (2, 13)])
def test_cpp_lbfgsb_matches_scipy(binary, prob, n):
    out = json.loads(subprocess.run([binary, str(prob), str(n)], capture_output=True, text=True, check=True).stdout)
    lo, up, x0 = _setup(prob, n)

    def fg(x):
        g = np.empty(n)
        for k in range(n):
            a = x.copy(); b = x.copy()
            a[k] = min(x[k] + 1e-3, up[k]); b[k] = max(x[k] - 1e-3, lo[k])
            g[k] = (_fun(prob, a) - _fun(prob, b)) / (a[k] - b[k])
        return _fun(prob, x), g

    res = minimize(fg, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lo, up)),
                   options=dict(maxcor=5, ftol=1e7 * np.finfo(float).eps, gtol=0.0, maxiter=100))
    x = np.array(out["x"])
    assert out["conv"] in (0, 1)
    assert np.all(x >= lo) and np.all(x <= up)
    assert abs(out["f"] - _fun(prob, x)) < 1e-12 * max(1.0, abs(out["f"]))
    scale = max(1.0, abs(res.fun))
    assert out["f"] <= res.fun + 2e-6 * scale           # at least as good as scipy's optimum, to optimiser tolerance
    assert abs(out["f"] - res.fun) <= 2e-5 * scale
    if prob != 1:                                      # the Rosenbrock valley is flat: compare values only
        np.testing.assert_allclose(x, res.x, atol=5e-3)
