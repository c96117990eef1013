"""ORACLE (test infrastructure) -- import the *unmodified* reference modules in the build container.

/root/reference needs `cyipopt` (Ipopt + HSL MA57) and `matplotlib`; neither is installed and there is no
network.  This loader puts two stand-ins into sys.modules *before* importing the reference files:

  * matplotlib / matplotlib.pyplot : inert stubs (the reference only plots when plot=True).
  * cyipopt.Problem                : records (n, m, problem_obj, lb, ub, cl, cu, options) exactly as the
                                     reference passes them (MPC_LIP_sig_step.py:256-277, MPC_LIP_modi.py:274-295,
                                     MPC_DD_sig_step.py:171-191) and solves with SciPy SLSQP driven by the
                                     reference's own objective/gradient/constraints/jacobian callbacks
                                     (the fallback oracle named in SURVEY.md 8(c); "cyipopt unavailable").

It cannot travel to the GPU box (/root/reference does not exist there); it is used by oracle/gen_golden.py
to freeze fixtures under tests/golden/.  Nothing under the product package imports this file.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

LAST_PROBLEM = {}


class _StubProblem:
    def __init__(self, n, m, problem_obj=None, lb=None, ub=None, cl=None, cu=None):
        self.n, self.m, self.obj = n, m, problem_obj
        self.lb, self.ub = lb, ub
        self.cl = np.asarray(cl, dtype=np.float64)
        self.cu = np.asarray(cu, dtype=np.float64)
        self.options = {}

    def add_option(self, key, value):
        self.options[key] = value

    def solve(self, u0):
        from .slsqp_driver import solve_callbacks
        u0 = np.asarray(u0, dtype=np.float64).ravel()
        LAST_PROBLEM.clear()
        LAST_PROBLEM.update(n=self.n, m=self.m, obj=self.obj, lb=self.lb, ub=self.ub, cl=self.cl, cu=self.cu,
                            options=dict(self.options), u0=u0.copy(),
                            goal=np.ravel(np.asarray(self.obj.goal, dtype=np.float64)).copy())
        res = solve_callbacks(
            lambda u: float(self.obj.objective(u)),
            lambda u: np.ravel(np.asarray(self.obj.gradient(u), dtype=np.float64)),
            lambda u: np.ravel(np.asarray(self.obj.constraints(u), dtype=np.float64)),
            lambda u: np.asarray(self.obj.jacobian(u), dtype=np.float64).reshape(self.m, self.n),
            u0, self.cl, self.cu,
            None if self.lb is None else np.asarray(self.lb, dtype=np.float64),
            None if self.ub is None else np.asarray(self.ub, dtype=np.float64))
        LAST_PROBLEM.update(result=res)
        info = {"status": res["status"], "obj_val": res["f"], "x": res["u"]}
        return res["u"], info


def _install_stubs():
    if "cyipopt" not in sys.modules:
        cy = types.ModuleType("cyipopt")
        cy.Problem = _StubProblem
        sys.modules["cyipopt"] = cy
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None
        plt.__getattr__ = lambda name: _noop  # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load(name: str):
    """name in {'MPC_LIP_sig_step', 'MPC_LIP_modi', 'MPC_DD_sig_step'} -> module object."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    key = "_dcbf_ref_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, f"{REFERENCE_ROOT}/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
