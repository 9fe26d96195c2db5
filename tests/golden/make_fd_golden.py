"""Generate tests/golden/fd_reference.json by IMPORTING the reference's own
python/Main_finite_difference.py (build container only; /root/reference does not exist on
the GPU box) and running its functions refineAll / interpU / forwardSolve / adjSolve / errEst
on the cases of its __main__ block (u' = sin u, J = int u^2, ref_factor = 4) -- SURVEY App. B.3.

    python tests/golden/make_fd_golden.py [/root/reference]
"""
import importlib.util
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np


def load_reference(ref):
    for m in ("cv2", "matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location("ref_fd", os.path.join(ref, "python", "Main_finite_difference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # the driver loop is under __main__ and does not run
    return mod


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    fd = load_reference(ref)
    ref_factor = 4
    fd.ref_factor = ref_factor        # interpU reads a module global (quirk C-12)

    # problem functions: verbatim semantics of the reference's __main__ block
    # (Main_finite_difference.py:110-140 ODEs, :153-227 functionals), restated as closures
    def make(ode, functional):
        if ode == "sin":
            fwdUpdate = lambda u, dt, n: u[n - 1] + np.sin(u[n - 1]) * dt[n - 1]
            getJF = lambda u, dt: np.diag(1 + np.cos(u[:-1]) * dt, -1)
        else:
            fwdUpdate = lambda u, dt, n: (1 + dt[n - 1]) * u[n - 1]
            getJF = lambda u, dt: np.diag(1 + dt, -1)
        if functional == "int_u2":
            getK = lambda dt, u: np.concatenate((2 * u[:-1] * dt, 0), axis=None)
        elif functional == "int_u":
            getK = lambda dt, u=None: np.concatenate((dt, 0), axis=None)
        else:
            def getK(dt, u=None):
                k = np.zeros_like(dt)
                k[-1] = 1
                return np.concatenate((k, 0), axis=None)
        return fwdUpdate, getJF, getK

    cases = []
    rng = np.random.default_rng(7)
    meshes = [np.array([0.0, 1.0, 2.0]), np.array([0.0, 0.5, 1.0, 2.0]), np.array([0.0, 0.25, 0.5, 1.0, 2.0]),
              np.sort(np.concatenate(([0.0, 2.0], rng.uniform(0.05, 1.95, 14))))]
    combos = [("sin", "int_u2", 4)] * 1 + [("sin", "int_u", 4), ("sin", "u_N", 4), ("linear", "int_u", 4),
                                            ("linear", "u_N", 3), ("linear", "int_u2", 5)]
    for ode, functional, ref_factor in combos:
        fd.ref_factor = ref_factor
        fwdUpdate, getJF, getK = make(ode, functional)
        for times in meshes:
            for u0 in ((1.0, -2.3, 0.4) if (ode, functional) == ("sin", "int_u2") else (1.0,)):
                dt_n = np.diff(times)
                u = fd.forwardSolve(fwdUpdate, dt_n, u0)
                v = fd.adjSolve(getK, getJF, dt_n, u, ref_factor)
                err_fine = fd.errEst(fwdUpdate, u, v, dt_n, ref_factor)
                e = np.abs(err_fine)[2:]
                # window sum of the driver loop (Main_finite_difference.py:270-277)
                n = len(dt_n)
                err_steps = np.array([e[r * ref_factor:r * ref_factor + ref_factor - 1].sum() for r in range(n)])
                cases.append(dict(ode=ode, functional=functional, times=times.tolist(), u0=u0, ref_factor=ref_factor,
                                  u=u.tolist(), v=v.tolist(), err_fine=err_fine.tolist(),
                                  err_steps=err_steps.tolist(), ref_idx=int(np.argmax(err_steps))))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fd_reference.json")
    with open(dst, "w") as f:
        json.dump(dict(source="python/Main_finite_difference.py functions, imported unmodified",
                       cases=cases), f, indent=0)
    print("wrote", dst, len(cases), "cases")
    c = cases[0]
    print(c["u"], c["v"][0], c["err_steps"])


if __name__ == "__main__":
    main()
