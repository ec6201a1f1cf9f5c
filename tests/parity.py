"""Shared machinery of the parity tests: run the seeded cases of tests/golden/cases.py through a
backend (the CPU oracle or the CUDA library) and measure the distance to the reference's
golden outputs.

Error measures (stated here once, used everywhere):
* scalars (total):       |a - b| / max(|b|, tiny)
* integrands:            max_e |a_e - b_e| / max_e |b_e|
* gradients / fields:    max_v ||a_v - b_v||_2 / max_v ||b_v||_2      (per-component
                         relative error is meaningless where contributions cancel)
"""
from __future__ import annotations

import numpy as np

ANALYTIC = {"Length", "Area", "VolumeEnclosed", "Volume"}

# Tolerances.  Energies, integrands and analytic gradients: 1e-12 (BASELINE.json north_star).
# Finite-difference gradients (every gradient/fieldgradient of the other functionals is a central
# difference in the reference): the reference amplifies its own last-bit integrand
# round-off by 1/(2h) ~ 8e4, so two correct implementations of the same stencil agree to ~1e-10
# of the gradient scale at best; the bound is written per call site.
TOL_EXACT = 1e-12
TOL_FD = 2e-9


def rel_scalar(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def rel_array(a, b, rows=None):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    if rows:
        a, b = a.reshape(-1, rows), b.reshape(-1, rows)
        num = np.linalg.norm(a - b, axis=1).max() if a.size else 0.0
        den = np.linalg.norm(b, axis=1).max() if b.size else 0.0
    else:
        num = np.abs(a - b).max() if a.size else 0.0
        den = np.abs(b).max() if b.size else 0.0
    if den == 0.0:
        return 0.0 if num == 0.0 else np.inf
    return num / den


def tolerance(ev):
    name, method = ev["f"]["name"], ev["method"]
    if method in ("total", "integrand"):
        return TOL_EXACT
    if method == "gradient" and name in ANALYTIC:
        return TOL_EXACT
    return TOL_FD


def measure(ev, got, want, case):
    method = ev["method"]
    if method == "total":
        return rel_scalar(float(np.asarray(got).ravel()[0]), float(want[0]))
    if method == "integrand":
        return rel_array(got, want)
    if method == "gradient":
        return rel_array(got, want, rows=case["vert"].shape[1])
    fl = np.asarray(case["fields"][ev["wrt"]])
    return rel_array(got, want, rows=(1 if fl.ndim == 1 else fl.shape[1]))
