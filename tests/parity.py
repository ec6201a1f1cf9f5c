"""Shared machinery of the parity tests: run the seeded cases of tests/golden/cases.py through a
backend (the CPU oracle or the CUDA library) and measure the distance to the reference's
golden outputs.

Error measures (stated here once, used everywhere):
* scalars (total):       |a - b| / max(|b|, tiny)
* integrands:            max_e |a_e - b_e| / max_e |b_e|
* gradients / fields:    max_v ||a_v - b_v||_2 / max_v ||b_v||_2      (per-component
                         relative error is meaningless where contributions cancel)
* analytic gradients additionally, PER VERTEX (H4's first option):
                         max_v ||a_v - b_v|| / max(||b_v||, EPS_FLOOR * S_v),   S_v = sum_e ||contribution_e(v)||
                         — a wrong row at a vertex whose net force is small is visible here; a vertex whose net force is
                         below EPS_FLOOR of what its elements contribute (cancellation) is measured against that floor.

Achieved errors are recorded by `record()` into gpurun_out/parity_errors.json (and printed) so that the tolerances of
the finite-difference family can be set from what was observed: FD_TOL holds, per (functional, method), 10x the worst
error observed on the B200 over all golden cases (profiles/r2_parity_errors.json is the committed copy).
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
TOL_FD = 2e-9          # ceiling for any finite-difference map that has no entry in FD_TOL yet
EPS_FLOOR = 1e-2       # per-vertex measure: floor as a fraction of the vertex's summed contribution norms

# 10x the worst error observed on the B200 (tests/test_parity_gpu.py::test_cuda_matches_reference_golden over all
# golden cases; observed values: profiles/r2_parity_errors.json).  Keys: (functional, method).
FD_TOL = {
    ("AreaEnclosed", "gradient"): 2.6e-11,            # observed 2.6e-12
    ("EquiElement", "gradient"): 1.0e-11,             # observed 6.1e-15 over the 60 evaluations of the tactoid example
    ("GradSq", "fieldgradient"): 5.0e-10,             # observed 5.0e-11
    ("GradSq", "gradient"): 1.9e-10,                  # observed 1.9e-11
    ("LineCurvatureSq", "gradient"): 8.8e-11,         # observed 8.7e-12
    ("Nematic", "fieldgradient"): 8.5e-11,            # observed 8.5e-12
    ("Nematic", "gradient"): 1.8e-10,                 # observed 1.8e-11
    ("NematicElectric", "fieldgradient"): 5.2e-11,    # observed 5.2e-12
    ("NematicElectric", "gradient"): 1.7e-11,         # observed 1.7e-12
    ("NormSq", "fieldgradient"): 3.4e-10,             # observed 3.3e-11
    ("NormSq", "gradient"): 1e-12,                    # observed 0 (identically zero in the reference too)
}


_recorded = {}


def record(group, key, err):
    """Remember the worst error per (group, key); flushed to gpurun_out/parity_errors.json by tests/conftest.py."""
    d = _recorded.setdefault(group, {})
    d[key] = max(d.get(key, 0.0), float(err))


def recorded():
    return _recorded


def contribution_norm_sum(name, vert, conn):
    """S_v = sum over incident elements of the norm of that element's gradient contribution at v, in closed form:
    Length 1; Area half the opposite edge; VolumeEnclosed |x_j x x_k| / 6; Volume a third of the opposite face."""
    vert = np.asarray(vert, dtype=np.float64)
    conn = np.asarray(conn)
    nv, k = vert.shape[0], conn.shape[1]
    S = np.zeros(nv)
    x = [vert[conn[:, a]] for a in range(k)]
    if vert.shape[1] == 2:
        x = [np.concatenate([xa, np.zeros((xa.shape[0], 1))], axis=1) for xa in x]
    for a in range(k):
        o = [b for b in range(k) if b != a]
        if name == "Length":
            w = np.ones(conn.shape[0])
        elif name == "Area":
            w = 0.5 * np.linalg.norm(x[o[0]] - x[o[1]], axis=1)
        elif name == "VolumeEnclosed":
            w = np.linalg.norm(np.cross(x[o[0]], x[o[1]]), axis=1) / 6.0
        elif name == "Volume":
            w = np.linalg.norm(np.cross(x[o[1]] - x[o[0]], x[o[2]] - x[o[0]]), axis=1) / 6.0
        else:
            raise KeyError(name)
        np.add.at(S, conn[:, a], w)
    return S


def rel_pervertex(a, b, S, eps_floor=EPS_FLOOR):
    """max_v ||a_v - b_v|| / max(||b_v||, eps_floor * S_v); vertices no element touches (S_v = 0) must match exactly."""
    a = np.asarray(a, dtype=np.float64).reshape(S.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(S.shape[0], -1)
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), eps_floor * S)
    if np.any((den == 0) & (num != 0)):
        return np.inf
    ok = den > 0
    return float((num[ok] / den[ok]).max()) if ok.any() else 0.0


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
    return min(TOL_FD, FD_TOL.get((name, method), TOL_FD))


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
