#!/usr/bin/env python3
"""BASELINE config 4 at scale, the closure-integral part: an anchoring-type AreaIntegral with a director field and a
LineIntegral on the boundary of a ~4M-triangle disk.  Times are whole C-ABI calls (program upload, kernels, result
copied to the host).  With oracle/_ref present the reference's own integrator is timed on a bounded sample."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morpho_b200 as mb
from morpho_b200 import meshgen as mg
from morpho_b200.integrand import exp, grad, normal, sin, trace

n = int(sys.argv[1]) if len(sys.argv) > 1 else 816
mb.init(0)
v, faces, bnd = mg.disk(n, dim=3)
v[:, 2] = 0.05 * np.sin(3 * v[:, 0]) * np.cos(2 * v[:, 1])
rng = np.random.Generator(np.random.PCG64(12345))
th = 0.3 * rng.uniform(-1, 1, size=v.shape[0]) + 0.5 * v[:, 0]
nn = np.stack([np.cos(th), np.sin(th), 0.1 * rng.uniform(-1, 1, size=v.shape[0])], axis=1)
m = mb.Mesh(v, edges=bnd, faces=faces)
q = mb.Field(m, nn)
cases = [
    ("AreaIntegral anchoring (q.n)^2", mb.AreaIntegral(lambda x, f: (f[0] * normal()[0] + f[1] * normal()[1] + f[2] * normal()[2]) ** 2, q), faces.shape[0]),
    ("AreaIntegral |grad q|^2 + bulk", mb.AreaIntegral(lambda x, f: sum(grad(f)[k][c] ** 2 for k in range(3) for c in range(3)) + (f[0] ** 2 + f[1] ** 2 + f[2] ** 2 - 1) ** 2, q), faces.shape[0]),
    ("AreaIntegral oscillatory (adaptive)", mb.AreaIntegral(lambda x: exp(x[0]) * sin(400 * x[1])), faces.shape[0]),
    ("LineIntegral boundary", mb.LineIntegral(lambda x, f: f[0] * f[1] + sin(50 * x[0]), q), bnd.shape[0]),
]
for name, f, nel in cases:
    for meth, call in (("total", lambda: f.total(m)), ("integrand", lambda: f.integrand(m)), ("gradient", lambda: f.gradient(m)),
                       ("fieldgradient", (lambda: f.fieldgradient(q)) if f.fields else None)):
        if call is None:
            continue
        call()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r = call()
        t = (time.perf_counter() - t0) / reps
        print(f"{name:38s} {meth:14s} {t * 1e3:10.2f} ms  {nel / t / 1e6:9.1f} M elem/s  (nel {nel})")

try:
    from oracle.oracle import Reference, reference_available
    if reference_available():
        ref = Reference(0)
        ns = 20000
        sub = faces[:ns]
        rm = ref.mesh(v, {2: sub})
        for name, f, _ in cases[:3]:
            prog = trace(f.fn, 3, [3] if f.fields else [])
            if any(int(op) >= 27 for op in prog["op"]):
                continue
            t0 = time.perf_counter()
            rm.integral(2, [tuple(t) for t in prog.tolist()], [nn] if f.fields else [])
            t = time.perf_counter() - t0
            print(f"reference integrator, 1 core: {name:38s} integrand {ns / t / 1e6:9.3f} M elem/s  (sample {ns} triangles)")
except Exception as e:                                                    # noqa: BLE001
    print("reference sample skipped:", e)
