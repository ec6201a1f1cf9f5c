#!/usr/bin/env python3
"""BASELINE config 5: tetrahedral cube (n^3 cells x 6 tets): Volume, GradSq (5-component field), Nematic — kernel times."""
import sys, os, ctypes as C, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 204
mb.init(0); L = mb.lib()
if len(sys.argv) > 2:
    mb.set_option("fd_elementwise", int(sys.argv[2]))     # 0: vertex-centric finite-difference kernels
t0 = time.time()
v, tets = mg.tet_cube(n)
rng = np.random.Generator(np.random.PCG64(12345))
inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
m = mb.Mesh(v, volumes=tets)
print(f"tet cube n={n}: {tets.shape[0]} tets, {v.shape[0]} vertices, built + mirrored in {time.time() - t0:.1f}s")
q = np.stack([v[:, 0], 2 * v[:, 1], -v[:, 2], v[:, 0] + v[:, 1], 0 * v[:, 0] + 3.0], axis=1) + 0.01 * rng.standard_normal((v.shape[0], 5))
nn = np.stack([np.cos(v[:, 0] + 0.3 * v[:, 1]), np.sin(v[:, 0] + 0.3 * v[:, 1]), 0.2 * np.sin(4 * v[:, 2])], axis=1)
nn /= np.linalg.norm(nn, axis=1, keepdims=True)
fq, fn = mb.Field(m, q), mb.Field(m, nn)
nel = tets.shape[0]
out = torch.empty(max(v.shape[0] * 5, nel), dtype=torch.float64, device="cuda")
ms = C.c_float()
for f, name, modes in ((mb.Volume(), "Volume", ("total", "integrand", "gradient")),
                       (mb.GradSq(fq), "GradSq(5)", ("total", "integrand", "fieldgradient")),
                       (mb.Nematic(fn, ksplay=1.0, ktwist=0.8, kbend=1.2), "Nematic", ("total", "integrand", "fieldgradient"))):
    for mname in modes:
        mode = {"total": _lib.MCU_TOTAL, "integrand": _lib.MCU_INTEGRAND, "gradient": _lib.MCU_GRADIENT, "fieldgradient": _lib.MCU_FIELDGRADIENT}[mname]
        d = f._desc(0)
        _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_sync(); L.mcu_timer_start()
        reps = 3
        for _ in range(reps):
            _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_timer_stop(C.byref(ms))
        t = ms.value / reps * 1e-3
        print(f"{name:10s} {mname:14s} {t * 1e3:9.3f} ms  {nel / t / 1e9:7.2f} G elem/s")
