#!/usr/bin/env python3
"""BASELINE config 4 at scale: Nematic / GradSq / NormSq on a ~4M-triangle disk with a director field — kernel times."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 816
mb.init(0); L = mb.lib()
if len(sys.argv) > 2:
    mb.set_option("fd_elementwise", int(sys.argv[2]))     # 0: vertex-centric finite-difference kernels
v, faces, bnd = mg.disk(n, dim=3)
rng = np.random.Generator(np.random.PCG64(12345))
th = 0.3 * rng.uniform(-1, 1, size=v.shape[0]) + 0.5 * v[:, 0]
nn = np.stack([np.cos(th), np.sin(th), 0.1 * rng.uniform(-1, 1, size=v.shape[0])], axis=1)
nn /= np.linalg.norm(nn, axis=1, keepdims=True)
m = mb.Mesh(v, edges=bnd, faces=faces)
fld = mb.Field(m, nn)
out = torch.empty(max(v.size, faces.shape[0]), dtype=torch.float64, device="cuda")
ms = C.c_float()
nel = faces.shape[0]
for f, name in ((mb.Nematic(fld, ksplay=1.0, ktwist=0.6, kbend=1.4), "Nematic"), (mb.GradSq(fld), "GradSq"), (mb.NormSq(fld), "NormSq")):
    for mode, mname in ((_lib.MCU_TOTAL, "total"), (_lib.MCU_INTEGRAND, "integrand"), (_lib.MCU_GRADIENT, "gradient"), (_lib.MCU_FIELDGRADIENT, "fieldgradient")):
        d = f._desc(0)
        for _ in range(2):
            _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_sync(); L.mcu_timer_start()
        reps = 5
        for _ in range(reps):
            _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_timer_stop(C.byref(ms))
        t = ms.value / reps * 1e-3
        print(f"{name:8s} {mname:14s} {t * 1e3:9.3f} ms  {nel / t / 1e9:7.2f} G elem/s")
