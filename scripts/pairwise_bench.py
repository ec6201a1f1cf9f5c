#!/usr/bin/env python3
"""BASELINE.json config 3: all-pairs Coulomb on N points of S^2 — kernel time and FP64 rate (20 flop per ordered pair)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mb.init(0); L = mb.lib()
x = torch.from_numpy(mg.sphere_points(n)).cuda()
out = torch.empty(n * 3, dtype=torch.float64, device="cuda")
ms = C.c_float()
for mode, name in ((_lib.MCU_GRADIENT, "gradient"), (_lib.MCU_INTEGRAND, "integrand")):
    for _ in range(2):
        _lib.check(L.mcu_pairwise_device(n, 3, x.data_ptr(), 0, None, 0.0, mode, out.data_ptr()))
    L.mcu_sync()
    L.mcu_timer_start()
    for _ in range(reps):
        _lib.check(L.mcu_pairwise_device(n, 3, x.data_ptr(), 0, None, 0.0, mode, out.data_ptr()))
    L.mcu_timer_stop(C.byref(ms))
    t = ms.value / reps * 1e-3
    print(f"pairwise {name}: n={n}  {t * 1e3:.2f} ms  {n * n / t / 1e9:.1f} G pairs/s  {20.0 * n * n / t / 1e12:.2f} TFLOP/s (20 flop/pair)")
