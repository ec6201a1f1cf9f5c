#!/usr/bin/env python3
"""One case for profilers: Volume.gradient on the tetrahedral cube through the block kernel (mcu_block.cu)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 204
mb.init(0); L = mb.lib()
v, tets = mg.tet_cube(n)
rng = np.random.Generator(np.random.PCG64(12345))
inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
m = mb.Mesh(v, volumes=tets)
out = torch.zeros(m.nv * 3, dtype=torch.float64, device="cuda")
mb.set_option("gradient_path", _lib.MCU_PATH_BLOCK)
d = mb.Volume()._desc(0)
ms = C.c_float()
for _ in range(3):
    _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
L.mcu_sync(); L.mcu_timer_start()
for _ in range(10):
    _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
L.mcu_timer_stop(C.byref(ms))
print(f"Volume.gradient block kernel n={n}: {ms.value / 10:.3f} ms = {24.0 * tets.shape[0] / (ms.value / 10) / 1e6 / 6549.4:.3f} of the HBM roofline")
