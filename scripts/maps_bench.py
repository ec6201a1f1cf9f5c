#!/usr/bin/env python3
"""Throughput of the other maps of the hot path on the headline mesh (total / integrand / gradient on both assembly
paths), data resident, CUDA events on the library stream."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
f = int(sys.argv[1]) if len(sys.argv) > 1 else 707
mb.init(0); L = mb.lib()
for kv in sys.argv[2:]:                       # planner / kernel options, e.g. patch_stages=4
    k, val = kv.split("=")
    mb.set_option(k, int(val))
v, c = mg.icosphere(f); v = mg.perturb_radially(v)
m = mb.Mesh(v, faces=c)
nel = c.shape[0]
out = torch.empty(max(v.size, nel), dtype=torch.float64, device="cuda")
ms = C.c_float()
peak = 6549.4
for cls in (mb.Area, mb.VolumeEnclosed):
    d = cls()._desc()
    for mode, name, bytes_per in ((_lib.MCU_TOTAL, "total(patch)", 24.0), (_lib.MCU_TOTAL, "total(gather)", 24.0), (_lib.MCU_INTEGRAND, "integrand", 32.0), (_lib.MCU_GRADIENT, "gradient(patch)", 36.0), (_lib.MCU_GRADIENT, "gradient(gather)", 36.0)):
        mb.set_option("gradient_path", _lib.MCU_PATH_GATHER if "gather" in name else _lib.MCU_PATH_AUTO)
        for _ in range(3):
            _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_sync(); L.mcu_timer_start()
        reps = 50
        for _ in range(reps):
            _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr()))
        L.mcu_timer_stop(C.byref(ms))
        t = ms.value / reps * 1e-3
        gbs = bytes_per * nel / t / 1e9
        print(f"{cls.__name__:15s} {name:17s} {t * 1e6:8.1f} us  {nel / t / 1e9:7.1f} G elem/s  {gbs:7.0f} GB/s algorithmic = {gbs / peak:.3f} of measured peak")
