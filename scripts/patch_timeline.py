#!/usr/bin/env python3
"""Timeline diagnostics of the patch kernel on the headline mesh: how long consumer warps wait for data."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
f = int(sys.argv[1]) if len(sys.argv) > 1 else 707
mb.init(0)
L = mb.lib()
for kv in sys.argv[2:]:
    k, v = kv.split("="); mb.set_option(k, int(v))
v, c = mg.icosphere(f); v = mg.perturb_radially(v)
m = mb.Mesh(v, faces=c)
out = torch.empty(v.size, dtype=torch.float64, device="cuda")
ncta, nw = 296, 16
buf = torch.zeros(ncta * nw * 2, dtype=torch.int64, device="cuda")
for cls in (mb.Area, mb.VolumeEnclosed):
    d = cls()._desc()
    for _ in range(3):
        _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
    L.mcu_debug_patch_profile(buf.data_ptr())
    _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
    L.mcu_debug_patch_profile(None)
    torch.cuda.synchronize()
    a = buf.cpu().numpy().reshape(ncta, nw, 2).astype(np.float64)
    wait, alive = a[..., 0], a[..., 1]
    print(cls.__name__, "alive cycles: mean %.0f max %.0f" % (alive.mean(), alive.max()), " wait fraction mean %.3f" % (wait / alive).mean())
    print("  wait fraction by warp:", np.round((wait / alive).mean(axis=0), 3))
    print("  wait fraction by CTA (percentiles 0,25,50,75,100):", np.round(np.percentile((wait / alive).mean(axis=1), [0, 25, 50, 75, 100]), 3))
