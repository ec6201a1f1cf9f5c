#!/usr/bin/env python3
"""Per-patch timeline of the first CTAs of the patch kernel (diagnostics)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
mb.init(0)
L = mb.lib()
for kv in sys.argv[1:]:
    k, v = kv.split("="); mb.set_option(k, int(v))
mb.set_option("patch_dbg", mb.lib().mcu_get_option(b"patch_dbg") | 4)
v, c = mg.icosphere(707); v = mg.perturb_radially(v)
m = mb.Mesh(v, faces=c)
out = torch.empty(v.size, dtype=torch.float64, device="cuda")
buf = torch.zeros(16384 + 4 * 18 * 48 * 3, dtype=torch.int64, device="cuda")
d = mb.Area()._desc()
for _ in range(3):
    _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
L.mcu_debug_patch_profile(buf.data_ptr())
_lib.check(L.mcu_eval_device(m._h, C.byref(d), None, _lib.MCU_GRADIENT, out.data_ptr()))
L.mcu_debug_patch_profile(None)
torch.cuda.synchronize()
tr = buf.cpu().numpy()[16384:].reshape(4, 18, 48, 3)
for cta in (0, 1):
    t = tr[cta]
    t0 = t[:16, 0, 0].min()
    print("CTA", cta, " (cycles relative to the first wait)")
    print(" n | cons wait_begin(min..max) ready(min..max) done(min..max) | prod0 loop_top empty_ok issued | prod1 ...")
    for n in range(5, 25):
        c = t[:16, n, :] - t0
        p0, p1 = t[16, n, :] - t0, t[17, n, :] - t0
        print(f"{n:2d} | {c[:,0].min():7d}..{c[:,0].max():7d}  {c[:,1].min():7d}..{c[:,1].max():7d}  {c[:,2].min():7d}..{c[:,2].max():7d} | {p0[0]:7d} {p0[1]:7d} {p0[2]:7d} | {p1[0]:7d} {p1[1]:7d} {p1[2]:7d}")
