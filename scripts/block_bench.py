#!/usr/bin/env python3
"""Block kernel (mcu_block.cu) against the CSR-transpose gather: Volume.gradient on the tetrahedral cube (BASELINE config 5),
Area / VolumeEnclosed gradients on the headline sphere restricted to a Selection, Length.gradient on its edges."""
import sys, os, ctypes as C, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 204
f = int(sys.argv[2]) if len(sys.argv) > 2 else 707
mb.init(0); L = mb.lib()
PEAK = 6549.4
ms = C.c_float()

def timed(m, func, sel, out, reps=10):
    d = func._desc(0)
    h = sel._handle(func._loop_grade(m)) if sel is not None else None
    t0 = time.time()
    _lib.check(L.mcu_eval_device(m._h, C.byref(d), h, _lib.MCU_GRADIENT, out.data_ptr())); L.mcu_sync()
    first = time.time() - t0
    L.mcu_timer_start()
    for _ in range(reps):
        _lib.check(L.mcu_eval_device(m._h, C.byref(d), h, _lib.MCU_GRADIENT, out.data_ptr()))
    L.mcu_timer_stop(C.byref(ms))
    return ms.value / reps, first

def compare(label, m, func, sel, nel, bytes_per_el):
    out = torch.zeros(m.nv * m.dim, dtype=torch.float64, device="cuda")
    res = {}
    for path, pname in ((_lib.MCU_PATH_GATHER, "gather"), (_lib.MCU_PATH_BLOCK, "block")):
        mb.set_option("gradient_path", path)
        t, first = timed(m, func, sel, out)
        res[pname] = out.clone()
        print(f"{label:34s} {pname:7s} {t:8.3f} ms  {nel / t / 1e6:8.2f} G elem/s  {bytes_per_el * nel / t / 1e6 / PEAK:6.3f} of HBM roofline   (first call {first:.2f} s)")
    mb.set_option("gradient_path", _lib.MCU_PATH_AUTO)
    a, b = res["gather"].view(-1, m.dim), res["block"].view(-1, m.dim)
    print(f"{'':34s} max row difference / max row norm = {((a - b).norm(dim=1).max() / a.norm(dim=1).max()).item():.2e}")

v, tets = mg.tet_cube(n)
rng = np.random.Generator(np.random.PCG64(12345))
inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
m = mb.Mesh(v, volumes=tets)
compare(f"Volume.gradient {tets.shape[0]} tets", m, mb.Volume(), None, tets.shape[0], 24.0)
sel = mb.Selection(m, ids={3: np.arange(0, tets.shape[0], 2, dtype=np.int32)})
compare("Volume.gradient, every 2nd tet", m, mb.Volume(), sel, tets.shape[0] // 2, 24.0)
del m

v, c = mg.icosphere(f)
v = mg.perturb_radially(v)
m = mb.Mesh(v, faces=c)
half = np.nonzero(v[c[:, 0], 2] > 0)[0].astype(np.int32)          # the upper hemisphere
sel = mb.Selection(m, ids={2: half})
compare(f"Area.gradient, {len(half)} selected", m, mb.Area(), sel, len(half), 36.0)
compare("VolumeEnclosed.gradient, selected", m, mb.VolumeEnclosed(), sel, len(half), 36.0)
compare(f"Area.gradient, all {c.shape[0]} (block)", m, mb.Area(), None, c.shape[0], 36.0)
