#!/usr/bin/env python3
"""Host planner of the patch kernel (mcu_patch_plan, pure host code): wall time and a digest of the whole plan, so that a
change to the planner (threads) can be checked for a bit-identical result.   python scripts/plan_time.py [f]"""
import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morpho_b200 import _lib, meshgen as mg  # noqa: E402

L = _lib.lib()
f = int(sys.argv[1]) if len(sys.argv) > 1 else 707
v, c = mg.icosphere(f)
v = np.ascontiguousarray(mg.perturb_radially(v))
c = np.ascontiguousarray(np.sort(c, axis=1).astype(np.int32))
counts = (C.c_long * 10)()
L.mcu_debug_patch_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
t0 = time.time()
rc = L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, 0, counts, None, None, None)
t1 = time.time() - t0
assert rc == 0
nh, nr, nt, hw = counts[0], counts[1], counts[2], counts[3]
hdr = np.zeros(nh * hw, dtype=np.int32)
runs = np.zeros(nr * 2, dtype=np.int32)
tab = np.zeros(nt, dtype=np.uint32)
L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, 0, counts, hdr.ctypes.data, runs.ctypes.data, tab.ctypes.data)
nw = (C.c_long * 1)()
L.mcu_debug_patch_triangles.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
L.mcu_debug_patch_triangles(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, 0, nw, None, None)
thdr = np.zeros(nh * 4, dtype=np.int32)
tri = np.zeros(nw[0], dtype=np.uint32)
L.mcu_debug_patch_triangles(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, 0, nw, thdr.ctypes.data, tri.ctypes.data)
h = hashlib.sha256()
for a in (hdr, runs, tab, thdr, tri):
    h.update(a.tobytes())
print(f"f={f}: {c.shape[0]} triangles, {nh} patches, plan in {t1:.2f} s, counts {list(counts)}, digest {h.hexdigest()[:16]}")
