#!/usr/bin/env python3
"""Small invocation of every kernel family (patch gradient/total, generic maps, FD family, pairwise, addgrade): a quick
"does everything launch and return" case (compute-sanitizer is closed on this GPU pool, so no memcheck run exists)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morpho_b200 as mb
from morpho_b200 import meshgen as mg
mb.init(0)
mb.set_option("patch_min_elements", 0)
v, c = mg.icosphere(24); v = mg.perturb_radially(v)
m = mb.Mesh(v, faces=c)
for cls in (mb.Area, mb.VolumeEnclosed):
    f = cls()
    for path in (mb._lib.MCU_PATH_PATCH, mb._lib.MCU_PATH_GATHER):
        mb.set_option("gradient_path", path)
        g = f.gradient(m); t = f.total(m)
    mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
    f.integrand(m)
edges = m.addgrade(1)
mb.Length().gradient(m); mb.Length().total(m)
sel = mb.Selection(m, ids={2: np.arange(0, c.shape[0], 3, dtype=np.int32)})
mb.Area().gradient(m, sel); mb.Area().total(m, sel)
rng = np.random.default_rng(0)
nn = rng.standard_normal((v.shape[0], 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
fld = mb.Field(m, nn); phi = mb.Field(m, rng.standard_normal(v.shape[0]))
for f in (mb.Nematic(fld, pitch=0.3), mb.GradSq(fld), mb.NormSq(fld), mb.NematicElectric(fld, phi)):
    f.total(m); f.integrand(m); f.gradient(m); f.fieldgradient(fld, m)
vt, t = mg.tet_cube(4)
mt = mb.Mesh(vt + 0.01 * rng.standard_normal(vt.shape), volumes=t)
mb.Volume().gradient(mt); mb.Volume().total(mt); mt.addgrade(2); mt.addgrade(1)
q5 = mb.Field(mt, rng.standard_normal((vt.shape[0], 5)))
mb.GradSq(q5).fieldgradient(q5, mt)
vl, e = mg.polyline_loop(40)
ml = mb.Mesh(vl, edges=e)
mb.LineCurvatureSq().gradient(ml); mb.LineCurvatureSq().total(ml)
x = mg.sphere_points(700)
pm = mb.Mesh(x)
p = mb.PairwisePotential("coulomb", cutoff=0.7)
p.total(pm); p.gradient(pm)
print("all kernel families ran")
