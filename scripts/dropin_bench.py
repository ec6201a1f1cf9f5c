#!/usr/bin/env python3
"""Script-level drop-in timing: the SAME Morpho script (load a mesh file, then time Area/VolumeEnclosed
total + gradient calls with System.clock) through the unmodified VM, stock and with `import morphocuda`."""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morpho_b200 import meshgen as mg
f = int(sys.argv[1]) if len(sys.argv) > 1 else 224
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
REF = os.path.join(ROOT, "oracle", "_ref")
tmp = tempfile.mkdtemp()
v, c = mg.icosphere(f)
v = mg.perturb_radially(v)
path = os.path.join(tmp, "sphere.mesh")
mg.write_mesh_file(path, v, g2=c)
script = f'''
var m = Mesh("{path}")
var la = Area()
var lv = VolumeEnclosed()
var t0 = System.clock()
var a = la.total(m)
var g = la.gradient(m)
print "first call (mirror + patch plan) ${{System.clock()-t0}}"
var vm = m.vertexmatrix()
t0 = System.clock()
for (i in 1..{reps}) {{
  var e = la.total(m) + lv.total(m)
  var ga = la.gradient(m)
  var gv = lv.gradient(m)
  vm.acc(-1e-6, ga)              // an optimiser step: the vertex matrix changes, the mirror is refreshed
}}
var dt = System.clock()-t0
print "per iteration (2 totals + 2 gradients + vertex update) ${{dt/{reps}}}"
print "area ${{la.total(m).format("%.12g")}} volume ${{lv.total(m).format("%.12g")}}"
'''
env = dict(os.environ, HOME=os.path.join(REF, "home"), OPENBLAS_NUM_THREADS="1")
for name, src in (("stock reference", script), ("with morphocuda", "import morphocuda\n" + script + "print cudastats()\n")):
    p = os.path.join(tmp, name.split()[0] + ".morpho")
    open(p, "w").write(src)
    t = time.time()
    r = subprocess.run([os.path.join(REF, "morpho6"), p], capture_output=True, text=True, env=env, timeout=3000)
    print(f"--- {name}: f={f}, {c.shape[0]} triangles (whole script {time.time() - t:.1f}s)")
    print(r.stdout.strip())
    if r.returncode:
        print(r.stderr[-500:])
