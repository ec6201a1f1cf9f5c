#!/usr/bin/env python3
"""Generate tests/golden/integrals.npz from the UNMODIFIED reference, twice:

 (1) through the reference's VM: a generated Morpho script loads each mesh from a .mesh file, builds the fields and
     prints ``LineIntegral(closure, fields...).integrand(mesh)`` / ``AreaIntegral(...)`` — the real builtin with a real
     closure (functional.c:5186-5243, 5371-5426);
 (2) through oracle/refbuild/ref_harness.c's refh_integral: the reference's own integrate_integrate with the
     bytecode interpreter as the integrand callback.

Both must agree (1e-13 of the element scale) — that pins the bytecode route the GPU tests use as their live
reference — and (1) is what gets stored.   Run in the dev container:   python tests/golden/make_integral_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from morpho_b200 import meshgen as mg  # noqa: E402
from morpho_b200.integrand import trace  # noqa: E402
from oracle.oracle import Reference  # noqa: E402
from tests.golden.integral_cases import CASES, FIELD_SRC, GRAD_CASES, METHOD_CASES, METHOD_GRAD_CASES, POT_CASES, fields_for, meshes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def vm_integrands(tmp, ms):
    src = ["import meshtools", ""]
    for name, (v, gr) in ms.items():
        mg.write_mesh_file(os.path.join(tmp, name + ".mesh"), v, **{f"g{g}": c for g, c in gr.items()})
        src.append(f'var m_{name} = Mesh("{name}.mesh")')
        for fname, fsrc in FIELD_SRC[v.shape[1]].items():
            src.append(f"var {fname}_{name} = Field(m_{name}, {fsrc})")
    for name, mesh, grade, _, closure, flds, sel in CASES:
        cls = {1: "LineIntegral", 2: "AreaIntegral", 3: "VolumeIntegral"}[grade]
        fargs = "".join(f", {f}_{mesh}" for f in flds)
        src.append(f"var r_{name} = {cls}({closure}{fargs}).integrand(m_{mesh})")
        src.append(f'print "@{name}"')
        src.append(f'for (i in 0...r_{name}.dimensions()[1]) print r_{name}[0,i].format("%.17g")')
    bycase = {c[0]: c for c in CASES}
    for gname, wrt in GRAD_CASES:
        name, mesh, grade, _, closure, flds, sel = bycase[gname]
        cls = {1: "LineIntegral", 2: "AreaIntegral", 3: "VolumeIntegral"}[grade]
        fargs = "".join(f", {f}_{mesh}" for f in flds)
        key = f"grad__{gname}__{'x' if wrt < 0 else wrt}"
        selarg = ""
        if sel is not None:
            src.append(f"var s_{key} = Selection(m_{mesh})")
            for i in sel:
                src.append(f"s_{key}[{grade},{i}] = true")
            selarg = f", s_{key}"
        src.append(f'print "@{key}"')
        if wrt < 0:
            src.append(f"var g_{key} = {cls}({closure}{fargs}).gradient(m_{mesh}{selarg})")
            src.append(f"var d_{key} = g_{key}.dimensions()")
            src.append(f'for (i in 0...d_{key}[1]) for (k in 0...d_{key}[0]) print g_{key}[k,i].format("%.17g")')
        else:
            src.append(f"var g_{key} = {cls}({closure}{fargs}).fieldgradient({flds[wrt]}_{mesh}, m_{mesh}{selarg})")
            src.append(f'for (x in g_{key}) {{ if (ismatrix(x)) {{ for (y in x) print y.format("%.17g") }} else print x.format("%.17g") }}')
    CLS = {1: "LineIntegral", 2: "AreaIntegral", 3: "VolumeIntegral"}
    bymethod = {m[0]: m for m in METHOD_CASES}
    for mname, base, msrc, _ in METHOD_CASES:
        name, mesh, grade, _, closure, flds, sel = bycase[base]
        fargs = "".join(f", {f}_{mesh}" for f in flds)
        selarg = ""
        if sel is not None:
            src.append(f"var s_{mname} = Selection(m_{mesh})")
            for i in sel:
                src.append(f"s_{mname}[{grade},{i}] = true")
            selarg = f", s_{mname}"
        src.append(f"var f_{mname} = {CLS[grade]}({closure}{fargs}, method={msrc})")
        src.append(f"var r_{mname} = f_{mname}.integrand(m_{mesh}{selarg})")
        src.append(f'print "@{mname}"')
        src.append(f'for (i in 0...r_{mname}.dimensions()[1]) print r_{mname}[0,i].format("%.17g")')
    for mname, wrt in METHOD_GRAD_CASES:
        base = bymethod[mname][1]
        name, mesh, grade, _, closure, flds, sel = bycase[base]
        key = f"grad__{mname}__{'x' if wrt < 0 else wrt}"
        selarg = f", s_{mname}" if sel is not None else ""
        src.append(f'print "@{key}"')
        if wrt < 0:
            src.append(f"var g_{key} = f_{mname}.gradient(m_{mesh}{selarg})")
            src.append(f"var d_{key} = g_{key}.dimensions()")
            src.append(f'for (i in 0...d_{key}[1]) for (k in 0...d_{key}[0]) print g_{key}[k,i].format("%.17g")')
        else:
            src.append(f"var g_{key} = f_{mname}.fieldgradient({flds[wrt]}_{mesh}, m_{mesh}{selarg})")
            src.append(f'for (x in g_{key}) {{ if (ismatrix(x)) {{ for (y in x) print y.format("%.17g") }} else print x.format("%.17g") }}')
    for name, mesh, _, gradf, fsrc, gsrc, sel in POT_CASES:
        src.append(f"var p_{name} = ScalarPotential({fsrc}{', ' + gsrc if gsrc else ''})")
        selarg = ""
        if sel is not None:
            src.append(f"var s_{name} = Selection(m_{mesh})")
            for i in sel:
                src.append(f"s_{name}[0,{i}] = true")
            selarg = f", s_{name}"
        src.append(f'print "@{name}__total"')
        src.append(f'print p_{name}.total(m_{mesh}{selarg}).format("%.17g")')
        src.append(f'print "@{name}__integrand"')
        src.append(f"var pi_{name} = p_{name}.integrand(m_{mesh}{selarg})")
        src.append(f'for (i in 0...pi_{name}.dimensions()[1]) print pi_{name}[0,i].format("%.17g")')
        src.append(f'print "@{name}__gradient"')
        src.append(f"var pg_{name} = p_{name}.gradient(m_{mesh}{selarg})")
        src.append(f"var pd_{name} = pg_{name}.dimensions()")
        src.append(f'for (i in 0...pd_{name}[1]) for (k in 0...pd_{name}[0]) print pg_{name}[k,i].format("%.17g")')
    path = os.path.join(tmp, "integrals.morpho")
    open(path, "w").write("\n".join(src) + "\n")
    env = dict(os.environ, HOME=os.path.join(REF, "home"), OPENBLAS_NUM_THREADS="1")
    p = subprocess.run([os.path.join(REF, "morpho6"), path], capture_output=True, text=True, cwd=tmp, env=env, timeout=600)
    if p.returncode != 0 or "rror" in p.stdout:
        raise SystemExit(p.stdout + p.stderr)
    out, cur = {}, None
    for line in p.stdout.splitlines():
        if line.startswith("@"):
            cur = line[1:]
            out[cur] = []
        elif cur and line.strip():
            out[cur].append(float(line))
    return {k: np.array(v) for k, v in out.items()}


def main():
    ms = meshes()
    with tempfile.TemporaryDirectory() as tmp:
        vm = vm_integrands(tmp, ms)
    ref = Reference(0)
    store = {}
    for name, mesh, grade, fn, _, flds, sel in CASES:
        v, gr = ms[mesh]
        fd = fields_for(v)
        prog = trace(fn, v.shape[1], [fd[f].shape[1] for f in flds])
        g = vm[name]
        store[name] = g
        if any(int(op) >= 27 for op in prog["op"]):          # tangent()/normal()/grad(): only the VM provides them
            print(f"{name:12s} n={g.size:5d} sum={g.sum():+.15e}  (element helpers: VM only)")
            continue
        rm = ref.mesh(v, gr)
        h = rm.integral(grade, [tuple(t) for t in prog.tolist()], [fd[f] for f in flds])
        rm.free()
        assert g.shape == h.shape, (name, g.shape, h.shape)
        err = np.max(np.abs(g - h)) / max(np.max(np.abs(g)), 1e-300)
        print(f"{name:12s} n={g.size:5d} sum={g.sum():+.15e}  VM closure vs harness bytecode: {err:.2e}")
        assert err < 1e-13, name
    for gname, wrt in GRAD_CASES:
        key = f"grad__{gname}__{'x' if wrt < 0 else wrt}"
        store[key] = vm[key]
        print(f"{key:28s} n={vm[key].size:5d} |g|={np.linalg.norm(vm[key]):.15e}")
    for mname, base, _, _ in METHOD_CASES:
        store[mname] = vm[mname]
        print(f"{mname:18s} n={vm[mname].size:5d} sum={vm[mname].sum():+.15e}  (default integrator on the same case: {store[base].sum():+.15e})")
    for mname, wrt in METHOD_GRAD_CASES:
        key = f"grad__{mname}__{'x' if wrt < 0 else wrt}"
        store[key] = vm[key]
        print(f"{key:28s} n={vm[key].size:5d} |g|={np.linalg.norm(vm[key]):.15e}")
    for name, *_ in POT_CASES:
        for what in ("total", "integrand", "gradient"):
            store[f"{name}__{what}"] = vm[f"{name}__{what}"]
        print(f"{name:12s} total={vm[name + '__total'][0]:+.15e} |g|={np.linalg.norm(vm[name + '__gradient']):.6e}")
    np.savez_compressed(os.path.join(HERE, "integrals.npz"), **store)
    print("wrote integrals.npz")


if __name__ == "__main__":
    main()
