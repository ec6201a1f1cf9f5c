#!/usr/bin/env python3
"""Generate tests/golden/pairwise.npz from the UNMODIFIED, INTERPRETED reference: PairwisePotential of
modules/functionals.morpho:87-141 with real closures, run by the reference VM (oracle/_ref/morpho6) on point sets read
from .mesh files — SURVEY §8d config 3 asks for exactly this at N = 500 / 1000 / 2000.  Values are printed with %.17g.

Run in the dev container (needs oracle/_ref, i.e. /root/reference):   python tests/golden/make_pairwise_golden.py
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from morpho_b200 import meshgen as mg  # noqa: E402
from tests.golden.pairwise_cases import CASES, points  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    store = {}
    with tempfile.TemporaryDirectory() as tmp:
        src = ["import meshtools", "import functionals", ""]
        for name, n, kind, params, cutoff, func, grad in CASES:
            x = points(name, n)
            store[name + "__x"] = x
            mg.write_mesh_file(os.path.join(tmp, name + ".mesh"), x)
            cut = f", cutoff={cutoff!r}" if cutoff > 0 else ""
            src += [f'var m_{name} = Mesh("{name}.mesh")',
                    f"var p_{name} = PairwisePotential({func}, {grad}{cut})",
                    f'print "@{name}__total"',
                    f'print p_{name}.total(m_{name}).format("%.17g")',
                    f"var i_{name} = p_{name}.integrand(m_{name})",
                    f'print "@{name}__integrand"',
                    f'for (i in 0...{n}) print i_{name}[i].format("%.17g")',
                    f"var g_{name} = p_{name}.gradient(m_{name})",
                    f'print "@{name}__gradient"',
                    f'for (i in 0...{n}) for (k in 0...{x.shape[1]}) print g_{name}[k,i].format("%.17g")']
        path = os.path.join(tmp, "pairwise.morpho")
        open(path, "w").write("\n".join(src) + "\n")
        env = dict(os.environ, HOME=os.path.join(REF, "home"), OPENBLAS_NUM_THREADS="1")
        t0 = time.time()
        p = subprocess.run([os.path.join(REF, "morpho6"), path], capture_output=True, text=True, cwd=tmp, env=env, timeout=3600)
        print(f"interpreted reference took {time.time() - t0:.1f} s")
        if p.returncode != 0 or "rror" in p.stdout:
            raise SystemExit(p.stdout[-2000:] + p.stderr[-2000:])
    cur = None
    vals = {}
    for line in p.stdout.splitlines():
        if line.startswith("@"):
            cur = line[1:]
            vals[cur] = []
        elif cur and line.strip():
            vals[cur].append(float(line))
    for name, n, *_ in CASES:
        dim = store[name + "__x"].shape[1]
        store[name + "__total"] = np.array(vals[name + "__total"])
        store[name + "__integrand"] = np.array(vals[name + "__integrand"])
        store[name + "__gradient"] = np.array(vals[name + "__gradient"]).reshape(n, dim)
        assert store[name + "__integrand"].shape == (n,)
        print(f"{name:18s} total={store[name + '__total'][0]:+.15e} |g|={np.linalg.norm(store[name + '__gradient']):.6e}")
    np.savez_compressed(os.path.join(HERE, "pairwise.npz"), **store)
    print("wrote pairwise.npz", os.path.getsize(os.path.join(HERE, "pairwise.npz")), "bytes")


if __name__ == "__main__":
    main()
