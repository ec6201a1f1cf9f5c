#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/refbuild/build_ref.sh) on the seeded cases of cases.py.

Run in the dev container (needs oracle/_ref):   python tests/golden/make_golden.py
The fixtures hold only the reference's raw FP64 outputs (plus a hash of the inputs so that
the tests notice if a generator changes); inputs are regenerated from their seeds.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.oracle import FunctionalSpec, Reference  # noqa: E402
from tests.golden.cases import all_cases, eval_key  # noqa: E402


def input_hash(case):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(case["vert"]).tobytes())
    for g in sorted(case["grades"]):
        h.update(np.ascontiguousarray(case["grades"][g], dtype=np.int32).tobytes())
    for k in sorted(case["fields"]):
        h.update(np.ascontiguousarray(case["fields"][k]).tobytes())
    return h.hexdigest()


def run_case(ref, case):
    out = {"input_hash": np.array(input_hash(case))}
    rm = ref.mesh(case["vert"], case["grades"])
    fh = {k: rm.field(v) for k, v in case["fields"].items()}
    nv, dim = case["vert"].shape
    for i, ev in enumerate(case["evals"]):
        f = dict(ev["f"])
        fnames = f.pop("fields", [])
        spec = FunctionalSpec(**f)
        func = rm.functional(spec, [fh[n] for n in fnames])
        sel = rm.selection(*ev["sel"]) if "sel" in ev else None
        method = ev["method"]
        field = fh[ev["wrt"]] if method == "fieldgradient" else None
        cap = max(nv * 16, max(c.shape[0] for c in case["grades"].values()) + 8)
        try:
            res = rm.eval(func, method, cap, sel=sel, field=field)
        except RuntimeError as e:
            res = np.array([np.nan])
            print("   reference error:", e)
        out[eval_key(i, ev)] = res.copy()
        if sel is not None:
            out[eval_key(i, ev) + "__selorder"] = rm.selection_ids(sel, ev["sel"][0])
    rm.free()
    return out


def main():
    ref = Reference(0)
    for case in all_cases():
        out = run_case(ref, case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']}: {len(case['evals'])} evaluations -> {os.path.basename(path)} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
