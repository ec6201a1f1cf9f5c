#!/usr/bin/env python3
"""Run the reference's OWN golden-output tests for the hot path through the reference
build in oracle/_ref (our morpho6 driver) — this is what pins oracle/_ref.

Mechanism of the reference's test/test.py (lines 44-47, 96-144) re-stated: every
``// expect: X`` comment is an expected stdout line; error reports are normalised to
``@error[Id]``.  (test.py itself needs the `colored` package, absent here.)

Usage:  python oracle/refbuild/run_ref_tests.py [--dirs functionals mesh ...] [--json out.json]
        [--root oracle/_ref/reftests] [--prepend "import morphocuda"]
--root defaults to /root/reference/test (dev container); oracle/_ref/reftests is the copy build_ref.sh makes
so that the suite also runs where the reference tree is absent.  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFOUT = os.path.join(HERE, "..", "_ref")
ERR, STK = "@error", "@stacktrace"


def simplify_errors(s):
    return re.sub(r".*[E|e]rror[ :]*'([A-z;a-z]*)'.*", ERR + r"[\1]", s.rstrip())


def expected_of(path):
    out = []
    for line in open(path, encoding="utf8"):
        if re.findall(r".*[E|e]rror[ :].*?(.*)", line):
            out.append(simplify_errors(line))
        else:
            out += re.findall(r"// expect: ?(.*)", line)
    return out


def output_of(text):
    lines = [re.sub(r"\x1b[^m]*m", "", l.rstrip()) for l in text.splitlines() if not l.startswith("[morphocuda]")]
    lines = [simplify_errors(l) for l in lines]
    lines = [re.sub(r".*at line.*", STK, l.rstrip()) for l in lines]
    for i in range(len(lines) - 1):
        if re.findall(r"@error.*", lines[i]) and re.findall(r".*in .*", lines[i + 1]):
            lines[i + 1] = STK
    return [l for l in lines if l != STK]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default="/root/reference/test")
    ap.add_argument("--dirs", nargs="*", default=["functionals", "integrate", "mesh", "field", "selection", "sparse", "matrix"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--prepend", default=None, help="line put in front of every script, e.g. 'import morphocuda'")
    ap.add_argument("--jobs", type=int, default=1, help="scripts run concurrently")
    a = ap.parse_args()
    exe = os.path.join(REFOUT, "morpho6")
    env = dict(os.environ, HOME=os.path.join(REFOUT, "home"), OPENBLAS_NUM_THREADS="1")
    files = []
    for d in a.dirs:
        for f in sorted(glob.glob(os.path.join(a.root, d, "**", "*.morpho"), recursive=True)):
            if not os.path.basename(f).startswith("_with_prepend_"):
                files.append(f)

    def one(f):
        run = f
        if a.prepend:
            # same directory (the scripts open their fixtures by relative path); the extra first line shifts
            # nothing the expectations depend on (line numbers in stack traces are filtered out)
            run = os.path.join(os.path.dirname(f), "_with_prepend_" + os.path.basename(f))
            with open(run, "w", encoding="utf8") as fh:
                fh.write(a.prepend + "\n" + open(f, encoding="utf8").read())
        cmd = [exe] + ([f"-w{a.threads}"] if a.threads else []) + [os.path.basename(run)]
        try:
            p = subprocess.run(cmd, cwd=os.path.dirname(f), env=env, capture_output=True, text=True, timeout=300)
            ok = expected_of(f) == output_of(p.stdout + p.stderr)
        except subprocess.TimeoutExpired:
            ok = False
        finally:
            if run != f and os.path.exists(run):
                os.remove(run)
        return os.path.relpath(f, a.root), ok

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(a.jobs, 1)) as ex:
        res = dict(ex.map(one, files))
    npass = sum(res.values())
    print(f"{npass}/{len(res)} reference tests pass on oracle/_ref")
    for k, v in res.items():
        if not v:
            print("  FAILED", k)
    if a.json:
        json.dump({"passed": npass, "total": len(res), "failed": [k for k, v in res.items() if not v]}, open(a.json, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
