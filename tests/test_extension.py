"""The Morpho extension morphocuda.so behind the unchanged builtin API (morpho_b200/ext/morphocuda.c).

Scripts under tests/morpho_scripts/ run twice through the reference's own VM (oracle/_ref/morpho6, the
unmodified libmorpho): as they are, and with `import morphocuda` prepended.  On a GPU box the second run
evaluates Length/Area/VolumeEnclosed/Volume on the device; every printed number must agree with the stock
run to 1e-12 (direct evaluations) / 1e-9 (optimiser trajectories: 12 CG iterations amplify round-off).
Without a GPU the extension must load, patch nothing and leave the output untouched.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
MORPHO6 = os.path.join(REF, "morpho6")
EXT = os.path.join(REF, "pkg", "lib", "morphocuda.so")
SCRIPTS = os.path.join(ROOT, "tests", "morpho_scripts")
NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")

pytestmark = pytest.mark.skipif(not (os.path.exists(MORPHO6) and os.path.exists(EXT)),
                                reason="oracle/_ref (reference build) or the extension is not built")


def run(script_text, tmp_path, name, env_extra=None):
    path = tmp_path / name
    path.write_text(script_text)
    env = dict(os.environ, HOME=os.path.join(REF, "home"), OPENBLAS_NUM_THREADS="1")
    env.update(env_extra or {})
    p = subprocess.run([MORPHO6, str(path)], capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout.strip().splitlines(), p.stderr


def compare(lines_ref, lines_new, tol):
    assert len(lines_ref) == len(lines_new), (lines_ref, lines_new)
    for a, b in zip(lines_ref, lines_new):
        assert NUM.sub("#", a) == NUM.sub("#", b), (a, b)
        for x, y in zip(NUM.findall(a), NUM.findall(b)):
            x, y = float(x), float(y)
            assert abs(x - y) <= tol * max(abs(x), 1e-30) or abs(x - y) <= 1e-300, (a, b)


def script(name):
    return open(os.path.join(SCRIPTS, name)).read()


def test_extension_loads_and_is_inert_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    src = script("catenoid_cg.morpho")
    ref, _ = run(src, tmp_path, "stock.morpho")
    new, err = run("import morphocuda\n" + src + "\nprint cudastats()\n", tmp_path, "ext.morpho")
    assert "CUDA unavailable" in err                 # loud, and nothing of ours computes on the CPU
    assert new[-1].replace(" ", "") == "[0,0,0,0,0]"
    assert ref == new[:-1]


@pytest.mark.gpu
@pytest.mark.parametrize("name,tol", [("functionals_basic.morpho", 1e-12), ("catenoid_cg.morpho", 1e-9),
                                      ("volume_constraint.morpho", 1e-9), ("field_functionals.morpho", 1e-9)])
def test_scripts_agree_with_the_stock_reference(tmp_path, name, tol):
    src = script(name)
    ref, _ = run(src, tmp_path, "stock.morpho")
    new, err = run("import morphocuda\n" + src + "\nprint cudastats()\n", tmp_path, "ext.morpho",
                   {"MORPHO_CUDA_VERBOSE": "1"})
    assert "ready:" in err, err
    stats = [int(x) for x in NUM.findall(new[-1])]
    evals, vert_uploads, conn_uploads, fallbacks, field_uploads = stats
    assert evals > 0, (stats, err)
    compare(ref, new[:-1], tol)
    # the mirror is re-uploaded only when the VM wrote to it: far fewer uploads than evaluations in the optimisers
    if name in ("catenoid_cg.morpho", "volume_constraint.morpho"):
        assert vert_uploads < evals, stats
    if name == "field_functionals.morpho":
        assert fallbacks == 0 and field_uploads > 0, stats       # every call of the script ran on the GPU


@pytest.mark.gpu
def test_paranoid_mode_matches_hooked_mode(tmp_path):
    src = "import morphocuda\n" + script("functionals_basic.morpho")
    a, _ = run(src, tmp_path, "a.morpho")
    b, _ = run(src, tmp_path, "b.morpho", {"MORPHO_CUDA_PARANOID": "1"})
    assert a == b
