"""LineIntegral / AreaIntegral / VolumeIntegral: the device quadrature (morpho_b200/csrc/mcu_integral.cu) against golden vectors made
by the unmodified reference running real closures in its VM (tests/golden/make_integral_golden.py), and against the
reference's own integrate_integrate called live through the harness where oracle/_ref is present.

Tolerance 1e-12 of the largest element value: the subdivision decisions are discrete (error estimate < 1e-6), so a
device that took a different refinement path anywhere would miss by ~1e-8 or more and fail."""
import os

import numpy as np
import pytest

from tests.golden.integral_cases import CASES, GRAD_CASES, METHOD_CASES, METHOD_GRAD_CASES, POT_CASES, fields_for, meshes

GOLD = os.path.join(os.path.dirname(__file__), "golden", "integrals.npz")
TOL = 1e-12


def _prog(fn, dim, flds, fd):
    from morpho_b200.integrand import trace
    return trace(fn, dim, [fd[f].shape[1] for f in flds])


# ---------------------------------------------------------------- CPU: tracer + the pinned checker
def test_tracer_programs_are_well_formed():
    from morpho_b200.integrand import OPS
    ms = meshes()
    for name, mesh, grade, fn, _, flds, sel in CASES:
        v, _ = ms[mesh]
        prog = _prog(fn, v.shape[1], flds, fields_for(v))
        depth = 0
        for op, a, c in prog.tolist():
            if op in (OPS["const"], OPS["pos"], OPS["field"], OPS["tangent"], OPS["normal"], OPS["gradq"]):
                depth += 1
            elif OPS["add"] <= op <= OPS["div"] or op in (OPS["pow"], OPS["atan2"]):
                assert depth >= 2, name
                depth -= 1
            else:
                assert depth >= 1, name
        assert depth == 1, name


def test_reference_integrator_reproduces_vm_goldens(reference):
    gold = np.load(GOLD)
    ms = meshes()
    for name, mesh, grade, fn, _, flds, sel in CASES:
        v, gr = ms[mesh]
        fd = fields_for(v)
        prog = _prog(fn, v.shape[1], flds, fd)
        if any(int(op) >= 27 for op in prog["op"]):
            continue                              # tangent()/normal()/grad(): element constants only the VM supplies
        rm = reference.mesh(v, gr)
        got = rm.integral(grade, [tuple(t) for t in prog.tolist()], [fd[f] for f in flds])
        rm.free()
        np.testing.assert_allclose(got, gold[name], rtol=0, atol=1e-13 * np.max(np.abs(gold[name])), err_msg=name)


def test_goldens_against_closed_forms():
    """Independent of any code of ours: sum over the sphere of the constant integrand is the mesh area, and the
    polynomial line integrand has a closed form per segment."""
    gold = np.load(GOLD)
    ms = meshes()
    v, gr = ms["sphere"]
    t = v[gr[2]]
    area = 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)
    np.testing.assert_allclose(gold["area_const"], area, rtol=1e-13)
    v, gr = ms["loop"]
    a, b = v[gr[1][:, 0]], v[gr[1][:, 1]]
    ln = np.linalg.norm(b - a, axis=1)
    # int_0^1 (x0+t dx)^2 (y0+t dy) dt + 1
    x0, dx, y0, dy = a[:, 0], b[:, 0] - a[:, 0], a[:, 1], b[:, 1] - a[:, 1]
    exact = (x0 * x0 * y0 + (2 * x0 * dx * y0 + x0 * x0 * dy) / 2 + (dx * dx * y0 + 2 * x0 * dx * dy) / 3 + dx * dx * dy / 4 + 1) * ln
    np.testing.assert_allclose(gold["line_poly"], exact, rtol=1e-12)


def _rules():
    import ctypes as C
    from morpho_b200 import _lib
    L = _lib.lib()
    out = []
    for idx in range(64):
        nm, info = C.create_string_buffer(32), (C.c_int * 5)()
        nd, w = np.zeros(1024), np.zeros(1024)
        if L.mcu_debug_quadrule(idx, nm, info, nd.ctypes.data, w.ctypes.data, 1024) != 0:
            break
        g, order, nn, ext = info[0], info[1], info[2], info[3]
        out.append(dict(name=nm.value.decode(), grade=g, order=order, nnodes=nn, ext=ext, nodes=nd[: nn * (g + 1)].reshape(nn, g + 1).copy(), w=w[:nn].copy()))
    return out


def test_quadrature_rules_are_consistent():
    """The generated rule tables (mcu_quadrules.h) without reference to anything: 33 rules, barycentric nodes, weights that
    sum to one, extension chains that share their leading nodes, and every rule integrates the monomials of its stated
    degree over the simplex exactly (the reference's one-node 'simpson' is kept as it is and excluded)."""
    from math import factorial
    R = _rules()
    assert len(R) == 33
    assert [r["name"] for r in R][:10] == ["midpoint", "simpson", "gauss1", "kronrod3", "gauss2", "kronrod5", "gauss5", "kronrod11", "gauss7", "kronrod15"]
    for r in R:
        if r["name"] == "simpson":
            continue
        np.testing.assert_allclose(r["nodes"].sum(axis=1), 1.0, atol=2e-15, err_msg=r["name"])
        assert abs(r["w"].sum() - 1.0) < 5e-15, r["name"]
        if r["ext"] >= 0:
            e = R[r["ext"]]
            assert e["grade"] == r["grade"] and e["nnodes"] >= r["nnodes"]
            if e["name"] != "simpson":
                np.testing.assert_array_equal(e["nodes"][: r["nnodes"]], r["nodes"])
        # exactness: the integral of prod lambda_i^a_i over the unit-measure simplex is g! prod a_i! / (g + sum a)!
        g = r["grade"]
        deg = r["order"] if g > 1 else r["order"]
        # the reference's `order` is a label, not always the algebraic degree: tri20 is the degree-7 Grundmann-Moeller rule,
        # the Shunn-Ham rules with 35 / 56 points have degree 6 / 8, Kronrod rules with n points have degree (3n+1)/2
        deg = min(deg, {"tri20": 7, "tet5": 6, "tet6": 8, "kronrod3": 5, "kronrod5": 7, "kronrod11": 16, "kronrod15": 22}.get(r["name"], deg))
        rng = np.random.Generator(np.random.PCG64(1))
        for _ in range(12):
            a = rng.multinomial(rng.integers(0, deg + 1), np.ones(g + 1) / (g + 1))
            exact = factorial(g) * np.prod([factorial(int(k)) for k in a]) / factorial(g + int(a.sum()))
            got = float(np.sum(r["w"] * np.prod(r["nodes"] ** a, axis=1)))
            assert abs(got - exact) < 2e-13, (r["name"], a, got, exact)


def test_quadrature_rule_selection_follows_the_reference():
    """integrator_configure (integrate.c:2652-2704): by name, by degree, defaults, and where the error estimate comes from."""
    import ctypes as C
    from morpho_b200 import _lib
    L = _lib.lib()
    names = [r["name"] for r in _rules()]

    def conf(grade, adapt=True, degree=-1, name=b""):
        out = (C.c_int * 2)()
        rc = L.mcu_debug_quadconfigure(grade, 1 if adapt else 0, degree, name, out)
        return None if rc else (names[out[0]], names[out[1]] if out[1] >= 0 else None)

    assert conf(1) == ("gauss5", None) and conf(2) == ("cubtri7", None) and conf(3) == ("grundmann3d0", None)
    assert conf(1, name=b"kronrod15") == ("gauss7", None)          # a rule that IS an extension is reached through its base rule
    assert conf(1, name=b"kronrod15", adapt=False) == ("kronrod15", None)
    assert conf(3, name=b"keast4") == ("keast4", "keast5")         # no extension: the next rule of higher degree estimates the error
    assert conf(3, name=b"tet6") == ("tet6", "grundmann3d5")
    assert conf(3, name=b"grundmann3d5") == ("grundmann3d4", None)
    assert conf(2, name=b"tri20") == ("tri10", None)
    assert conf(1, degree=6) == ("gauss2", None)                   # the lowest rule of order >= 6 is kronrod5, reached through gauss2
    assert conf(2, degree=4) == ("tri10", None)
    assert conf(3, degree=5) == ("keast5", "tet5")                 # ties in order go to the rule listed first
    assert conf(1, name=b"Foo") is None and conf(1, degree=10000) is None


@pytest.mark.skipif(not os.path.exists("/root/reference/src/geometry/integrate.c"), reason="the reference tree is not present")
def test_quadrature_rules_match_the_reference_tables():
    import re
    src = open("/root/reference/src/geometry/integrate.c").read()

    def table(name):
        body = re.sub(r"//.*", "", re.search(r"double " + name + r"\[\]\s*=\s*\{(.*?)\};", src, re.S).group(1))
        return np.array([float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", body)])

    decl = re.findall(r'quadraturerule (\w+) = \{\s*\.name = "(\w+)",\s*\.grade = (\d+),\s*\.order = (\d+),[^\n]*\n\s*\.nnodes = (\d+),'
                      r"\s*\.nodes = (\w+),\s*\.weights = (\w+),\s*\.ext = (&?\w+)", src)
    order = re.findall(r"&(\w+)", re.search(r"quadraturerule \*quadrules\[\] = \{(.*?)NULL", src, re.S).group(1))
    byvar = {d[0]: d for d in decl}
    R = _rules()
    assert len(order) == len(R) == 33
    for r, var in zip(R, order):
        _, name, grade, ordr, nn, nodes, wts, ext = byvar[var]
        assert (r["name"], r["grade"], r["order"], r["nnodes"]) == (name, int(grade), int(ordr), int(nn))
        assert (r["ext"] < 0) == (ext == "NULL") and (ext == "NULL" or order[r["ext"]] == ext[1:])
        np.testing.assert_allclose(r["nodes"].ravel(), table(nodes)[: int(nn) * (int(grade) + 1)], rtol=0, atol=5e-16, err_msg=name)
        np.testing.assert_allclose(r["w"], table(wts)[: int(nn)], rtol=0, atol=5e-16, err_msg=name)


# ---------------------------------------------------------------- GPU: the device quadrature
@pytest.fixture(scope="module")
def mb():
    import morpho_b200 as mb
    mb.init(0)
    return mb


def _device(mb, case, ms):
    name, mesh, grade, fn, _, flds, sel = case
    v, gr = ms[mesh]
    fd = fields_for(v)
    m = mb.Mesh(v, edges=gr.get(1), faces=gr.get(2), volumes=gr.get(3))
    fl = [mb.Field(m, fd[f]) for f in flds]
    cls = {1: mb.LineIntegral, 2: mb.AreaIntegral, 3: mb.VolumeIntegral}[grade]
    args = [m] + ([mb.Selection(m, ids={grade: sel})] if sel is not None else [])
    func = cls(fn, *fl)
    return func.integrand(*args), func.total(*args)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_device_quadrature_matches_vm_goldens(mb, case):
    gold = np.load(GOLD)[case[0]].copy()
    if case[6] is not None:
        mask = np.zeros(gold.size, dtype=bool)
        mask[case[6]] = True
        gold[~mask] = 0.0                         # unselected elements keep 0 (functional.c:226-229)
    integrand, total = _device(mb, case, meshes())
    scale = np.max(np.abs(gold))
    err = np.max(np.abs(integrand - gold)) / scale
    print(case[0], f"integrand err {err:.2e}")
    assert err <= TOL, case[0]
    s = c = 0.0
    for y in gold:                                # the reference's serial Kahan sum (functional.c:243-253)
        yy = y - c
        tt = s + yy
        c = (tt - s) - yy
        s = tt
    assert abs(total - s) <= TOL * max(abs(s), scale), case[0]


@pytest.mark.gpu
@pytest.mark.parametrize("gcase", GRAD_CASES, ids=[f"{n}-{'x' if w < 0 else w}" for n, w in GRAD_CASES])
def test_device_integral_gradients_match_vm_goldens(mb, gcase):
    """gradient / fieldgradient = central differences over the quadrature (functional.c:1142-1162, 1305-1328): the
    reference's values carry ~1e-10 of round-off amplified by 1/(2 eps), hence the finite-difference tolerance of
    tests/parity.py (2e-9 of the gradient scale)."""
    gname, wrt = gcase
    case = next(c for c in CASES if c[0] == gname)
    name, mesh, grade, fn, _, flds, sel = case
    gold = np.load(GOLD)[f"grad__{gname}__{'x' if wrt < 0 else wrt}"]
    v, gr = meshes()[mesh]
    fd = fields_for(v)
    m = mb.Mesh(v, edges=gr.get(1), faces=gr.get(2), volumes=gr.get(3))
    fl = [mb.Field(m, fd[f]) for f in flds]
    func = {1: mb.LineIntegral, 2: mb.AreaIntegral, 3: mb.VolumeIntegral}[grade](fn, *fl)
    args = [m] + ([mb.Selection(m, ids={grade: sel})] if sel is not None else [])
    got = func.gradient(*args) if wrt < 0 else func.fieldgradient(fl[wrt], *args[1:])
    assert got.size == gold.size
    err = np.max(np.abs(got.ravel() - gold)) / np.max(np.abs(gold))
    print(gname, wrt, f"err {err:.2e}")
    assert err <= 2e-9


@pytest.mark.gpu
@pytest.mark.parametrize("pcase", POT_CASES, ids=[c[0] for c in POT_CASES])
def test_scalar_potential_matches_vm_goldens(mb, pcase):
    name, mesh, fn, gradfn, _, _, sel = pcase
    G = np.load(GOLD)
    v, gr = meshes()[mesh]
    m = mb.Mesh(v, edges=gr.get(1), faces=gr.get(2))
    args = [m] + ([mb.Selection(m, ids={0: sel})] if sel is not None else [])
    pot = mb.ScalarPotential(fn, gradfn)
    integrand, total, grad = pot.integrand(*args), pot.total(*args), pot.gradient(*args)
    gi, gt, gg = G[name + "__integrand"], float(G[name + "__total"][0]), G[name + "__gradient"]
    assert np.max(np.abs(integrand - gi)) <= TOL * np.max(np.abs(gi))
    assert abs(total - gt) <= TOL * max(abs(gt), np.max(np.abs(gi)))
    tol = TOL if gradfn is not None else 2e-9          # explicit gradient function: exact; otherwise central differences
    assert np.max(np.abs(grad.ravel() - gg)) <= tol * np.max(np.abs(gg)), name


@pytest.mark.gpu
def test_device_quadrature_large_mesh_vs_live_reference(mb, reference):
    """20k triangles / 30k edges with integrands that need several levels of refinement on part of the mesh."""
    from morpho_b200 import meshgen as mg
    from morpho_b200.integrand import cos, exp, sin
    v, c = mg.icosphere(32)
    e = mg.edges_from_faces(c)
    fd = fields_for(v)
    fn2 = lambda x, q: exp(2 * x[0]) * cos(25 * x[1] * x[2]) + q[0] * q[1]      # noqa: E731
    fn1 = lambda x, q: sin(40 * x[0] * x[1]) / (1.2 + x[2]) + q[2]               # noqa: E731
    m = mb.Mesh(v, edges=e, faces=c)
    q = mb.Field(m, fd["q"])
    rm = reference.mesh(v, {1: e, 2: c})
    for grade, fn, cls in ((2, fn2, mb.AreaIntegral), (1, fn1, mb.LineIntegral)):
        prog = _prog(fn, 3, ["q"], fd)
        want = rm.integral(grade, [tuple(t) for t in prog.tolist()], [fd["q"]])
        got = cls(fn, q).integrand(m)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        print(f"grade {grade}: n={want.size} err {err:.2e}")
        assert err <= TOL
    rm.free()


@pytest.mark.gpu
def test_integral_rejects_bad_programs(mb):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(2)
    m = mb.Mesh(v, faces=c)
    with pytest.raises(mb.MorphoError):
        mb.LineIntegral(lambda x: x[0]).total(m)          # no edges: FnctlELNtFnd like functional.c:115-129
    f = mb.AreaIntegral(lambda x: x[0])
    f._prog = (3, np.zeros(1, dtype=f.program(3).dtype))
    f._prog[1][0] = (3, 0, 0.0)                           # ADD on an empty stack
    with pytest.raises(mb.MorphoError):
        f.total(m)


@pytest.mark.gpu
def test_integral_edge_cases(mb, reference):
    """Empty selections, zero-size elements, a 2-D mesh, an integrand that is singular on part of the mesh: the device
    does what the reference's integrator does (same values, zeros where the reference leaves zeros)."""
    from morpho_b200.integrand import exp, sqrt
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [1, 0, 0], [2, 2, 0.5]], dtype=np.float64)
    faces = np.array([[0, 1, 2], [0, 2, 3], [1, 2, 4], [2, 3, 5]], dtype=np.int32)          # face 2 has zero area (1 and 4 coincide)
    faces.sort(axis=1)
    edges = np.array([[0, 1], [1, 2], [1, 4], [3, 5]], dtype=np.int32)                      # edge 2 has zero length
    m = mb.Mesh(v, edges=edges, faces=faces)
    rm = reference.mesh(v, {1: edges, 2: faces})
    fn = lambda x: exp(x[0]) * sqrt(1.0 + x[1] * x[1]) + x[2]                                # noqa: E731
    for grade, cls in ((1, mb.LineIntegral), (2, mb.AreaIntegral)):
        prog = _prog(fn, 3, [], {})
        want = rm.integral(grade, [tuple(t) for t in prog.tolist()])
        got = cls(fn).integrand(m)
        assert np.max(np.abs(got - want)) <= TOL * np.max(np.abs(want))
        assert got[2] == 0.0                                                                 # the degenerate element contributes nothing
        empty = mb.Selection(m, ids={grade: np.zeros(0, dtype=np.int32)})
        assert cls(fn).total(m, empty) == 0.0 and not np.any(cls(fn).integrand(m, empty))
        assert not np.any(cls(fn).gradient(m, empty))
    rm.free()
    # 2-D mesh: positions have two coordinates, x[2] is out of range for the program validator
    v2 = np.ascontiguousarray(v[:4, :2])
    m2 = mb.Mesh(v2, faces=faces[:2])
    assert abs(mb.AreaIntegral(lambda x: x[0] * x[1]).total(m2) - 0.25) <= 1e-14
    with pytest.raises((IndexError, mb.MorphoError)):                # the reference raises an index error too
        mb.AreaIntegral(lambda x: x[2]).total(m2)
    # an integrand with a (integrable) singularity at a vertex: the adaptive rule refines to its depth limit and returns
    r = mb.AreaIntegral(lambda x: 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + 1e-300)).integrand(m2)
    assert np.all(np.isfinite(r)) and abs(r.sum() - 2.0 * np.log(1.0 + np.sqrt(2.0))) < 1e-3     # int_0^1 int_0^1 1/r = 2 asinh(1)


# ---------------------------------------------------------------- GPU: the rule-table integrator (method dictionaries)
def _method_functional(mb, mcase, ms):
    mname, base, _, method = mcase
    case = next(c for c in CASES if c[0] == base)
    name, mesh, grade, fn, _, flds, sel = case
    v, gr = ms[mesh]
    fd = fields_for(v)
    m = mb.Mesh(v, edges=gr.get(1), faces=gr.get(2), volumes=gr.get(3))
    fl = [mb.Field(m, fd[f]) for f in flds]
    func = {1: mb.LineIntegral, 2: mb.AreaIntegral, 3: mb.VolumeIntegral}[grade](fn, *fl, method=method)
    args = [m] + ([mb.Selection(m, ids={grade: sel})] if sel is not None else [])
    return func, args, fl, sel


@pytest.mark.gpu
@pytest.mark.parametrize("mcase", METHOD_CASES, ids=[c[0] for c in METHOD_CASES])
def test_device_rule_table_integrator_matches_vm_goldens(mb, mcase):
    """integrate() with a method dictionary (integrate.c:2759-2866): rule chains, error rules, the heap of work items and
    Laurie's error sharpening on the device against closures run by the reference VM."""
    gold = np.load(GOLD)[mcase[0]].copy()
    func, args, _, sel = _method_functional(mb, mcase, meshes())
    if sel is not None:
        mask = np.zeros(gold.size, dtype=bool)
        mask[sel] = True
        gold[~mask] = 0.0
    got = func.integrand(*args)
    scale = np.max(np.abs(gold))
    err = np.max(np.abs(got - gold)) / scale
    print(mcase[0], f"integrand err {err:.2e}")
    assert err <= TOL, mcase[0]
    assert abs(func.total(*args) - gold.sum()) <= 1e-11 * max(abs(gold.sum()), scale)


@pytest.mark.gpu
@pytest.mark.parametrize("gcase", METHOD_GRAD_CASES, ids=[f"{n}-{'x' if w < 0 else w}" for n, w in METHOD_GRAD_CASES])
def test_device_rule_table_gradients_match_vm_goldens(mb, gcase):
    mname, wrt = gcase
    mcase = next(c for c in METHOD_CASES if c[0] == mname)
    gold = np.load(GOLD)[f"grad__{mname}__{'x' if wrt < 0 else wrt}"]
    func, args, fl, _ = _method_functional(mb, mcase, meshes())
    got = func.gradient(*args) if wrt < 0 else func.fieldgradient(fl[wrt], *args[1:])
    assert got.size == gold.size
    err = np.max(np.abs(got.ravel() - gold)) / np.max(np.abs(gold))
    print(mname, wrt, f"err {err:.2e}")
    assert err <= 2e-9


@pytest.mark.gpu
def test_rule_table_integrator_hands_hopeless_elements_back(mb):
    """1/x on an interval that starts at 0 (the reference's test/integrate/too_many_subdivisions.morpho raises
    IntgrtrSbdvsns): the device reports the call as unsupported instead of returning a number."""
    from morpho_b200 import MorphoError
    v = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    m = mb.Mesh(v, edges=np.array([[0, 1]], dtype=np.int32))
    with pytest.raises(MorphoError):
        mb.LineIntegral(lambda x: 1.0 / x[0], method={}).total(m)
    assert abs(mb.LineIntegral(lambda x: x[0], method={}).total(m) - 0.5) < 1e-14       # test/integrate/embedding.morpho
