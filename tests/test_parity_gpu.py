"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes → libmorpho_cuda.so),
against (a) the committed golden vectors of the unmodified reference, (b) the CPU oracle on the same
seeded inputs, and (c) size-independent properties at BASELINE.json's full size.

Tolerances are the ones in tests/parity.py: 1e-12 for energies, integrands and analytic gradients;
2e-9 (of the gradient scale) for the finite-difference family, whose reference values carry
their own ~1e-10 round-off amplified by 1/(2h)."""
import os

import numpy as np
import pytest

from oracle.oracle import FunctionalSpec
from tests.golden.cases import all_cases, eval_key
from tests.parity import ANALYTIC, TOL_EXACT, contribution_norm_sum, measure, record, rel_array, rel_pervertex, tolerance

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = all_cases()


@pytest.fixture(scope="module")
def mb():
    import morpho_b200 as mb
    mb.init(0)
    return mb


def _build(mb, case):
    g = case["grades"]
    mesh = mb.Mesh(case["vert"], edges=g.get(1), faces=g.get(2), volumes=g.get(3))
    fields = {k: mb.Field(mesh, np.asarray(v).reshape(mesh.nv, -1)) for k, v in case["fields"].items()}
    return mesh, fields


def _functional(mb, ev, fields):
    f = dict(ev["f"])
    name = f.pop("name")
    fl = [fields[n] for n in f.pop("fields", [])]
    return getattr(mb, name)(*fl, **f)


def _cuda_eval(mb, mesh, fields, ev):
    func = _functional(mb, ev, fields)
    args = [mesh]
    if "sel" in ev:
        args.append(mb.Selection(mesh, ids={ev["sel"][0]: ev["sel"][1]}))
    m = ev["method"]
    if m == "fieldgradient":
        return func.fieldgradient(fields[ev["wrt"]], *args)
    return getattr(func, m)(*args)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_golden(mb, case):
    gold = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    mesh, fields = _build(mb, case)
    worst = {}
    for i, ev in enumerate(case["evals"]):
        key = eval_key(i, ev)
        got = _cuda_eval(mb, mesh, fields, ev)
        assert got is not None, key
        err = measure(ev, got, gold[key], case)
        name, method = ev["f"]["name"], ev["method"]
        worst[(name, method)] = max(worst.get((name, method), 0.0), err)
        record("golden", f"{name}.{method}", err)
        assert err <= tolerance(ev), f"{case['name']}:{key} err {err:.3e} > {tolerance(ev):.1e}"
        if method == "gradient" and name in ANALYTIC and "sel" not in ev:
            # per-vertex measure (first option; tests/parity.py): sees a wrong row at a low-force vertex
            g = {"Length": 1, "Area": 2, "VolumeEnclosed": 2, "Volume": 3}[name]
            S = contribution_norm_sum(name, case["vert"], case["grades"][g])
            pv = rel_pervertex(got, gold[key], S)
            record("golden_pervertex", f"{name}.{method}", pv)
            assert pv <= TOL_EXACT, f"{case['name']}:{key} per-vertex err {pv:.3e}"
    print({k: f"{v:.1e}" for k, v in worst.items()})


@pytest.mark.parametrize("case", [c for c in CASES if any(ev["f"]["name"] in ANALYTIC for ev in c["evals"])],
                         ids=[c["name"] for c in CASES if any(ev["f"]["name"] in ANALYTIC for ev in c["evals"])])
def test_block_kernel_matches_reference_golden(mb, case):
    """The shared-memory block kernel (mcu_block.cu; gradient_path = MCU_PATH_BLOCK) on every analytic gradient of the
    golden cases — lines, triangles, tetrahedra, 2-D and 3-D, with and without Selections — at 1e-12, norm-wise and per
    vertex, and bit-identical rows where no (selected) element touches a vertex."""
    gold = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    mesh, fields = _build(mb, case)
    mb.set_option("gradient_path", mb._lib.MCU_PATH_BLOCK)
    try:
        n = 0
        for i, ev in enumerate(case["evals"]):
            name = ev["f"]["name"]
            if name not in ANALYTIC or ev["method"] != "gradient":
                continue
            key = eval_key(i, ev)
            got = _cuda_eval(mb, mesh, fields, ev)
            assert got is not None, key
            err = measure(ev, got, gold[key], case)
            record("golden_block", f"{name}.gradient", err)
            assert err <= TOL_EXACT, f"{case['name']}:{key} err {err:.3e}"
            g = {"Length": 1, "Area": 2, "VolumeEnclosed": 2, "Volume": 3}[name]
            conn = case["grades"][g] if "sel" not in ev else case["grades"][g][ev["sel"][1]]
            S = contribution_norm_sum(name, case["vert"], conn)
            pv = rel_pervertex(got, gold[key], S)
            record("golden_block_pervertex", f"{name}.gradient", pv)
            assert pv <= TOL_EXACT, f"{case['name']}:{key} per-vertex err {pv:.3e}"
            n += 1
        assert n > 0
    finally:
        mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)


@pytest.mark.parametrize("f", [1, 7, 40])
def test_transpose_bit_exact(mb, oracle, f):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(f)
    m = mb.Mesh(v, faces=c)
    tptr, tidx = m.transpose(2)
    tptr0, tidx0 = oracle.transpose(v.shape[0], c)
    np.testing.assert_array_equal(tptr, tptr0)
    np.testing.assert_array_equal(tidx, tidx0)


@pytest.mark.parametrize("f", [3, 22, 70, 224])
def test_sphere_rungs_against_oracle(mb, oracle, f):
    """SURVEY §8d config 2 rungs: Area / VolumeEnclosed total, integrand, gradient on both assembly paths."""
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(f)
    v = mg.perturb_radially(v)
    m = mb.Mesh(v, faces=c)
    for cls, name in ((mb.Area, "Area"), (mb.VolumeEnclosed, "VolumeEnclosed")):
        spec = FunctionalSpec(name)
        fn = cls()
        t0 = oracle.total(v, {2: c}, spec)
        assert abs(fn.total(m) - t0) <= TOL_EXACT * abs(t0)
        assert rel_array(fn.integrand(m), oracle.integrand(v, {2: c}, spec)) <= TOL_EXACT
        g0 = oracle.gradient(v, {2: c}, spec)
        S = contribution_norm_sum(name, v, c)
        res = {}
        for path in (mb._lib.MCU_PATH_GATHER, mb._lib.MCU_PATH_PATCH):
            mb.set_option("gradient_path", path)
            g = fn.gradient(m)
            assert rel_array(g, g0, rows=3) <= TOL_EXACT, (name, path)
            pv = rel_pervertex(g, g0, S)
            record("sphere_rungs_pervertex", f"{name}.f{f}.path{path}", pv)
            assert pv <= TOL_EXACT, (name, path, pv)
            g2 = fn.gradient(m)
            np.testing.assert_array_equal(g, g2)     # deterministic: bitwise reproducible run to run
            res[path] = g
        mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
        # the assembly paths agree with each other as well as with the oracle
        assert rel_array(res[mb._lib.MCU_PATH_PATCH], res[mb._lib.MCU_PATH_GATHER], rows=3) <= TOL_EXACT


def test_vertex_update_invalidates_mirror(mb, oracle):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(10)
    m = mb.Mesh(v, faces=c)
    a = mb.Area()
    t1 = a.total(m)
    v2 = mg.perturb_radially(v, 0.1, 5)
    m.setvertexmatrix(v2)
    t2 = a.total(m)
    assert abs(t2 - oracle.total(v2, {2: c}, FunctionalSpec("Area"))) <= TOL_EXACT * t2
    assert t1 != t2


def test_edge_cases(mb, oracle):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(4)
    m = mb.Mesh(v, faces=c)
    # empty selection: totals 0, gradient all zeros, integrand all zeros
    empty = mb.Selection(m, ids={2: []})
    assert mb.Area().total(m, empty) == 0.0
    assert not mb.Area().gradient(m, empty).any()
    assert not mb.Area().integrand(m, empty).any()
    # missing grade → FnctlELNtFnd (functional.c:115-129)
    with pytest.raises(mb.MorphoError) as ei:
        mb.Length().total(m)
    assert ei.value.id == "FnctlELNtFnd"
    with pytest.raises(mb.MorphoError):
        mb.Volume().total(m)
    # degenerate triangle → gradient is nil (functional.c:2090, 412-415); integrand still defined
    vd = v.copy()
    vd[c[0, 1]] = vd[c[0, 0]]
    md = mb.Mesh(vd, faces=c)
    for path in (mb._lib.MCU_PATH_GATHER, mb._lib.MCU_PATH_PATCH):
        mb.set_option("gradient_path", path)
        assert mb.Area().gradient(md) is None
    mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
    assert mb.Area().integrand(md)[0] == 0.0
    # a vertex at the origin → VolEnclZero (functional.c:2148-2151)
    vz = v.copy()
    vz[c[5, 2]] = 0.0
    mz = mb.Mesh(vz, faces=c)
    with pytest.raises(mb.MorphoError) as ei:
        mb.VolumeEnclosed().gradient(mz)
    assert ei.value.id == "VolEnclZero"
    # isolated vertices get zero rows; single-element mesh works
    vi = np.concatenate([v, [[2.0, 2.0, 2.0]]])
    mi = mb.Mesh(vi, faces=c)
    for path in (mb._lib.MCU_PATH_GATHER, mb._lib.MCU_PATH_PATCH):
        mb.set_option("gradient_path", path)
        g = mb.Area().gradient(mi)
        assert not g[-1].any()
        assert rel_array(g[:-1], oracle.gradient(v, {2: c}, FunctionalSpec("Area")), rows=3) <= TOL_EXACT
    mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
    one = mb.Mesh(np.array([[1, 0, 0], [-0.5, 0.866025, 0], [-0.5, -0.866025, 0]], float), faces=[[0, 1, 2]])
    assert abs(mb.Area().integrand(one)[0] - 1.29904) < 1e-5    # test/functionals/area/area.morpho:9-10


def test_selection_boundary_and_addgrade(mb, oracle):
    from morpho_b200 import meshgen as mg
    v, f, b = mg.disk(6, 3)
    m = mb.Mesh(v, faces=f)
    edges = m.addgrade(1)
    np.testing.assert_array_equal(edges, oracle.edges_from(f))           # mesh_addgrade order, bit-exact
    bnd = mb.Selection(m, boundary=True)
    ids = bnd.idlistforgrade(1)
    key = lambda a: a[:, 0].astype(np.int64) * v.shape[0] + a[:, 1]
    assert sorted(key(edges[ids])) == sorted(key(b))
    L = mb.Length().total(m, bnd)
    assert abs(L - oracle.total(v, {1: edges}, FunctionalSpec("Length"), ids)) <= TOL_EXACT * L


def test_pairwise_against_oracle(mb, oracle):
    from morpho_b200 import meshgen as mg
    x = mg.sphere_points(1500, 12345)
    m = mb.Mesh(x)
    p = mb.PairwisePotential("coulomb")
    e0 = oracle.pairwise_coulomb(x, "integrand")
    g0 = oracle.pairwise_coulomb(x, "gradient")
    assert rel_array(p.integrand(m), e0) <= TOL_EXACT
    assert abs(p.total(m) - e0.sum()) <= TOL_EXACT * e0.sum()
    assert rel_array(p.gradient(m), g0, rows=3) <= TOL_EXACT
    pc = mb.PairwisePotential("coulomb", cutoff=0.5)
    assert rel_array(pc.gradient(m), oracle.pairwise_coulomb(x, "gradient", 0.5), rows=3) <= TOL_EXACT


def test_pairwise_against_the_interpreted_reference(mb):
    """tests/golden/pairwise.npz: the reference's interpreted PairwisePotential (modules/functionals.morpho:87-141, real
    closures, run by oracle/_ref/morpho6) at N = 500 / 1000 / 2000 and for power-law, Yukawa, Lennard-Jones, cut-off and
    2-D cases — against every library route to the same potential (named family and postfix programs)."""
    from tests.golden.pairwise_cases import CASES as PW, LIB
    from tests.parity import rel_scalar
    G = np.load(os.path.join(GOLD, "pairwise.npz"))
    worst = {}
    for name, n, *_ in PW:
        x = G[name + "__x"]
        m = mb.Mesh(x)
        for kw in LIB[name]:
            p = mb.PairwisePotential(**kw)
            e = rel_array(p.integrand(m), G[name + "__integrand"])
            g = rel_array(p.gradient(m), G[name + "__gradient"], rows=x.shape[1])
            t = rel_scalar(p.total(m), float(G[name + "__total"][0]))
            worst[f"{name}/{kw['kind']}"] = max(e, g, t)
            assert max(e, g, t) <= TOL_EXACT, (name, kw["kind"], e, g, t)
    print("pairwise vs interpreted reference, worst relative error per route:", {k: f"{v:.1e}" for k, v in worst.items()})


def test_full_size_properties(mb, oracle):
    """BASELINE.json config 2 at full size (icosphere f=707: 9 996 980 triangles).  Size-independent
    properties: Area is homogeneous of degree 2 and VolumeEnclosed of degree 3 in the coordinates
    (Euler: x.grad f = k f); Area is translation invariant (rows sum to 0); the assembly paths agree to
    1e-12; results are bitwise reproducible.  The oracle finishes this size in a few seconds, so energies and
    gradients are also compared with it directly at 1e-12."""
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(707)
    v = mg.perturb_radially(v)
    m = mb.Mesh(v, faces=c)
    assert c.shape[0] == 9996980 and v.shape[0] == 4998492
    a, ve = mb.Area(), mb.VolumeEnclosed()
    A, V = a.total(m), ve.total(m)
    assert abs(A - a.integrand(m).sum()) <= 1e-12 * A
    grads = {}
    for path in (mb._lib.MCU_PATH_GATHER, mb._lib.MCU_PATH_PATCH):
        mb.set_option("gradient_path", path)
        ga, gv = a.gradient(m), ve.gradient(m)
        scale = np.linalg.norm(ga, axis=1).max()
        assert np.abs(ga.sum(axis=0)).max() <= 1e-9 * scale * np.sqrt(len(ga))
        assert abs((ga * v).sum() - 2 * A) <= 1e-11 * A
        assert abs((gv * v).sum() - 3 * V) <= 1e-11 * V
        grads[path] = (ga, gv)
    mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
    (ga1, gv1), (ga2, gv2) = grads[mb._lib.MCU_PATH_GATHER], grads[mb._lib.MCU_PATH_PATCH]
    assert rel_array(ga2, ga1, rows=3) <= TOL_EXACT
    assert rel_array(gv2, gv1, rows=3) <= TOL_EXACT
    np.testing.assert_array_equal(a.gradient(m), ga2)
    for fn, name, tot, g_gather, g_patch in ((a, "Area", A, ga1, ga2), (ve, "VolumeEnclosed", V, gv1, gv2)):
        spec = FunctionalSpec(name)
        t0 = oracle.total(v, {2: c}, spec)
        assert abs(tot - t0) <= TOL_EXACT * abs(t0), name
        g0 = oracle.gradient(v, {2: c}, spec)
        assert rel_array(g_gather, g0, rows=3) <= TOL_EXACT, name
        assert rel_array(g_patch, g0, rows=3) <= TOL_EXACT, name
        # per vertex (first option): on this mesh the net force is ~1.6e-3 of what the six triangles of a
        # vertex contribute, so the global measure above cannot see one wrong row — this one can
        S = contribution_norm_sum(name, v, c)
        for label, g in (("gather", g_gather), ("patch", g_patch)):
            pv = rel_pervertex(g, g0, S)
            record("fullsize_pervertex", f"{name}.{label}", pv)
            assert pv <= TOL_EXACT, (name, label, pv)


def test_lowering_matrices_on_device_match_the_reference_bit_exactly(mb, reference):
    """conn(row, col), row < col, built on the device (mcu_connect.cu: sort + binary search) against
    mesh_addlowermatrix of the unmodified reference (mesh.c:621-660): the same compressed columns, entry for entry.
    (1,2) on a closed surface, (1,2), (1,3) and (2,3) on tetrahedra, and a grade-1 that lacks some edges."""
    from morpho_b200 import meshgen as mg
    v, f = mg.icosphere(9)
    v2, t = mg.tet_cube(4)
    for vert, top, g_top, pairs in ((v, f, 2, ((1, 2),)), (v2, t, 3, ((2, 3), (1, 3), (1, 2)))):
        kw = {"faces": top} if g_top == 2 else {"volumes": top}
        m = mb.Mesh(vert, **kw)
        rm = reference.mesh(vert, {g_top: top})
        for g in range(g_top - 1, 0, -1):
            m.addgrade(g)
            rm.addgrade(g)
        for row, col in pairs:
            cptr, rix, nrows = m.lower(row, col)
            want_cptr, want_rix = rm.conn(row, col, create=True)
            np.testing.assert_array_equal(cptr, want_cptr)
            np.testing.assert_array_equal(rix, want_rix)
            assert nrows == m.count(row)
        rm.free()
    # edges given explicitly, every third one missing: columns of different length
    edges = mg.edges_from_faces(f)[::3].copy()
    m = mb.Mesh(v, edges=edges, faces=f)
    rm = reference.mesh(v, {1: edges, 2: f})
    cptr, rix, _ = m.lower(1, 2)
    want_cptr, want_rix = rm.conn(1, 2, create=True)
    np.testing.assert_array_equal(cptr, want_cptr)
    np.testing.assert_array_equal(rix, want_rix)
    assert len(set(np.diff(cptr).tolist())) > 1
    rm.free()


def test_addgrade_on_device_matches_the_reference_bit_exactly(mb, oracle, reference):
    """Mesh.addgrade built on the device (mcu_connect.cu: sort + scan) against mesh_addgrade of the unmodified
    reference (mesh.c:419-489): element ids in the order of first discovery, bit-exact.  Edges from triangles, edges
    and faces from tetrahedra; then the headline mesh against the oracle's restatement."""
    import time
    from morpho_b200 import meshgen as mg
    v, f = mg.icosphere(12)
    v2, t = mg.tet_cube(5)
    for vert, top, g_top, grades in ((v, f, 2, (1,)), (v2, t, 3, (2, 1)), (v2, t, 3, (1,))):
        kw = {"faces": top} if g_top == 2 else {"volumes": top}
        m = mb.Mesh(vert, **kw)
        rm = reference.mesh(vert, {g_top: top})
        for g in grades:
            got = m.addgrade(g)
            rm.addgrade(g)
            cptr, rix = rm.conn(0, g, create=False)
            assert np.array_equal(cptr, np.arange(got.shape[0] + 1) * (g + 1))
            np.testing.assert_array_equal(got, rix.reshape(-1, g + 1))
        rm.free()
    v, f = mg.icosphere(707)
    m = mb.Mesh(v, faces=f)
    t0 = time.time()
    edges = m.addgrade(1)
    t_dev = time.time() - t0
    assert edges.shape[0] == 30 * 707 * 707
    t0 = time.time()
    want = oracle.edges_from(f)
    print(f"addgrade(1) on 9 996 980 triangles: device {t_dev:.2f}s (incl. read-back), oracle (sort based, C) {time.time() - t0:.2f}s")
    np.testing.assert_array_equal(edges, want)


def test_patch_path_on_irregular_meshes(mb, oracle):
    """The patch pipeline (gradient AND total) away from the closed, nicely numbered sphere: a surface with boundary
    (open fans), a non-manifold edge (three sheets meeting), randomly permuted vertex numbering (runs of one or two
    vertices), isolated vertices, and the error statuses raised from inside the patch kernel."""
    from morpho_b200 import meshgen as mg
    mb.set_option("patch_min_elements", 0)
    mb.set_option("gradient_path", mb._lib.MCU_PATH_PATCH)
    try:
        rng = np.random.Generator(np.random.PCG64(99))
        meshes = []
        v, f, _ = mg.disk(40, 3)                                   # boundary: open fans
        v = v + 0.01 * rng.standard_normal(v.shape) + np.array([0.3, 0.2, 1.0])
        meshes.append(("disk", v, f))
        # three rectangular sheets sharing one edge line (non-manifold: 3 triangles per interior edge of the spine)
        n = 24
        sheets, vs = [], []
        for k, ang in enumerate((0.0, 2.1, 4.0)):
            u, w = np.meshgrid(np.arange(n + 1) / n, np.arange(1, n + 1) / n, indexing="ij")
            vs.append(np.stack([w.ravel() * np.cos(ang), w.ravel() * np.sin(ang), u.ravel()], axis=1))
        spine = np.stack([np.zeros(n + 1), np.zeros(n + 1), np.arange(n + 1) / n], axis=1)
        vert = np.concatenate([spine] + vs) + np.array([1.0, 1.0, 0.5])
        tris = []
        for k in range(3):
            base = (n + 1) + k * (n + 1) * n
            idx = lambda i, j: np.where(j == 0, i, base + i * n + (j - 1))     # column 0 is the shared spine
            I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
            I, J = I.ravel(), J.ravel()
            a, b, c2, d = idx(I, J), idx(I + 1, J), idx(I, J + 1), idx(I + 1, J + 1)
            tris += [np.stack([a, b, c2], axis=1), np.stack([b, d, c2], axis=1)]
        tri = np.sort(np.concatenate(tris).astype(np.int32), axis=1)
        meshes.append(("book", vert + 0.003 * rng.standard_normal(vert.shape), tri))
        vs2, fs2 = mg.icosphere(30)                                # random numbering: no consecutive-id locality
        vs2 = mg.perturb_radially(vs2)
        perm = rng.permutation(vs2.shape[0])
        inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
        meshes.append(("permuted sphere", vs2[perm], np.sort(inv[fs2].astype(np.int32), axis=1)))
        vi = np.concatenate([vs2, [[2.0, 2.0, 2.0], [3.0, -1.0, 0.5]]])     # isolated vertices
        meshes.append(("isolated", vi, fs2))
        for name, vv, ff in meshes:
            vv, ff = np.ascontiguousarray(vv), np.ascontiguousarray(ff)
            if name == "permuted sphere":
                # by default such a numbering is refused by the patch planner (hundreds of two-vertex copies per patch)
                # and AUTO falls back to the gather kernel ...
                with pytest.raises(mb.MorphoError):
                    mb.Area().gradient(mb.Mesh(vv, faces=ff))
                mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
                g = mb.Area().gradient(mb.Mesh(vv, faces=ff))
                assert rel_array(g, oracle.gradient(vv, {2: ff}, FunctionalSpec("Area")), rows=3) <= TOL_EXACT
                mb.set_option("gradient_path", mb._lib.MCU_PATH_PATCH)
                # ... but the kernel itself must cope: force it (patches with more than 64 runs, the overflow path)
                mb.set_option("patch_max_avg_runs", 1000000)
            m = mb.Mesh(vv, faces=ff)
            for cls, fname in ((mb.Area, "Area"), (mb.VolumeEnclosed, "VolumeEnclosed")):
                spec = FunctionalSpec(fname)
                g0 = oracle.gradient(vv, {2: ff}, spec)
                g = cls().gradient(m)
                assert rel_array(g, g0, rows=3) <= TOL_EXACT, (name, fname)
                t0 = oracle.total(vv, {2: ff}, spec)
                assert abs(cls().total(m) - t0) <= TOL_EXACT * abs(t0), (name, fname)
        # statuses from inside the patch kernel
        vz = vs2.copy(); vz[fs2[100, 2]] = 0.0
        with pytest.raises(mb.MorphoError) as ei:
            mb.VolumeEnclosed().gradient(mb.Mesh(vz, faces=fs2))
        assert ei.value.id == "VolEnclZero"
        vd = vs2.copy(); vd[fs2[7, 1]] = vd[fs2[7, 0]]
        assert mb.Area().gradient(mb.Mesh(vd, faces=fs2)) is None
    finally:
        mb.set_option("patch_min_elements", 4096)
        mb.set_option("patch_max_avg_runs", 160)
        mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)


def test_fd_field_gradient_paths_are_bit_identical(mb):
    """The element-centric finite-difference kernels and the GradSq specialisation (geometry factored once, only the
    perturbed component re-solved) perform the reference's arithmetic on every number they recompute: their output
    equals the vertex-centric kernels' bit for bit."""
    from morpho_b200 import meshgen as mg
    rng = np.random.Generator(np.random.PCG64(99))
    v, tets = mg.tet_cube(6)
    inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
    v = v + 0.02 * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
    mt = mb.Mesh(v, volumes=tets)
    vs, faces = mg.icosphere(6)
    ms = mb.Mesh(mg.perturb_radially(vs), faces=faces)
    try:
        for mesh in (mt, ms):
            for ps in (1, 3, 5, 4):
                f = mb.Field(mesh, rng.standard_normal((mesh.nv, ps)))
                funcs = [mb.GradSq(f)] + ([mb.Nematic(f, ksplay=1.0, ktwist=0.7, kbend=1.3)] if ps == 3 else [])
                for func in funcs:
                    res = []
                    for elementwise, fast in ((0, 0), (1, 0), (1, 1)):
                        mb.set_option("fd_elementwise", elementwise)
                        mb.set_option("gradsq_fast_fieldgradient", fast)
                        res.append(func.fieldgradient(f, mesh))
                    np.testing.assert_array_equal(res[0], res[1])
                    np.testing.assert_array_equal(res[0], res[2])
    finally:
        mb.set_option("fd_elementwise", -1)
        mb.set_option("gradsq_fast_fieldgradient", 1)


def _boundary_host(nv, top, sub):
    """numpy restatement of selection_selectboundary for the checker: facets contained in exactly one top element"""
    from itertools import combinations
    k = sub.shape[1]

    def keys(a):
        out = np.zeros(a.shape[0], dtype=object)
        for c in range(a.shape[1]):
            out = out * int(nv) + a[:, c].astype(object)
        return out
    cand = np.concatenate([top[:, list(c)] for c in combinations(range(top.shape[1]), k)])
    uk, cnt = np.unique(keys(cand), return_counts=True)
    single = set(uk[cnt == 1].tolist())
    return np.array([i for i, key in enumerate(keys(sub).tolist()) if key in single], dtype=np.int32)


def test_boundary_selection_on_device(mb):
    """Selection(mesh, boundary=true) for surfaces (edges of triangles), volumes (faces of tetrahedra) and lines
    (end points), against a numpy restatement of selection.c:211-232; a closed surface has no boundary."""
    from morpho_b200 import meshgen as mg
    v, f, b = mg.disk(9, 3)
    m = mb.Mesh(v, faces=f)
    s = mb.Selection(m, boundary=True)
    np.testing.assert_array_equal(s.idlistforgrade(1), _boundary_host(m.nv, f, m.addgrade(1)))
    assert len(s.idlistforgrade(1)) == b.shape[0]
    np.testing.assert_array_equal(s.idlistforgrade(0), np.unique(b))
    vs, fs = mg.icosphere(5)
    assert len(mb.Selection(mb.Mesh(vs, faces=fs), boundary=True).idlistforgrade(1)) == 0
    vt, t = mg.tet_cube(4)
    mt = mb.Mesh(vt, volumes=t)
    st = mb.Selection(mt, boundary=True)
    faces = mt.addgrade(2)
    np.testing.assert_array_equal(st.idlistforgrade(2), _boundary_host(mt.nv, t, faces))
    assert len(st.idlistforgrade(2)) == 6 * 2 * 4 * 4                       # two triangles per boundary cell face
    vl, e = mg.polyline_loop(12, 3, 3, 0.1)
    ml = mb.Mesh(vl, edges=e[:-1])                                           # open polyline: two end points
    assert sorted(mb.Selection(ml, boundary=True).idlistforgrade(0).tolist()) == sorted(
        [int(i) for i in np.nonzero(np.bincount(e[:-1].ravel(), minlength=12) == 1)[0]])
