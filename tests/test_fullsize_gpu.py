"""BASELINE.json configs 3-5 at (or near) full size, through size-independent properties of the domain and
sampled comparisons with the CPU oracle.  (Config 2 at full size: tests/test_parity_gpu.py::test_full_size_properties.)

  config 3  200 000 points on S^2, all-pairs Coulomb (PairwisePotential, modules/functionals.morpho:87-141)
  config 4  ~4M-triangle disk, director field: Nematic / Area / NormSq / Length on the boundary selection
  config 5  tetrahedral cube (204^3 cells = 50 937 984 tets): Volume, GradSq on a 5-component field, Nematic

Tolerances: 1e-12 where the quantity is evaluated directly; the finite-difference gradients of the reference
(functional.c:1142-1162, 1305-1354) carry the truncation error of the stencil, h = cbrt(eps) ~ 6e-6, so identities
that hold for the exact derivative hold to ~h^2 = 4e-11 x conditioning: written as 1e-7 in the tests.
"""
import time

import numpy as np
import pytest

from oracle.oracle import FunctionalSpec
from tests.parity import TOL_EXACT, rel_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import morpho_b200 as mb
    mb.init(0)
    return mb


def test_config3_pairwise_200k(mb, oracle):
    from morpho_b200 import meshgen as mg
    n = 200_000
    x = mg.sphere_points(n)
    m = mb.Mesh(x)
    pp = mb.PairwisePotential("coulomb")
    t0 = time.time()
    e = pp.integrand(m)
    g = pp.gradient(m)
    E = pp.total(m)
    print(f"pairwise n={n}: 3 all-pairs evaluations in {time.time() - t0:.2f}s")
    assert abs(e.sum() - E) <= 1e-12 * E
    # Newton's third law and Euler's theorem for a potential homogeneous of degree -1: sum_i x_i.g_i = -E
    assert np.abs(g.sum(axis=0)).max() <= 1e-9 * np.abs(g).max() * np.sqrt(n)
    assert abs((g * x).sum() + E) <= 1e-11 * E
    # scaling x -> 2x: E -> E/2
    m2 = mb.Mesh(2.0 * x)
    assert abs(pp.total(m2) - 0.5 * E) <= 1e-12 * E
    # sampled rows against a direct evaluation: e_i = 1/2 sum_j 1/r_ij, g_i = - sum_j (x_i - x_j)/r^3
    idx = np.array([0, 1, 777, 99_999, 123_456, n - 1])
    for i in idx:
        d = x[i] - x
        r = np.sqrt((d * d).sum(axis=1))
        r[i] = np.inf
        assert abs(e[i] - 0.5 * (1.0 / r).sum()) <= 1e-12 * e[i]
        gi = -(d / r[:, None] ** 3).sum(axis=0)
        assert np.linalg.norm(g[i] - gi) <= 1e-11 * np.linalg.norm(gi)


def _rotz(a, th):
    c, s = np.cos(th), np.sin(th)
    out = a.copy()
    out[:, 0], out[:, 1] = c * a[:, 0] - s * a[:, 1], s * a[:, 0] + c * a[:, 1]
    return out


def test_config4_tactoid_disk_4m(mb, oracle):
    from morpho_b200 import meshgen as mg
    n = 816                                      # 6 n^2 = 3 995 136 triangles, 3 n (n+1) + 1 = 2 000 017 vertices
    v, faces, bnd = mg.disk(n, dim=3)
    rng = np.random.Generator(np.random.PCG64(12345))
    v[:, :2] *= 1.0 + 0.05 * np.sin(3.0 * np.arctan2(v[:, 1], v[:, 0]))[:, None]     # not a perfect circle
    th = 0.3 * rng.uniform(-1, 1, size=v.shape[0]) + 0.5 * v[:, 0]
    nn = np.stack([np.cos(th), np.sin(th), 0.1 * rng.uniform(-1, 1, size=v.shape[0])], axis=1)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    m = mb.Mesh(v, edges=bnd, faces=faces)
    assert faces.shape[0] == 3_995_136
    fld = mb.Field(m, nn)
    nem = mb.Nematic(fld, ksplay=1.0, ktwist=0.6, kbend=1.4)
    E = nem.total(m)
    ei = nem.integrand(m)
    assert abs(ei.sum() - E) <= 1e-12 * E
    # oracle on the full mesh: energy and integrand (plain C, a second or two)
    spec = FunctionalSpec("Nematic", fields=[nn], ksplay=1.0, ktwist=0.6, kbend=1.4)
    conn = {1: bnd, 2: faces}
    assert abs(E - oracle.total(v, conn, spec)) <= TOL_EXACT * E
    assert rel_array(ei, oracle.integrand(v, conn, spec)) <= TOL_EXACT
    # splay (div n)^2 is quadratic in the director, twist (n.curl n)^2 and bend |n x curl n|^2 are quartic:
    # E(a n) = a^2 E_s + a^4 E_tb, and Euler's theorem for the field gradient: sum n.dE/dn = 2 E_s + 4 E_tb
    E_s = mb.Nematic(fld, ksplay=1.0, ktwist=0.0, kbend=0.0).total(m)
    E_tb = E - E_s
    fg = nem.fieldgradient(fld, m)
    assert abs((fg * nn).sum() - (2.0 * E_s + 4.0 * E_tb)) <= 1e-7 * E
    fld.assign(1.5 * nn)
    assert abs(nem.total(m) - (2.25 * E_s + 5.0625 * E_tb)) <= 1e-11 * E
    fld.assign(nn)
    # rigid rotation of mesh and director about z leaves the energy unchanged
    m_rot = mb.Mesh(_rotz(v, 0.7), edges=bnd, faces=faces)
    f_rot = mb.Field(m_rot, _rotz(nn, 0.7))
    assert abs(mb.Nematic(f_rot, ksplay=1.0, ktwist=0.6, kbend=1.4).total(m_rot) - E) <= 1e-11 * E
    # sampled rows of the finite-difference field gradient against the oracle's restatement of the reference stencil
    sub_faces = faces[:20000]
    gids = np.unique(sub_faces)
    loc = np.searchsorted(gids, sub_faces).astype(np.int32)
    loc.sort(axis=1)
    msub = mb.Mesh(v[gids], faces=loc)
    fsub = mb.Field(msub, nn[gids])
    got = mb.Nematic(fsub, ksplay=1.0, ktwist=0.6, kbend=1.4).fieldgradient(fsub, msub)
    want = oracle.fieldgradient(v[gids], {2: loc}, FunctionalSpec("Nematic", fields=[nn[gids]], ksplay=1.0, ktwist=0.6, kbend=1.4), 0)
    assert rel_array(got, want, rows=3) <= 1e-9
    # boundary length on a Selection, area, NormSq
    L = mb.Length()
    sel = mb.Selection(m, ids={1: np.arange(bnd.shape[0], dtype=np.int32)})
    assert abs(L.total(m, sel) - L.total(m)) <= 1e-12 * L.total(m)
    assert abs(L.total(m) - oracle.total(v, conn, FunctionalSpec("Length"))) <= TOL_EXACT * L.total(m)
    A = mb.Area().total(m)
    assert abs(A - oracle.total(v, conn, FunctionalSpec("Area"))) <= TOL_EXACT * A
    ns = mb.NormSq(fld)
    assert abs(ns.total(m) - m.nv) <= 1e-10 * m.nv            # unit director: sum |n|^2 = number of vertices
    # the closure integrals of config 4 at full size (anchoring on the boundary, bulk terms on the sheet), through
    # properties that do not need the CPU: |t| = 1 and |n| = 1 make the integrals the length / the area; the P1
    # gradient of a linear field is exact; and the integrand map sums to the total
    from morpho_b200.integrand import grad, normal, tangent
    li = mb.LineIntegral(lambda x: tangent()[0] ** 2 + tangent()[1] ** 2 + tangent()[2] ** 2)
    assert abs(li.total(m) - L.total(m)) <= 1e-12 * L.total(m)
    anch = mb.LineIntegral(lambda x, q: (q[0] * tangent()[0] + q[1] * tangent()[1] + q[2] * tangent()[2]) ** 2, fld)
    ai = anch.integrand(m)
    assert ai.shape == (bnd.shape[0],) and np.all(ai >= 0) and np.all(ai <= L.integrand(m) * (1 + 1e-12))
    assert abs(ai.sum() - anch.total(m)) <= 1e-12 * anch.total(m)
    one = mb.AreaIntegral(lambda x: normal()[0] ** 2 + normal()[1] ** 2 + normal()[2] ** 2)
    assert abs(one.total(m) - A) <= 1e-12 * A
    lin = mb.Field(m, (2.0 * v[:, 0] - 3.0 * v[:, 1] + 0.5).reshape(-1, 1))
    gsq = mb.AreaIntegral(lambda x, p: grad(p)[0] ** 2 + grad(p)[1] ** 2 + grad(p)[2] ** 2, lin)
    assert abs(gsq.total(m) - 13.0 * A) <= 1e-9 * A                     # flat sheet: |grad|^2 = 4 + 9
    assert abs(gsq.total(m) - mb.GradSq(lin).total(m)) <= 1e-11 * A     # the same number through the GradSq functional


def test_config5_tet_cube_50m(mb, oracle):
    from morpho_b200 import meshgen as mg
    n = 204
    t0 = time.time()
    v, tets = mg.tet_cube(n)
    assert tets.shape[0] == 50_937_984 and v.shape[0] == 205 ** 3
    rng = np.random.Generator(np.random.PCG64(12345))
    inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
    v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]      # jiggle interior vertices only
    m = mb.Mesh(v, volumes=tets)
    print(f"tet cube n={n}: mesh built and mirrored in {time.time() - t0:.1f}s")
    vol = mb.Volume()
    V = vol.total(m)
    assert abs(V - 1.0) <= 1e-12                       # the jiggle moves no boundary vertex: the cube keeps volume 1
    g = vol.gradient(m)
    # interior vertices do not change the total volume; translation invariance; Euler (degree 3)
    assert np.abs(g[inner]).max() <= 1e-12 * np.abs(g).max() * 10
    assert np.abs(g.sum(axis=0)).max() <= 1e-9 * np.abs(g).max()
    assert abs((g * v).sum() - 3.0 * V) <= 1e-11
    # the gradient above came from the block kernel (mcu_block.cu: AUTO takes it for large meshes); the CSR-transpose gather
    # visits the same elements in the same order with the reference's own formulas
    mb.set_option("gradient_path", mb._lib.MCU_PATH_GATHER)
    g_gather = vol.gradient(m)
    mb.set_option("gradient_path", mb._lib.MCU_PATH_AUTO)
    assert rel_array(g, g_gather, rows=3) <= TOL_EXACT
    # rows of boundary vertices are O(1/n^2): compare them vertex by vertex
    bnd = ~inner
    assert (np.linalg.norm(g[bnd] - g_gather[bnd], axis=1) / np.linalg.norm(g_gather[bnd], axis=1)).max() <= 1e-11
    del g_gather
    vi = vol.integrand(m)
    assert abs(vi.sum() - V) <= 1e-12
    # a 5-component field: GradSq is quadratic in q, zero for constants, exact for linear fields
    q_lin = np.stack([v[:, 0], 2 * v[:, 1], -v[:, 2], v[:, 0] + v[:, 1], 0 * v[:, 0] + 3.0], axis=1)
    fld = mb.Field(m, q_lin)
    gs = mb.GradSq(fld)
    # |grad q|_F^2 = 1 + 4 + 1 + 2 + 0 = 8, integrated over the unit cube
    assert abs(gs.total(m) - 8.0) <= 1e-10
    q = q_lin + 0.01 * rng.standard_normal(q_lin.shape)
    fld.assign(q)
    E = gs.total(m)
    fg = gs.fieldgradient(fld, m)
    assert abs((fg * q).sum() - 2.0 * E) <= 1e-7 * E
    assert np.abs(fg.sum(axis=0)).max() <= 1e-7 * np.abs(fg).max() * np.sqrt(m.nv)   # constants cost nothing
    # Nematic on a 3-component director, sampled against the oracle on the first cells
    nn = np.stack([np.cos(v[:, 0] + 0.3 * v[:, 1]), np.sin(v[:, 0] + 0.3 * v[:, 1]), 0.2 * np.sin(4 * v[:, 2])], axis=1)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    fn = mb.Field(m, nn)
    nem = mb.Nematic(fn, ksplay=1.0, ktwist=0.8, kbend=1.2)
    En = nem.total(m)
    assert abs(nem.integrand(m).sum() - En) <= 1e-12 * En
    En_s = mb.Nematic(fn, ksplay=1.0, ktwist=0.0, kbend=0.0).total(m)
    fgn = nem.fieldgradient(fn, m)
    # (the reference's fixed finite-difference step makes this identity degrade like n^2 with the resolution of the
    #  cube: the CPU oracle gives 8e-10, 7e-9, 3e-8 at n = 4, 12, 24 — i.e. ~2e-6 at n = 204; parity with the oracle's
    #  restatement of the same stencil is asserted below at 1e-9)
    assert abs((fgn * nn).sum() - (2.0 * En_s + 4.0 * (En - En_s))) <= 1e-5 * En
    sub = tets[:30000]
    gids = np.unique(sub)
    loc = np.searchsorted(gids, sub).astype(np.int32)
    loc.sort(axis=1)
    msub = mb.Mesh(v[gids], volumes=loc)
    spec = FunctionalSpec("Nematic", fields=[nn[gids]], ksplay=1.0, ktwist=0.8, kbend=1.2)
    fsub = mb.Field(msub, nn[gids])
    nsub = mb.Nematic(fsub, ksplay=1.0, ktwist=0.8, kbend=1.2)
    assert abs(nsub.total(msub) - oracle.total(v[gids], {3: loc}, spec)) <= TOL_EXACT * nsub.total(msub)
    assert rel_array(nsub.integrand(msub), oracle.integrand(v[gids], {3: loc}, spec)) <= TOL_EXACT
    assert rel_array(nsub.fieldgradient(fsub, msub), oracle.fieldgradient(v[gids], {3: loc}, spec, 0), rows=3) <= 1e-9
