"""Definition of the seeded parity cases shared by the golden-vector generator
(make_golden.py, runs the UNMODIFIED reference) and the tests (oracle and CUDA path).

Every case is a dict: mesh arrays, optional fields, a list of evaluations
(functional spec, method, optional selection, optional field to differentiate).
Inputs are regenerated from the seeds; only reference OUTPUTS are stored in the fixtures.
"""
import numpy as np

from morpho_b200 import meshgen as mg


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def _unit_rows(a):
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def case_sphere(f=4, seed=1):
    v, c = mg.icosphere(f)
    v = mg.perturb_radially(v, 0.05, seed)
    edges = mg.edges_from_faces(c)
    r = _rng(seed + 100)
    sel_faces = np.sort(r.choice(c.shape[0], size=c.shape[0] // 3, replace=False)).astype(np.int32)
    sel_edges = np.sort(r.choice(edges.shape[0], size=edges.shape[0] // 4, replace=False)).astype(np.int32)
    evals = []
    for name in ("Area", "VolumeEnclosed"):
        for method in ("total", "integrand", "gradient"):
            evals.append(dict(f=dict(name=name), method=method))
            evals.append(dict(f=dict(name=name), method=method, sel=(2, sel_faces)))
    for method in ("total", "integrand", "gradient"):
        evals.append(dict(f=dict(name="Length"), method=method))
        evals.append(dict(f=dict(name="Length"), method=method, sel=(1, sel_edges)))
        evals.append(dict(f=dict(name="AreaEnclosed"), method=method))
    # appended after the round-1 evaluations so that their fixture keys keep their numbers
    vsel = np.sort(r.choice(v.shape[0], size=v.shape[0] // 3, replace=False)).astype(np.int32)
    wts = 0.5 + r.uniform(0, 1, size=c.shape[0])
    for method in ("total", "integrand", "gradient"):
        evals.append(dict(f=dict(name="EquiElement"), method=method))
        evals.append(dict(f=dict(name="EquiElement"), method=method, sel=(0, vsel)))
        evals.append(dict(f=dict(name="EquiElement", weight=wts), method=method))
        evals.append(dict(f=dict(name="EquiElement", grade=1), method=method))
    return dict(name=f"sphere_f{f}", vert=v, grades={1: edges, 2: c}, fields={}, evals=evals)


def case_tetcube(n=3, seed=2):
    v, t = mg.tet_cube(n)
    r = _rng(seed)
    v = v + 0.08 / n * r.uniform(-1, 1, size=v.shape)
    q5 = r.standard_normal((v.shape[0], 5))
    nn = _unit_rows(r.standard_normal((v.shape[0], 3)) + np.array([2.0, 0, 0]))
    phi = r.standard_normal(v.shape[0])
    sel = np.sort(r.choice(t.shape[0], size=t.shape[0] // 2, replace=False)).astype(np.int32)
    evals = []
    for method in ("total", "integrand", "gradient"):
        evals.append(dict(f=dict(name="Volume"), method=method))
        evals.append(dict(f=dict(name="Volume"), method=method, sel=(3, sel)))
    for method in ("total", "integrand", "gradient", "fieldgradient"):
        evals.append(dict(f=dict(name="GradSq", fields=["q5"]), method=method, wrt="q5"))
        evals.append(dict(f=dict(name="Nematic", fields=["nn"], ksplay=1.0, ktwist=0.7, kbend=1.3), method=method, wrt="nn"))
        evals.append(dict(f=dict(name="Nematic", fields=["nn"], ksplay=0.9, ktwist=1.1, kbend=1.0, pitch=0.4), method=method, wrt="nn"))
        evals.append(dict(f=dict(name="NematicElectric", fields=["nn", "phi"]), method=method, wrt="nn"))
    evals.append(dict(f=dict(name="NematicElectric", fields=["nn", "phi"]), method="fieldgradient", wrt="phi"))
    evals.append(dict(f=dict(name="GradSq", fields=["q5"]), method="fieldgradient", wrt="q5", sel=(3, sel)))
    for method in ("total", "integrand", "fieldgradient", "gradient"):
        evals.append(dict(f=dict(name="NormSq", fields=["q5"]), method=method, wrt="q5"))
    return dict(name=f"tetcube_n{n}", vert=v, grades={3: t}, fields=dict(q5=q5, nn=nn, phi=phi), evals=evals)


def case_disk(n=5, seed=3, dim=3):
    v, f, b = mg.disk(n, dim)
    r = _rng(seed)
    interior = np.ones(v.shape[0], bool)
    interior[np.unique(b)] = False
    jig = 0.15 / n * r.uniform(-1, 1, size=v.shape)
    if dim == 3:
        jig[:, 2] = 0.0
    v = v + jig * interior[:, None]
    edges = mg.edges_from_faces(f)
    # boundary edge ids inside the full edge list
    key = lambda a: a[:, 0].astype(np.int64) * v.shape[0] + a[:, 1]
    bsel = np.nonzero(np.isin(key(edges), key(b)))[0].astype(np.int32)
    ang = 0.6 * r.uniform(-1, 1, size=v.shape[0])
    nn = np.stack([np.cos(ang), np.sin(ang), 0.2 * r.uniform(-1, 1, size=v.shape[0])], axis=1)
    nn = _unit_rows(nn)
    phi = v[:, 0] * 0.7 + 0.1 * r.standard_normal(v.shape[0])
    q2 = r.standard_normal((v.shape[0], 2))
    fsel = np.sort(r.choice(f.shape[0], size=f.shape[0] // 2, replace=False)).astype(np.int32)
    vsel = np.sort(r.choice(v.shape[0], size=v.shape[0] // 2, replace=False)).astype(np.int32)
    evals = []
    for method in ("total", "integrand", "gradient", "fieldgradient"):
        evals.append(dict(f=dict(name="Nematic", fields=["nn"], ksplay=1.0, ktwist=1.0, kbend=1.0), method=method, wrt="nn"))
        evals.append(dict(f=dict(name="Nematic", fields=["nn"], ksplay=0.5, ktwist=2.0, kbend=1.5, pitch=0.3), method=method, wrt="nn"))
        evals.append(dict(f=dict(name="NematicElectric", fields=["nn", "phi"]), method=method, wrt="nn"))
        evals.append(dict(f=dict(name="GradSq", fields=["q2"]), method=method, wrt="q2"))
        evals.append(dict(f=dict(name="GradSq", fields=["nn"]), method=method, wrt="nn", sel=(2, fsel)))
        evals.append(dict(f=dict(name="NormSq", fields=["nn"]), method=method, wrt="nn"))
    evals.append(dict(f=dict(name="NormSq", fields=["nn"]), method="fieldgradient", wrt="nn", sel=(0, vsel)))
    evals.append(dict(f=dict(name="NormSq", fields=["nn"]), method="total", wrt="nn", sel=(0, vsel)))
    evals.append(dict(f=dict(name="NematicElectric", fields=["nn", "phi"]), method="fieldgradient", wrt="phi"))
    for method in ("total", "integrand", "gradient"):
        evals.append(dict(f=dict(name="Length"), method=method, sel=(1, bsel)))
        evals.append(dict(f=dict(name="Area"), method=method))
    for method in ("total", "integrand", "gradient"):          # the regulariser of examples/tactoid/tactoid.morpho:16
        evals.append(dict(f=dict(name="EquiElement"), method=method))
    return dict(name=f"disk_n{n}_d{dim}", vert=v, grades={1: edges, 2: f}, fields=dict(nn=nn, phi=phi, q2=q2), evals=evals)


def case_loop(n=40, seed=4, dim=3, closed=True):
    v, e = mg.polyline_loop(n, dim, seed)
    if not closed:
        e = e[~((e[:, 0] == 0) & (e[:, 1] == n - 1))]
    r = _rng(seed + 7)
    vsel = np.sort(r.choice(n, size=n // 2, replace=False)).astype(np.int32)
    evals = []
    for method in ("total", "integrand", "gradient"):
        evals.append(dict(f=dict(name="LineCurvatureSq"), method=method))
        evals.append(dict(f=dict(name="LineCurvatureSq", integrandonly=True), method=method))
        evals.append(dict(f=dict(name="LineCurvatureSq"), method=method, sel=(0, vsel)))
        evals.append(dict(f=dict(name="Length"), method=method))
        evals.append(dict(f=dict(name="AreaEnclosed"), method=method))
    return dict(name=f"loop_n{n}_d{dim}_{'closed' if closed else 'open'}", vert=v, grades={1: e}, fields={}, evals=evals)


def all_cases():
    return [case_sphere(4, 1), case_sphere(9, 11), case_tetcube(3, 2), case_disk(5, 3, 3), case_disk(4, 5, 2),
            case_loop(40, 4, 3, True), case_loop(25, 6, 2, True), case_loop(30, 8, 3, False)]


def eval_key(i, ev):
    f = ev["f"]
    s = f"{i:03d}_{f['name']}_{ev['method']}"
    if "sel" in ev:
        s += f"_sel{ev['sel'][0]}"
    if ev.get("wrt") and ev["method"] == "fieldgradient":
        s += f"_wrt_{ev['wrt']}"
    return s
