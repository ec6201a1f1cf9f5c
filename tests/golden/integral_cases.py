"""Seeded cases for LineIntegral / AreaIntegral / VolumeIntegral: each integrand is written twice, as a Python callable for the
device path (recorded by morpho_b200.integrand.trace) and as the Morpho closure the unmodified VM runs when the
golden file is made.  Fields are polynomials of the vertex position so that both sides build the same values."""
import numpy as np

from morpho_b200 import meshgen as mg
from morpho_b200.integrand import cos, exp, grad, log, normal, sin, sqrt, tangent


def _fields(vert):
    x, y = vert[:, 0], vert[:, 1]
    z = vert[:, 2] if vert.shape[1] > 2 else np.zeros_like(x)
    phi = x * x - 0.5 * y + 0.25 * z
    q = np.stack([x + 0.5 * z, y * y, x * y - z], axis=1)
    return {"phi": phi.reshape(-1, 1), "q": q}


# Morpho source of the same fields (Field(mesh, fn (x,y,z) ...)); 2D meshes get fn (x,y)
FIELD_SRC = {
    3: {"phi": "fn (x,y,z) x*x-0.5*y+0.25*z", "q": "fn (x,y,z) Matrix([x+0.5*z, y*y, x*y-z])"},
    2: {"phi": "fn (x,y) x*x-0.5*y+0.25*0", "q": "fn (x,y) Matrix([x+0.5*0, y*y, x*y-0])"},
}


def meshes():
    out = {}
    v, c = mg.icosphere(3)
    v = mg.perturb_radially(v, 0.02, 5)
    out["sphere"] = (v, {1: mg.edges_from_faces(c), 2: c})
    v, c, _ = mg.disk(5, 2)
    out["disk2d"] = (v, {1: mg.edges_from_faces(c), 2: c})
    v, e = mg.polyline_loop(24, 3, 11, 0.1)
    out["loop"] = (v, {1: e})
    v, t = mg.tet_cube(2)
    rng = np.random.Generator(np.random.PCG64(17))
    v = v + 0.05 * rng.uniform(-1, 1, size=v.shape)
    out["tets"] = (v, {3: t})
    return out


# name, mesh, grade, python callable, morpho closure, fields, selection ids or None
VOLUME_CASES = [
    ("vol_const", "tets", 3, lambda x: 1.0 + 0 * x[0], "fn (x) 1+0*x[0]", [], None),
    ("vol_poly", "tets", 3, lambda x: x[0] * x[1] * x[2] + x[2] ** 3 - 0.5 * x[0] ** 2, "fn (x) x[0]*x[1]*x[2]+x[2]^3-0.5*x[0]^2", [], None),
    ("vol_osc", "tets", 3, lambda x: exp(x[0]) * cos(5 * x[1]) + sin(4 * x[2]) ** 2, "fn (x) exp(x[0])*cos(5*x[1])+sin(4*x[2])^2", [], None),
    ("vol_peak", "tets", 3, lambda x: 1.0 / (0.01 + (x[0] - 0.3) ** 2 + (x[1] - 0.6) ** 2 + x[2] ** 2),
     "fn (x) 1/(0.01+(x[0]-0.3)^2+(x[1]-0.6)^2+x[2]^2)", [], None),
    ("vol_field", "tets", 3, lambda x, p, q: p * p + q[0] * q[1] - q[2] * x[2], "fn (x, p, q) p*p+q[0]*q[1]-q[2]*x[2]", ["phi", "q"], None),
    ("vol_sel", "tets", 3, lambda x, q: sqrt(1.0 + q[1]) * exp(-x[2]), "fn (x, q) sqrt(1+q[1])*exp(-x[2])", ["q"], [0, 5, 6, 17, 30, 47]),
    ("vol_gradsq", "tets", 3, lambda x, p: grad(p)[0] ** 2 + grad(p)[1] ** 2 + grad(p)[2] ** 2 + p * grad(p)[2],
     "fn (x, p) grad(p).inner(grad(p))+p*grad(p)[2]", ["phi"], None),
    ("vol_gradvec", "tets", 3, lambda x, q: sum(grad(q)[k][c] ** 2 for k in range(3) for c in range(3)) + grad(q)[2][1] * q[0],
     "fn (x, q) { var g = grad(q); return g[0].inner(g[0])+g[1].inner(g[1])+g[2].inner(g[2])+g[2][1]*q[0] }", ["q"], None),
]

CASES = [
    ("line_poly", "loop", 1, lambda x: x[0] ** 2 * x[1] + 1.0, "fn (x) x[0]^2*x[1]+1", [], None),
    ("line_osc", "loop", 1, lambda x: sin(9 * x[0]) * exp(x[1]) + cos(7 * x[2]), "fn (x) sin(9*x[0])*exp(x[1])+cos(7*x[2])", [], None),
    ("line_sqrt", "sphere", 1, lambda x: sqrt(x[0] ** 2 + 0.01) / (1.5 + x[2]), "fn (x) sqrt(x[0]^2+0.01)/(1.5+x[2])", [], None),
    ("line_field", "sphere", 1, lambda x, p, q: p * q[1] - q[2] * x[0] + abs(q[0]), "fn (x, p, q) p*q[1]-q[2]*x[0]+abs(q[0])", ["phi", "q"], None),
    ("line_sel", "loop", 1, lambda x, p: log(2.0 + p * p) * x[1], "fn (x, p) log(2+p*p)*x[1]", ["phi"], [0, 3, 4, 11, 23]),
    ("line_2d", "disk2d", 1, lambda x: exp(-4 * (x[0] ** 2 + x[1] ** 2)), "fn (x) exp(-4*(x[0]^2+x[1]^2))", [], None),
    ("area_const", "sphere", 2, lambda x: 1.0 + 0 * x[0], "fn (x) 1+0*x[0]", [], None),
    ("area_quad", "sphere", 2, lambda x: x[0] * x[1] + x[2] ** 2, "fn (x) x[0]*x[1]+x[2]^2", [], None),
    ("area_osc", "sphere", 2, lambda x: exp(x[0]) * cos(6 * x[1]) + sin(5 * x[2]) ** 2, "fn (x) exp(x[0])*cos(6*x[1])+sin(5*x[2])^2", [], None),
    ("area_peak", "disk2d", 2, lambda x: 1.0 / (0.02 + (x[0] - 0.3) ** 2 + x[1] ** 2), "fn (x) 1/(0.02+(x[0]-0.3)^2+x[1]^2)", [], None),
    ("area_field", "sphere", 2, lambda x, p, q: p * p + q[0] * q[1] - q[2] * x[2], "fn (x, p, q) p*p+q[0]*q[1]-q[2]*x[2]", ["phi", "q"], None),
    ("area_sel", "sphere", 2, lambda x, q: sqrt(1.0 + q[1]) * exp(-x[2]), "fn (x, q) sqrt(1+q[1])*exp(-x[2])", ["q"], [1, 2, 17, 40, 41, 99, 150]),
    # element helpers: tangent(), normal(), grad(q)
    ("line_tangent", "loop", 1, lambda x: (tangent()[0] * x[1] - tangent()[2]) ** 2 + tangent()[1],
     "fn (x) (tangent()[0]*x[1]-tangent()[2])^2+tangent()[1]", [], None),
    ("area_anchor", "sphere", 2, lambda x, q: (q[0] * normal()[0] + q[1] * normal()[1] + q[2] * normal()[2]) ** 2 + normal()[2] * x[0],
     "fn (x, q) q.inner(normal())^2+normal()[2]*x[0]", ["q"], None),
    ("area_gradsq", "sphere", 2, lambda x, p: grad(p)[0] ** 2 + grad(p)[1] ** 2 + grad(p)[2] ** 2 + p * grad(p)[0],
     "fn (x, p) grad(p).inner(grad(p))+p*grad(p)[0]", ["phi"], None),
    ("area_gradvec", "sphere", 2, lambda x, q: sum(grad(q)[k][c] ** 2 for k in range(3) for c in range(3)) + grad(q)[1][0] * q[2],
     "fn (x, q) { var g = grad(q); return g[0].inner(g[0])+g[1].inner(g[1])+g[2].inner(g[2])+g[1][0]*q[2] }", ["q"], None),
] + VOLUME_CASES


# The rule-table integrator (`method` dictionary): name, base case, Morpho source of the dictionary, the same as a Python dict.
# Covers the default rule of every grade, rules by name with an extension chain, a rule whose error estimate comes from
# a second rule (keast4 -> keast5), selection by degree, and adapt = false.
METHOD_CASES = [
    ("m_line_default", "line_osc", "{}", {}),
    ("m_line_gauss7", "line_sqrt", '{ "rule" : "gauss7" }', {"rule": "gauss7"}),
    ("m_line_gauss1", "line_osc", '{ "rule" : "gauss1" }', {"rule": "gauss1"}),
    ("m_line_degree", "line_field", '{ "degree" : 6 }', {"degree": 6}),
    ("m_line_noadapt", "line_poly", '{ "rule" : "gauss2", "adapt" : false }', {"rule": "gauss2", "adapt": False}),
    ("m_line_sel", "line_sel", "{}", {}),
    ("m_area_default", "area_osc", "{}", {}),
    ("m_area_peak", "area_peak", "{}", {}),
    ("m_area_tri4", "area_quad", '{ "rule" : "tri4" }', {"rule": "tri4"}),
    ("m_area_cubtri19", "area_field", '{ "rule" : "cubtri19" }', {"rule": "cubtri19"}),
    ("m_area_gm", "area_osc", '{ "rule" : "grundmann2d2" }', {"rule": "grundmann2d2"}),
    ("m_area_degree", "area_anchor", '{ "degree" : 4 }', {"degree": 4}),
    ("m_area_noadapt", "area_gradsq", '{ "rule" : "cubtri7", "adapt" : false }', {"rule": "cubtri7", "adapt": False}),
    ("m_vol_default", "vol_poly", "{}", {}),
    ("m_vol_osc", "vol_osc", "{}", {}),
    ("m_vol_keast4", "vol_field", '{ "rule" : "keast4" }', {"rule": "keast4"}),
    ("m_vol_tet5", "vol_osc", '{ "rule" : "tet5" }', {"rule": "tet5"}),
    ("m_vol_gm3", "vol_gradsq", '{ "rule" : "grundmann3d3" }', {"rule": "grundmann3d3"}),
    ("m_vol_degree", "vol_sel", '{ "degree" : 5 }', {"degree": 5}),
]
# gradients through the rule-table integrator: (method case, wrt)
METHOD_GRAD_CASES = [("m_line_default", -1), ("m_line_degree", 1), ("m_area_default", -1), ("m_area_cubtri19", 0), ("m_vol_keast4", -1),
                     ("m_vol_keast4", 1)]


def fields_for(vert):
    return _fields(vert)


# gradients: (case name, wrt) with wrt = -1 for positions or the index of the field; evaluated through the VM by
# LineIntegral/AreaIntegral(...).gradient(mesh) / .fieldgradient(field, mesh)
GRAD_CASES = [("line_poly", -1), ("line_osc", -1), ("line_field", -1), ("line_field", 0), ("line_field", 1), ("line_sel", -1),
              ("area_quad", -1), ("area_osc", -1), ("area_field", -1), ("area_field", 0), ("area_field", 1), ("area_sel", 0),
              ("line_tangent", -1), ("area_anchor", -1), ("area_anchor", 0), ("area_gradsq", -1), ("area_gradsq", 0),
              ("area_gradvec", 0),
              ("vol_poly", -1), ("vol_osc", -1), ("vol_field", -1), ("vol_field", 0), ("vol_field", 1), ("vol_sel", 0), ("vol_gradsq", -1),
              ("vol_gradsq", 0)]


# ScalarPotential: name, mesh, f, gradf or None, Morpho sources, selected vertex ids or None
POT_CASES = [
    ("pot_grad", "sphere", lambda x, y, z: x ** 2 + y * z - cos(2.5 * z), lambda x, y, z: (2 * x, z, y + 2.5 * sin(2.5 * z)),
     "fn (x,y,z) x^2+y*z-cos(2.5*z)", "fn (x,y,z) Matrix([2*x, z, y+2.5*sin(2.5*z)])", None),
    ("pot_fd", "sphere", lambda x, y, z: exp(-x * x) * y + sqrt(1.5 + z), None, "fn (x,y,z) exp(-x*x)*y+sqrt(1.5+z)", None, None),
    ("pot_sel", "loop", lambda x, y, z: sin(3 * x) * y - z, None, "fn (x,y,z) sin(3*x)*y-z", None, [0, 2, 3, 10, 23]),
    ("pot_2d", "disk2d", lambda x, y: x * y + exp(x), lambda x, y: (y + exp(x), x), "fn (x,y) x*y+exp(x)", "fn (x,y) Matrix([y+exp(x), x])", None),
]
