"""PairwisePotential cases shared by the golden generator (the INTERPRETED reference, modules/functionals.morpho:87-141,
run through oracle/_ref/morpho6), the oracle tests and the GPU parity tests.

`func` / `grad` are the Morpho closures a script hands to PairwisePotential(func, grad, cutoff); kind / params select
the oracle's C restatement of those closures (oracle/morpho_oracle.c pw_func / pw_grad: 0 Coulomb, 1 power, 2 Yukawa,
3 Lennard-Jones); LIB maps a case to the keyword arguments of morpho_b200.PairwisePotential.
"""
import numpy as np

# name, n points, kind, params, cutoff (0 = none), func closure, grad closure
CASES = [
    ("coulomb_500", 500, 0, [1.0], 0.0, "fn (r) 1/r", "fn (r) -1/r^2"),
    ("coulomb_1000", 1000, 0, [1.0], 0.0, "fn (r) 1/r", "fn (r) -1/r^2"),
    ("coulomb_2000", 2000, 0, [1.0], 0.0, "fn (r) 1/r", "fn (r) -1/r^2"),
    ("coulomb_cut_500", 500, 0, [1.0], 0.5, "fn (r) 1/r", "fn (r) -1/r^2"),
    ("power3_500", 500, 1, [2.0, 3.0], 0.0, "fn (r) 2/r^3", "fn (r) -6/r^4"),
    ("yukawa_500", 500, 2, [1.0, 0.4], 0.0, "fn (r) exp(-r/0.4)/r", "fn (r) -exp(-r/0.4)*(1/r^2+1/(0.4*r))"),
    ("lj_cut_500", 500, 3, [0.1, 0.2], 0.6, "fn (r) 4*0.1*((0.2/r)^12-(0.2/r)^6)", "fn (r) 4*0.1*(-12*(0.2/r)^12+6*(0.2/r)^6)/r"),
    ("coulomb_2d_300", 300, 0, [1.0], 0.0, "fn (r) 1/r", "fn (r) -1/r^2"),
]


def points(name, n):
    """Config 3's generator: normalised Gaussian points on S^2, seed 12345; the 2-D case a unit circle."""
    from morpho_b200 import meshgen as mg
    x = mg.sphere_points(n)
    if "_2d_" in name:
        x = np.ascontiguousarray(x[:, :2] / np.linalg.norm(x[:, :2], axis=1, keepdims=True))
    return x

from morpho_b200 import integrand as _ig  # noqa: E402  (symbolic operators for the "program" route)

LIB = {
    "coulomb_500": [dict(kind="coulomb"), dict(kind="program", func=lambda r: 1 / r, grad=lambda r: -1 / r ** 2)],
    "coulomb_1000": [dict(kind="coulomb")],
    "coulomb_2000": [dict(kind="coulomb")],
    "coulomb_cut_500": [dict(kind="coulomb", cutoff=0.5)],
    "power3_500": [dict(kind="power", a=2.0, p=3), dict(kind="laurent", terms=[(2.0, -3)])],
    "yukawa_500": [dict(kind="yukawa", a=1.0, l=0.4),
                   dict(kind="program", func=lambda r: _ig.exp(-r / 0.4) / r,
                        grad=lambda r: -_ig.exp(-r / 0.4) * (1 / r ** 2 + 1 / (0.4 * r)))],
    "lj_cut_500": [dict(kind="lj", eps=0.1, sigma=0.2, cutoff=0.6),
                   dict(kind="program", cutoff=0.6, func=lambda r: 4 * 0.1 * ((0.2 / r) ** 12 - (0.2 / r) ** 6),
                        grad=lambda r: 4 * 0.1 * (-12 * (0.2 / r) ** 12 + 6 * (0.2 / r) ** 6) / r)],
    "coulomb_2d_300": [dict(kind="coulomb")],
}
