"""The functional classes of the accelerated path, with the reference's method names.

Each class mirrors a builtin class registered in functional.c:6260-6280 and offers
``total`` / ``integrand`` / ``gradient`` (/ ``fieldgradient``) with the reference's argument
convention: a Mesh, an optional Selection and an optional Field, in any order
(functional_validateargs, functional.c:82-102).  Results are numpy arrays in the reference
layouts.  A degenerate element makes ``gradient`` return ``None``
— the reference returns nil there (functional.c:392,412-415) — and named errors are raised
as :class:`MorphoError` with the reference's error id.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MCU_FIELDGRADIENT, MCU_GRADIENT, MCU_INTEGRAND, MCU_TOTAL, FUNCTIONAL_IDS, McuFunctional,
                   MorphoError, check, lib)
from .geometry import Field, Mesh, Selection

__all__ = ["Functional", "Length", "AreaEnclosed", "Area", "VolumeEnclosed", "Volume", "LineCurvatureSq", "EquiElement",
           "GradSq", "NormSq", "Nematic", "NematicElectric", "PairwisePotential", "LineIntegral", "AreaIntegral", "VolumeIntegral", "ScalarPotential"]


class Functional:
    name = None
    grade = -1

    def __init__(self):
        self.fields = []

    # -- descriptor ------------------------------------------------------------
    def _desc(self, wrt=0):
        d = McuFunctional()
        d.functional = FUNCTIONAL_IDS[self.name]
        d.grade = int(getattr(self, "grade", -1) or -1)
        d.ksplay = float(getattr(self, "ksplay", 1.0))
        d.ktwist = float(getattr(self, "ktwist", 1.0))
        d.kbend = float(getattr(self, "kbend", 1.0))
        pitch = getattr(self, "pitch", None)
        d.pitch = 0.0 if pitch is None else float(pitch)
        d.haspitch = 0 if pitch is None else 1
        d.integrandonly = int(bool(getattr(self, "integrandonly", False)))
        for i, f in enumerate(self.fields[:2]):
            d.field[i] = f._h
        d.wrt = wrt
        w = getattr(self, "weight", None)
        if w is not None:
            self._w = np.ascontiguousarray(w, dtype=np.float64).ravel()      # kept alive for the call
            d.host_weight, d.nweight = self._w.ctypes.data, self._w.size
        return d

    @staticmethod
    def _args(args):
        """functional_validateargs (functional.c:82-102)."""
        mesh = sel = fld = None
        for a in args:
            if isinstance(a, Mesh):
                mesh = a
            elif isinstance(a, Selection):
                sel = a
            elif isinstance(a, Field):
                fld = a
                mesh = a.mesh
        if mesh is None:
            raise MorphoError(_lib.MCU_ERR_ARGS, "FnctlIntMsh", "Method 'integrand' requires a mesh as the argument.")
        return mesh, sel, fld

    def _loop_grade(self, mesh):
        if self.name in ("GradSq", "Nematic", "NematicElectric"):
            g = int(getattr(self, "grade", -1) or -1)
            return g if g > 0 else mesh.maxgrade()
        return {"Length": 1, "AreaEnclosed": 1, "Area": 2, "VolumeEnclosed": 2, "Volume": 3,
                "LineCurvatureSq": 0, "NormSq": 0, "EquiElement": 0}[self.name]

    def _eval(self, mode, args, wrt=0):
        mesh, sel, fld = self._args(args)
        d = self._desc(wrt)
        L = lib()
        n = L.mcu_result_size(mesh._h, C.byref(d), mode)
        g = self._loop_grade(mesh)
        if n < 0:
            if g > 0 and mesh.count(g) == 0 and g not in mesh._conn:
                raise MorphoError(_lib.MCU_ERR_ELNOTFOUND, "FnctlELNtFnd", f"Mesh does not provide elements of grade {g}.")
            check(_lib.MCU_ERR_ARGS)
        out = np.zeros(max(n, 1), dtype=np.float64)
        hsel = sel._handle(g) if sel is not None else None
        rc = L.mcu_eval(mesh._h, C.byref(d), hsel, mode, out.ctypes.data)
        if rc == _lib.MCU_ERR_DEGENERATE:
            return None            # nil, as the reference (functional.c:412-415)
        check(rc)
        return out[:n]

    # -- the builtin methods (functional.h:55-60) --------------------------------
    def total(self, *args):
        r = self._eval(MCU_TOTAL, args)
        return None if r is None else float(r[0])

    def integrand(self, *args):
        return self._eval(MCU_INTEGRAND, args)

    def gradient(self, *args):
        mesh, _, _ = self._args(args)
        r = self._eval(MCU_GRADIENT, args)
        return None if r is None else r.reshape(mesh.nv, mesh.dim)


class Length(Functional):
    name = "Length"


class AreaEnclosed(Functional):
    name = "AreaEnclosed"


class Area(Functional):
    name = "Area"


class VolumeEnclosed(Functional):
    name = "VolumeEnclosed"


class Volume(Functional):
    name = "Volume"


class EquiElement(Functional):
    """EquiElement(grade=-1, weight=None) (functional.c:2775-2922): per vertex, the variance of the sizes of the incident
    elements of ``grade`` (default: the highest grade of the mesh) relative to their mean; ``weight``: one number per
    element (the reference's 1 x nel Matrix).  Gradient by the reference's central differences."""
    name = "EquiElement"

    def __init__(self, grade=-1, weight=None):
        super().__init__()
        self.grade, self.weight = grade, weight

    def _desc(self, wrt=0):
        d = super()._desc(wrt)
        d.grade = int(self.grade)          # 0 is a legal value here (it means "the highest grade", like -1)
        return d


class LineCurvatureSq(Functional):
    name = "LineCurvatureSq"

    def __init__(self, integrandonly=False):
        super().__init__()
        self.integrandonly = integrandonly


class _FieldFunctional(Functional):
    def fieldgradient(self, *args):
        mesh, sel, fld = self._args(args)
        wrt = 0
        if fld is not None and len(self.fields) > 1 and fld is self.fields[1]:
            wrt = 1
        r = self._eval(MCU_FIELDGRADIENT, args, wrt)
        if r is None:
            return None
        return r.reshape(mesh.nv, self.fields[wrt].psize)


class GradSq(_FieldFunctional):
    name = "GradSq"

    def __init__(self, field, grade=None):
        super().__init__()
        if not isinstance(field, Field):
            raise MorphoError(_lib.MCU_ERR_ARGS, "InvldArgs", "GradSq requires a field")
        self.fields = [field]
        self.grade = grade if grade else -1


class NormSq(_FieldFunctional):
    name = "NormSq"

    def __init__(self, field):
        super().__init__()
        if not isinstance(field, Field):
            raise MorphoError(_lib.MCU_ERR_ARGS, "InvldArgs", "NormSq requires a field")
        self.fields = [field]


class Nematic(_FieldFunctional):
    name = "Nematic"

    def __init__(self, field, ksplay=1.0, ktwist=1.0, kbend=1.0, pitch=None, grade=None):
        super().__init__()
        if not isinstance(field, Field):
            raise MorphoError(_lib.MCU_ERR_ARGS, "NmtcArgs", "Nematic requires a field as its argument.")
        self.fields = [field]
        self.ksplay, self.ktwist, self.kbend, self.pitch = ksplay, ktwist, kbend, pitch
        self.grade = grade if grade else -1


class NematicElectric(_FieldFunctional):
    name = "NematicElectric"

    def __init__(self, director, potential, grade=None):
        super().__init__()
        if not (isinstance(director, Field) and isinstance(potential, Field)):
            raise MorphoError(_lib.MCU_ERR_ARGS, "FnctlArgs", "NematicElectric requires two fields")
        self.fields = [director, potential]
        self.grade = grade if grade else -1


class PairwisePotential:
    """All-pairs potential on the mesh vertices (modules/functionals.morpho:87-141).

    The reference takes two closures ``func(r)``, ``grad(r)``.  Here the potential is named (the extension derives the
    same description from the closures' bytecode):

    * ``kind="coulomb", a=1``           f = a/r  (examples/thomson/thomson.morpho:26)
    * ``kind="power", a, p``            f = a r^-p  (integer p)
    * ``kind="laurent", terms=[(a,k)]`` f = sum a r^k  (integer k)
    * ``kind="lj", eps, sigma``         f = 4 eps ((sigma/r)^12 - (sigma/r)^6)
    * ``kind="yukawa", a, l``           f = a exp(-r/l)/r
    * ``kind="program", func, grad``    callables of r written with :mod:`morpho_b200.integrand` operators
    """

    def __init__(self, kind="coulomb", cutoff=None, **kw):
        self.cutoff = 0.0 if cutoff is None else float(cutoff)
        self._progs = None
        if kind == "coulomb":
            self.kind, self.params = _lib.MCU_PW_COULOMB, [float(kw.get("a", 1.0))]
        elif kind == "power":
            self.kind, self.params = _lib.MCU_PW_LAURENT, [1.0, float(kw["a"]), -float(int(kw["p"]))]
        elif kind == "laurent":
            t = list(kw["terms"])
            self.kind, self.params = _lib.MCU_PW_LAURENT, [float(len(t))] + [float(v) for ak in t for v in ak]
        elif kind == "lj":
            e, sg = float(kw["eps"]), float(kw["sigma"])
            self.kind, self.params = _lib.MCU_PW_LAURENT, [2.0, 4 * e * sg ** 12, -12.0, -4 * e * sg ** 6, -6.0]
        elif kind == "yukawa":
            self.kind, self.params = _lib.MCU_PW_YUKAWA, [float(kw["a"]), float(kw["l"])]
        elif kind == "program":
            from .integrand import trace
            self.kind, self.params = _lib.MCU_PW_PROGRAM, [0.0]
            self._progs = (trace(lambda x: kw["func"](x[0]), 1, []), trace(lambda x: kw["grad"](x[0]), 1, []))
        else:
            raise MorphoError(_lib.MCU_ERR_UNSUPPORTED, "", f"pairwise kind '{kind}' is not accelerated")

    def _run(self, mesh, mode, n):
        out = np.zeros(max(n, 1))
        x = mesh._vert
        if self._progs is not None:
            pf, pg = self._progs
            check(lib().mcu_pairwise_program(len(pf), pf.ctypes.data, len(pg), pg.ctypes.data))
        p = np.ascontiguousarray(self.params, dtype=np.float64)
        check(lib().mcu_pairwise(mesh.nv, mesh.dim, x.ctypes.data, self.kind, p.ctypes.data, self.cutoff, mode, out.ctypes.data))
        return out[:n]

    def integrand(self, mesh):
        return self._run(mesh, MCU_INTEGRAND, mesh.nv)

    def total(self, mesh):
        return float(self._run(mesh, MCU_TOTAL, 1)[0])

    def gradient(self, mesh):
        return self._run(mesh, MCU_GRADIENT, mesh.nv * mesh.dim).reshape(mesh.nv, mesh.dim)


class _QuadMethod(C.Structure):
    """mcu_quad_method (include/morpho_cuda.h)"""
    _fields_ = [("adapt", C.c_int), ("degree", C.c_int), ("rule", C.c_char * 32)]


class _Integral:
    """LineIntegral / AreaIntegral (functional.c:5186-5243, 5371-5426): the integral over every element of a callable
    of the position and of the interpolated field values, by the reference's default adaptive quadrature
    (integrate_integrate, integrate.c:769-800; mcu_integral.cu on the device).  ``fn(x, *q)`` is written with the
    operators and functions of :mod:`morpho_b200.integrand`; it is recorded once into a postfix program.
    ``gradient`` / ``fieldgradient`` are central differences over the quadrature like the reference's
    (functional_mapnumericalgradient / functional_mapnumericalfieldgradient, functional.c:1142-1247, 1305-1503)."""
    grade = 0

    def __init__(self, fn, *fields, method=None):
        self.fn = fn
        self.fields = list(fields)
        self._prog = None
        # method = None: the default integrator; a dict ({} included) selects the rule-table integrator like the
        # reference's `method` option (integrator_configurewithdictionary, integrate.c:2706-2751): "rule", "degree", "adapt"
        self.method = None
        if method is not None:
            unknown = set(method) - {"rule", "degree", "adapt"}
            if unknown:
                raise MorphoError(_lib.MCU_ERR_ARGS, "IntgrtrMthdTyp", f"unknown method entries {sorted(unknown)}")
            self.method = _QuadMethod(1 if method.get("adapt", True) else 0, int(method.get("degree", -1)),
                                      str(method.get("rule", "")).encode())

    def program(self, dim):
        if self._prog is None or self._prog[0] != dim:
            from .integrand import trace
            self._prog = (dim, trace(self.fn, dim, [f.psize for f in self.fields]))
        return self._prog[1]

    def _eval(self, mode, args):
        mesh, sel, _ = Functional._args(args)
        prog = self.program(mesh.dim)
        nel = mesh.count(self.grade)
        if nel == 0 and self.grade not in mesh._conn:
            raise MorphoError(_lib.MCU_ERR_ELNOTFOUND, "FnctlELNtFnd", f"Mesh does not provide elements of grade {self.grade}.")
        out = np.zeros(max(nel, 1) if mode == MCU_INTEGRAND else 1, dtype=np.float64)
        fh = (C.c_void_p * max(len(self.fields), 1))(*[f._h for f in self.fields])
        hsel = sel._handle(self.grade) if sel is not None else None
        mref = C.byref(self.method) if self.method is not None else None
        check(lib().mcu_integral_method(mesh._h, self.grade, len(prog), prog.ctypes.data, len(self.fields), fh, hsel, mode, mref, out.ctypes.data))
        return out[:nel] if mode == MCU_INTEGRAND else out

    def total(self, *args):
        return float(self._eval(MCU_TOTAL, args)[0])

    def integrand(self, *args):
        return self._eval(MCU_INTEGRAND, args)

    def _grad(self, args, wrt):
        mesh, sel, _ = Functional._args(args)
        prog = self.program(mesh.dim)
        if mesh.count(self.grade) == 0 and self.grade not in mesh._conn:
            raise MorphoError(_lib.MCU_ERR_ELNOTFOUND, "FnctlELNtFnd", f"Mesh does not provide elements of grade {self.grade}.")
        ncomp = mesh.dim if wrt < 0 else self.fields[wrt].psize
        out = np.zeros((mesh.nv, ncomp), dtype=np.float64)
        fh = (C.c_void_p * max(len(self.fields), 1))(*[f._h for f in self.fields])
        hsel = sel._handle(self.grade) if sel is not None else None
        mref = C.byref(self.method) if self.method is not None else None
        check(lib().mcu_integral_gradient_method(mesh._h, self.grade, len(prog), prog.ctypes.data, len(self.fields), fh, hsel, wrt, mref, out.ctypes.data))
        return out

    def gradient(self, *args):
        return self._grad(args, -1)

    def fieldgradient(self, field, *args):
        """Gradient with respect to ``field``, which must be one of the Fields the integral was built with."""
        for k, f in enumerate(self.fields):
            if f is field:
                return self._grad((field.mesh,) + tuple(args), k)
        raise MorphoError(_lib.MCU_ERR_ARGS, "FnctlFldGrdArgs", "fieldgradient needs one of the functional's own fields")


class LineIntegral(_Integral):
    grade = 1


class AreaIntegral(_Integral):
    grade = 2


class VolumeIntegral(_Integral):
    """functional.c:5471-5576; integrate_volint (integrate.c:677-752) on the device."""
    grade = 3


class ScalarPotential:
    """ScalarPotential(f[, gradf]) (functional.c:2255-2372): ``f(x, y, z)`` summed over the (selected) vertices; the
    gradient comes from ``gradf(x, y, z)`` (a sequence of ``dim`` expressions) if given, else from central differences
    of ``f`` exactly like the reference.  Both callables use the operators of :mod:`morpho_b200.integrand`."""

    def __init__(self, fn, gradfn=None):
        self.fn, self.gradfn = fn, gradfn

    def _eval(self, mode, args):
        from .integrand import trace_point
        mesh, sel, _ = Functional._args(args)
        use_grad = mode == MCU_GRADIENT and self.gradfn is not None
        progs = trace_point(self.gradfn if use_grad else self.fn, mesh.dim)
        if len(progs) != (mesh.dim if use_grad else 1):
            raise MorphoError(_lib.MCU_ERR_ARGS, "SclrPtFnCllbl", "ScalarPotential function returns the wrong number of values")
        allops = np.concatenate(progs)
        nops = np.array([len(p) for p in progs], dtype=np.int32)
        n = {MCU_TOTAL: 1, MCU_INTEGRAND: mesh.nv, MCU_GRADIENT: mesh.nv * mesh.dim}[mode]
        out = np.zeros(max(n, 1), dtype=np.float64)
        hsel = sel._handle(0) if sel is not None else None
        check(lib().mcu_pointwise(mesh._h, len(progs), nops.ctypes.data, allops.ctypes.data, hsel, mode, out.ctypes.data))
        return out[:n]

    def total(self, *args):
        return float(self._eval(MCU_TOTAL, args)[0])

    def integrand(self, *args):
        return self._eval(MCU_INTEGRAND, args)

    def gradient(self, *args):
        mesh, _, _ = Functional._args(args)
        return self._eval(MCU_GRADIENT, args).reshape(mesh.nv, mesh.dim)
