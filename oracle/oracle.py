"""ctypes front-ends for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* :class:`Oracle`    — oracle/libmorpho_oracle.so, our plain-C restatement
                       (oracle/morpho_oracle.c) of the reference algorithms.
* :class:`Reference` — oracle/_ref/libmorpho_refharness.so, our thin harness
                       around the UNMODIFIED reference library compiled from
                       /root/reference/src (oracle/refbuild/).

Both expose the same verbs (total / integrand / gradient / fieldgradient) on
plain numpy arrays so that tests can compare them with each other and with the
CUDA path on identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field as dc_field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FUNCTIONAL_IDS = {
    "Length": 1, "AreaEnclosed": 2, "Area": 3, "VolumeEnclosed": 4, "Volume": 5,
    "LineCurvatureSq": 6, "GradSq": 7, "NormSq": 8, "Nematic": 9, "NematicElectric": 10, "EquiElement": 11,
}

MO_ERRORS = {1: "args", 2: "FnctlELNtFnd", 3: "degenerate", 4: "VolEnclZero", 5: "unsupported"}


class OracleError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"oracle error {code} ({MO_ERRORS.get(code, '?')}) {msg}")
        self.code = code
        self.id = MO_ERRORS.get(code, "?")


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class _MoMesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("nv", C.c_int), ("vert", C.POINTER(C.c_double)),
                ("nel", C.c_int * 4), ("conn", C.POINTER(C.c_int) * 4)]


class _MoFunctional(C.Structure):
    _fields_ = [("functional", C.c_int), ("grade", C.c_int),
                ("ksplay", C.c_double), ("ktwist", C.c_double), ("kbend", C.c_double), ("pitch", C.c_double),
                ("haspitch", C.c_int), ("integrandonly", C.c_int),
                ("psize", C.c_int * 2), ("field", C.POINTER(C.c_double) * 2), ("wrt", C.c_int),
                ("weight", C.POINTER(C.c_double)), ("nweight", C.c_int)]


@dataclass
class FunctionalSpec:
    """Host-side description of a functional instance (mirrors the properties the
    reference reads in *_prepareref, functional.c:2940-2961, 3622-3642, 3757-3791)."""
    name: str
    grade: int = -1
    ksplay: float = 1.0
    ktwist: float = 1.0
    kbend: float = 1.0
    pitch: float | None = None
    integrandonly: bool = False
    fields: list = dc_field(default_factory=list)   # list of np.ndarray (nv, psize)
    weight: object = None                           # EquiElement: one weight per element of the grade compared, or None


class Oracle:
    """Plain-C restatement (oracle/morpho_oracle.c)."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libmorpho_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(path)
        L.mo_fdstepsize.restype = C.c_double
        L.mo_fdstepsize.argtypes = [C.c_double, C.c_int]
        for fn in ("mo_total", "mo_integrand", "mo_gradient", "mo_fieldgradient"):
            getattr(L, fn).restype = C.c_int
            getattr(L, fn).argtypes = [C.POINTER(_MoMesh), C.POINTER(_MoFunctional), C.POINTER(C.c_int), C.c_int,
                                       C.POINTER(C.c_double)]
        L.mo_count.argtypes = [C.POINTER(_MoMesh), C.POINTER(_MoFunctional), C.POINTER(C.c_int)]
        L.mo_transpose.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.mo_edges_from.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.mo_binbounds.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        for fn in ("mo_pairwise_coulomb_integrand", "mo_pairwise_coulomb_gradient"):
            getattr(L, fn).argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        for fn in ("mo_pairwise_integrand", "mo_pairwise_gradient"):
            getattr(L, fn).argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        L.mo_pairwise_total.restype = C.c_double
        L.mo_pairwise_total.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double]

    # -- helpers ---------------------------------------------------------------
    @staticmethod
    def _mesh(vert, grades):
        vert = np.ascontiguousarray(vert, dtype=np.float64)
        m = _MoMesh()
        m.dim, m.nv, m.vert = vert.shape[1], vert.shape[0], _dptr(vert)
        keep = [vert]
        for g in range(4):
            c = grades.get(g)
            if c is None or g == 0:
                m.nel[g] = 0
                continue
            c = np.ascontiguousarray(c, dtype=np.int32)
            keep.append(c)
            m.nel[g] = c.shape[0]
            m.conn[g] = _iptr(c)
        return m, keep

    @staticmethod
    def _func(spec: FunctionalSpec, wrt=0):
        f = _MoFunctional()
        f.functional = FUNCTIONAL_IDS[spec.name]
        f.grade = spec.grade
        f.ksplay, f.ktwist, f.kbend = spec.ksplay, spec.ktwist, spec.kbend
        f.pitch = 0.0 if spec.pitch is None else spec.pitch
        f.haspitch = 0 if spec.pitch is None else 1
        f.integrandonly = int(spec.integrandonly)
        keep = []
        for i, fl in enumerate(spec.fields[:2]):
            a = np.array(fl, dtype=np.float64, order="C", copy=True)     # FD perturbs in place
            if a.ndim == 1:
                a = a[:, None]
            keep.append(a)
            f.psize[i] = a.shape[1]
            f.field[i] = _dptr(a)
        f.wrt = wrt
        if spec.weight is not None:
            w = np.ascontiguousarray(spec.weight, dtype=np.float64).ravel()
            keep.append(w)
            f.weight, f.nweight = _dptr(w), w.size
        return f, keep

    def _run(self, fn, vert, grades, spec, sel, nout, wrt=0):
        m, k1 = self._mesh(vert, grades)
        f, k2 = self._func(spec, wrt)
        out = np.zeros(nout, dtype=np.float64)
        s = None if sel is None else np.ascontiguousarray(sel, dtype=np.int32)
        rc = fn(C.byref(m), C.byref(f), _iptr(s), 0 if s is None else len(s), _dptr(out))
        del k1, k2
        if rc:
            raise OracleError(rc)
        return out

    def count(self, vert, grades, spec):
        m, k1 = self._mesh(vert, grades)
        f, k2 = self._func(spec)
        n = C.c_int(0)
        rc = self.lib.mo_count(C.byref(m), C.byref(f), C.byref(n))
        if rc:
            raise OracleError(rc)
        return n.value

    # -- verbs -----------------------------------------------------------------
    def total(self, vert, grades, spec, sel=None):
        return float(self._run(self.lib.mo_total, vert, grades, spec, sel, 1)[0])

    def integrand(self, vert, grades, spec, sel=None):
        return self._run(self.lib.mo_integrand, vert, grades, spec, sel, self.count(vert, grades, spec))

    def gradient(self, vert, grades, spec, sel=None):
        vert = np.asarray(vert)
        return self._run(self.lib.mo_gradient, vert, grades, spec, sel, vert.size).reshape(vert.shape)

    def fieldgradient(self, vert, grades, spec, wrt=0, sel=None):
        fl = np.asarray(spec.fields[wrt])
        return self._run(self.lib.mo_fieldgradient, vert, grades, spec, sel, fl.size, wrt).reshape(fl.shape)

    def fdstepsize(self, x, order=1):
        return self.lib.mo_fdstepsize(float(x), int(order))

    def transpose(self, nv, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        tptr = np.zeros(nv + 1, dtype=np.int32)
        tidx = np.zeros(conn.size, dtype=np.int32)
        self.lib.mo_transpose(nv, conn.shape[0], conn.shape[1], _iptr(conn), _iptr(tptr), _iptr(tidx))
        return tptr, tidx

    def edges_from(self, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        nvper = conn.shape[1]
        out = np.zeros((conn.shape[0] * nvper * (nvper - 1) // 2, 2), dtype=np.int32)
        n = self.lib.mo_edges_from(conn.shape[0], nvper, _iptr(conn), _iptr(out))
        return out[:n].copy()

    def binbounds(self, nel, nbins):
        b = np.zeros(nbins + 1, dtype=np.int32)
        self.lib.mo_binbounds(nel, nbins, _iptr(b))
        return b

    def pairwise(self, x, mode, kind=0, params=(1.0,), cutoff=0.0):
        """PairwisePotential (modules/functionals.morpho:87-141) for the closure families of tests/golden/pairwise_cases.py."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, dim = x.shape
        p = np.ascontiguousarray(list(params) + [0.0] * 4, dtype=np.float64)
        if mode == "total":
            return float(self.lib.mo_pairwise_total(n, dim, _dptr(x), kind, _dptr(p), cutoff))
        if mode == "integrand":
            out = np.zeros(n)
            self.lib.mo_pairwise_integrand(n, dim, _dptr(x), kind, _dptr(p), cutoff, _dptr(out))
        else:
            out = np.zeros((n, dim))
            self.lib.mo_pairwise_gradient(n, dim, _dptr(x), kind, _dptr(p), cutoff, _dptr(out))
        return out

    def pairwise_coulomb(self, x, mode, cutoff=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, dim = x.shape
        if mode == "integrand":
            out = np.zeros(n)
            self.lib.mo_pairwise_coulomb_integrand(n, dim, _dptr(x), cutoff, _dptr(out))
        else:
            out = np.zeros((n, dim))
            self.lib.mo_pairwise_coulomb_gradient(n, dim, _dptr(x), cutoff, _dptr(out))
        return out


class Reference:
    """The unmodified reference library behind our C harness (oracle/_ref/)."""

    _inited = False

    def __init__(self, nthreads=0, path=None):
        d = os.path.join(HERE, "_ref")
        path = path or os.path.join(d, "libmorpho_refharness.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle/refbuild/build_ref.sh where /root/reference exists")
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        os.environ["HOME"] = os.path.join(d, "home")       # resources.c:246-273 reads $HOME/.morphopackages
        L = self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.refh_last_error.restype = C.c_char_p
        L.refh_mesh_new.restype = C.c_void_p
        L.refh_mesh_new.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.refh_mesh_set_vertices.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.refh_mesh_set_grade.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.refh_mesh_set_grade_dok.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.refh_mesh_addgrade.argtypes = [C.c_void_p, C.c_int]
        L.refh_mesh_grade_count.argtypes = [C.c_void_p, C.c_int]
        L.refh_mesh_get_conn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_long]
        L.refh_mesh_free.argtypes = [C.c_void_p]
        L.refh_field_new.restype = C.c_void_p
        L.refh_field_new.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.refh_field_size.restype = C.c_long
        L.refh_field_size.argtypes = [C.c_void_p]
        L.refh_field_free.argtypes = [C.c_void_p]
        L.refh_selection_new.restype = C.c_void_p
        L.refh_selection_new.argtypes = [C.c_void_p]
        L.refh_selection_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.refh_selection_ids.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.refh_selection_free.argtypes = [C.c_void_p]
        L.refh_functional_new.restype = C.c_void_p
        L.refh_functional_new.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                          C.POINTER(C.c_char_p), C.POINTER(C.c_double)]
        L.refh_functional_free.argtypes = [C.c_void_p]
        L.refh_functional_set_rowmatrix.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
        L.refh_eval.restype = C.c_long
        L.refh_eval.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_double), C.c_long, C.POINTER(C.c_double)]
        L.refh_time.restype = C.c_double
        L.refh_time.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.refh_integral.restype = C.c_long
        L.refh_integral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refh_fdstepsize.restype = C.c_double
        L.refh_fdstepsize.argtypes = [C.c_double, C.c_int]
        rc = L.refh_init(int(nthreads))
        if rc:
            raise RuntimeError(f"refh_init failed ({rc}): {L.refh_last_error().decode()}")

    def set_threads(self, n):
        self.lib.refh_set_threads(int(n))

    def err(self):
        return self.lib.refh_last_error().decode()

    # -- objects ---------------------------------------------------------------
    def mesh(self, vert, grades, via_dok=False):
        return RefMesh(self, vert, grades, via_dok)

    def fdstepsize(self, x, order=1):
        return self.lib.refh_fdstepsize(float(x), int(order))


class RefMesh:
    def __init__(self, ref: Reference, vert, grades, via_dok=False):
        self.ref, L = ref, ref.lib
        vert = np.ascontiguousarray(vert, dtype=np.float64)
        self.nv, self.dim = vert.shape
        self.h = L.refh_mesh_new(self.dim, self.nv, _dptr(vert))
        self._fields, self._sels, self._funcs = [], [], []
        for g in sorted(grades):
            if g == 0:
                continue
            c = np.ascontiguousarray(grades[g], dtype=np.int32)
            rc = (L.refh_mesh_set_grade_dok(self.h, g, c.shape[0], _iptr(c)) if via_dok
                  else L.refh_mesh_set_grade(self.h, g, c.shape[0], _iptr(c), 1))
            if rc:
                raise RuntimeError(f"set_grade failed {rc}")

    def set_vertices(self, vert):
        vert = np.ascontiguousarray(vert, dtype=np.float64)
        self.ref.lib.refh_mesh_set_vertices(self.h, _dptr(vert))

    def addgrade(self, g):
        return self.ref.lib.refh_mesh_addgrade(self.h, g)

    def integral(self, g, prog, fields=()):
        """Per-element integrals through the reference's OWN integrate_integrate (integrate.c:769-800) with the postfix
        program ``prog`` (list of (op, a, c), include/morpho_cuda.h) as the integrand; ``fields`` are (nv, psize) arrays
        whose components are the interpolated quantities, flattened in order."""
        ops = np.zeros(len(prog), dtype=np.dtype([("op", "<i4"), ("a", "<i4"), ("c", "<f8")]))
        for i, (op, a, c) in enumerate(prog):
            ops[i] = (op, a, c)
        comps, keep = [], []
        for f in fields:
            f = np.ascontiguousarray(np.asarray(f, dtype=np.float64).reshape(self.nv, -1))
            keep.append(f)
            for k in range(f.shape[1]):
                comps.append((f.ctypes.data + 8 * k, f.shape[1]))
        nq = len(comps)
        ptrs = (C.c_void_p * max(nq, 1))(*[p for p, _ in comps])
        strides = np.array([s for _, s in comps] or [1], dtype=np.int32)
        nel = self.count(g)
        out = np.zeros(max(nel, 1), dtype=np.float64)
        n = self.ref.lib.refh_integral(self.h, g, len(prog), ops.ctypes.data, nq, ptrs, strides.ctypes.data, out.ctypes.data)
        if n < 0:
            raise RuntimeError("refh_integral failed: " + self.ref.err())
        return out[:n]

    def count(self, g):
        return self.ref.lib.refh_mesh_grade_count(self.h, g)

    def conn(self, row, col, create=True):
        """(cptr, rix) of connectivity element (row, col) as the reference holds it."""
        L = self.ref.lib
        ncols = C.c_int(0)
        ne = L.refh_mesh_get_conn(self.h, row, col, int(create), C.byref(ncols), None, None, 0)
        if ne < 0:
            return None
        cptr = np.zeros(ncols.value + 1, dtype=np.int32)
        rix = np.zeros(max(ne, 1), dtype=np.int32)
        L.refh_mesh_get_conn(self.h, row, col, 0, C.byref(ncols), _iptr(cptr), _iptr(rix), ne)
        return cptr, rix[:ne]

    def field(self, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        psize = 1 if a.ndim == 1 else a.shape[1]
        h = self.ref.lib.refh_field_new(self.h, psize, _dptr(a))
        if not h:
            raise RuntimeError("field_new failed")
        self._fields.append(h)
        return h

    def selection(self, g, ids):
        L = self.ref.lib
        h = L.refh_selection_new(self.h)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        L.refh_selection_add(h, g, len(ids), _iptr(ids))
        self._sels.append(h)
        return h

    def selection_ids(self, sel, g):
        L = self.ref.lib
        n = L.refh_selection_ids(sel, g, None, 0)
        out = np.zeros(max(n, 1), dtype=np.int32)
        L.refh_selection_ids(sel, g, _iptr(out), n)
        return out[:n]

    def functional(self, spec: FunctionalSpec, field_handles=()):
        L = self.ref.lib
        names, vals = [], []
        if spec.name in ("Nematic",):
            names += ["ksplay", "ktwist", "kbend"]
            vals += [spec.ksplay, spec.ktwist, spec.kbend]
            if spec.pitch is not None:
                names.append("pitch"); vals.append(spec.pitch)
        if spec.name == "LineCurvatureSq" and spec.integrandonly:
            names.append("integrandonly"); vals.append(1.0)
        if spec.grade > 0 and spec.name in ("GradSq", "Nematic", "NematicElectric", "EquiElement"):
            names.append("grade"); vals.append(float(spec.grade))
        fa = (C.c_void_p * max(len(field_handles), 1))(*field_handles)
        na = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        va = (C.c_double * max(len(vals), 1))(*vals)
        h = L.refh_functional_new(spec.name.encode(), len(field_handles), fa, len(names), na, va)
        if not h:
            raise RuntimeError("functional_new failed: " + self.ref.err())
        if spec.weight is not None:           # EquiElement(weight=Matrix 1 x nel), functional.c:2895-2907
            w = np.ascontiguousarray(spec.weight, dtype=np.float64).ravel()
            if L.refh_functional_set_rowmatrix(h, b"weight", w.size, _dptr(w)):
                raise RuntimeError("functional_set_rowmatrix failed: " + self.ref.err())
        self._funcs.append(h)
        return h

    def eval(self, func, method, nout, sel=None, field=None, with_mesh=True, timing=False):
        out = np.zeros(max(nout, 1), dtype=np.float64)
        secs = C.c_double(0.0)
        n = self.ref.lib.refh_eval(func, method.encode(), self.h if with_mesh else None, sel, field,
                                   _dptr(out), out.size, C.byref(secs))
        if n < 0:
            raise RuntimeError(f"reference {method} failed ({n}): {self.ref.err()}")
        return (out[:n], secs.value) if timing else out[:n]

    def time(self, func, method, reps, sel=None, field=None):
        return self.ref.lib.refh_time(func, method.encode(), self.h, sel, field, int(reps))

    def free(self):
        L = self.ref.lib
        for h in self._funcs: L.refh_functional_free(h)
        for h in self._sels: L.refh_selection_free(h)
        for h in self._fields: L.refh_field_free(h)
        if self.h:
            L.refh_mesh_free(self.h)
        self.h = None
        self._funcs, self._sels, self._fields = [], [], []


def reference_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libmorpho_refharness.so"))
