"""Host-side mirrors of Morpho's Mesh / Field / Selection for the accelerated path.

These are thin owners of device mirrors kept by libmorpho_cuda.so; they follow the
reference objects' meaning (mesh.h:25-31, field.h:25-41, selection.h:22-31):

* ``Mesh.vertexmatrix()`` / ``setvertexmatrix(v)``  — mesh.c:1133-1160
* ``Mesh.addgrade(g)``                              — mesh.c:419-489 (edges from faces)
* ``Mesh.count(g)``, ``maxgrade()``                 — mesh.c:334-340
* ``Field(mesh, prototype)``                        — field.c:107-170 (vertex, P1)
* ``Selection(mesh, ...)``                          — selection.c (id sets per grade)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

__all__ = ["Mesh", "Field", "Selection"]


def _as_conn(a, g):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if a.ndim != 2 or a.shape[1] != g + 1:
        raise ValueError(f"grade {g} connectivity must have shape (nel, {g + 1})")
    # the reference keeps ids ascending inside an element (DOK→CCS conversion sorts, sparse.c:505-510)
    return np.ascontiguousarray(np.sort(a, axis=1))


class Mesh:
    """Device-resident mirror of a simplicial mesh.

    ``vert`` is ``(nv, dim)`` float64 (== Morpho's column-major ``dim x nv`` Matrix);
    ``edges`` / ``faces`` / ``volumes`` are ``(nel, g+1)`` int arrays of 0-based ids.
    """

    def __init__(self, vert, edges=None, faces=None, volumes=None):
        _lib.init()
        vert = np.ascontiguousarray(vert, dtype=np.float64)
        if vert.ndim != 2 or vert.shape[1] not in (2, 3):
            raise ValueError("vertex array must be (nv, 2) or (nv, 3)")
        self.nv, self.dim = vert.shape
        self._vert = vert.copy()
        self._conn = {}
        self._h = lib().mcu_mesh_create(self.dim, self.nv)
        if not self._h:
            check(_lib.MCU_ERR_CUDA if "CUDA" in lib().mcu_last_error().decode() else _lib.MCU_ERR_ARGS)
        check(lib().mcu_mesh_set_vertices(self._h, self._vert.ctypes.data))
        for g, c in ((1, edges), (2, faces), (3, volumes)):
            if c is not None:
                self._set_grade(g, c)

    # -- construction ----------------------------------------------------------
    @classmethod
    def load(cls, path):
        """Read Morpho's ``.mesh`` text format (mesh.c:910-1080): sections vertices/edges/faces/volumes, 1-based ids."""
        sections = {"vertices": [], "edges": [], "faces": [], "volumes": []}
        cur = None
        with open(path) as fh:
            for line in fh:
                t = line.split()
                if not t:
                    continue
                if t[0] in sections:
                    cur = t[0]
                    continue
                if cur is None:
                    continue
                sections[cur].append(t[1:])
        vert = np.array(sections["vertices"], dtype=np.float64)
        kw = {}
        for name, key in (("edges", "edges"), ("faces", "faces"), ("volumes", "volumes")):
            if sections[name]:
                kw[key] = np.array(sections[name], dtype=np.int64) - 1
        return cls(vert, **kw)

    def _set_grade(self, g, conn):
        conn = _as_conn(conn, g)
        if g > self.dim:
            raise ValueError("grade exceeds mesh dimension")
        self._conn[g] = conn
        check(lib().mcu_mesh_set_connectivity(self._h, g, conn.shape[0], conn.ctypes.data))

    # -- reference-style accessors ---------------------------------------------
    def vertexmatrix(self):
        """Copy of the vertex matrix, shape (nv, dim)."""
        return self._vert.copy()

    def setvertexmatrix(self, vert):
        vert = np.ascontiguousarray(vert, dtype=np.float64)
        if vert.shape != (self.nv, self.dim):
            raise _lib.MorphoError(_lib.MCU_ERR_ARGS, "MshVrtMtrxDim", "Vertex matrix dimensions inconsistent with mesh.")
        self._vert = vert.copy()
        check(lib().mcu_mesh_set_vertices(self._h, self._vert.ctypes.data))

    def maxgrade(self):
        for g in range(self.dim, 0, -1):
            if g in self._conn:
                return g
        return 0

    def count(self, g=0):
        if g == 0:
            return self.nv
        return self._conn[g].shape[0] if g in self._conn else 0

    def connectivity(self, g):
        return self._conn.get(g)

    def lower(self, row, col):
        """``(cptr, rix)`` of the grade-lowering connectivity matrix conn(row, col), 1 <= row < col
        (``Mesh.connectivitymatrix(row, col)``; mesh_addlowermatrix, mesh.c:621-660): per element of grade ``col`` the ids,
        ascending, of the elements of grade ``row`` it contains.  Built on the device (mcu_connect.cu)."""
        ncol = self.count(col)
        from math import comb
        cptr = np.zeros(ncol + 1, dtype=np.int32)
        rix = np.zeros(max(ncol * comb(col + 1, row + 1), 1), dtype=np.int32)
        nnz, nrows, ncols = C.c_int(0), C.c_int(0), C.c_int(0)
        check(lib().mcu_mesh_lower(self._h, row, col, cptr.ctypes.data, rix.ctypes.data, C.byref(nnz), C.byref(nrows), C.byref(ncols)))
        return cptr[: ncols.value + 1].copy(), rix[: nnz.value].copy(), nrows.value

    def addgrade(self, g):
        """Create grade ``g`` elements from the next higher grade present (mesh_addgrade, mesh.c:419-489)."""
        if g in self._conn:
            return self._conn[g]
        higher = next((h for h in range(g + 1, self.dim + 1) if h in self._conn), None)
        if higher is None:
            return None
        nel = C.c_int(0)
        check(lib().mcu_mesh_addgrade(self._h, g, C.byref(nel)))     # sort + scan on the device (mcu_connect.cu)
        out = np.zeros((nel.value, g + 1), dtype=np.int32)
        if nel.value:
            check(lib().mcu_mesh_get_connectivity(self._h, g, out.ctypes.data))
        self._conn[g] = out
        return self._conn[g]

    def transpose(self, g):
        """(tptr, tidx) of the vertex→element incidence held on the device (for parity checks)."""
        tptr = np.zeros(self.nv + 1, dtype=np.int32)
        tidx = np.zeros(self._conn[g].size, dtype=np.int32)
        check(lib().mcu_mesh_get_transpose(self._h, g, tptr.ctypes.data, tidx.ctypes.data))
        return tptr, tidx

    def close(self):
        if getattr(self, "_h", None):
            lib().mcu_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        return f"<Mesh: {self.nv} vertices>"


class Field:
    """Vertex-based (P1) field with ``psize`` doubles per vertex (field.c:107-170).

    ``Field(mesh, 0.0)`` is a scalar field; ``Field(mesh, np.array([1,0,0]))`` copies the
    prototype to every vertex; ``Field(mesh, data)`` with ``data.shape == (nv, psize)`` sets it.
    """

    def __init__(self, mesh: Mesh, prototype=0.0):
        self.mesh = mesh
        p = np.asarray(prototype, dtype=np.float64)
        if p.ndim == 2 and p.shape[0] == mesh.nv:
            data = p
        elif p.ndim == 1 and p.shape[0] == mesh.nv and mesh.nv > 9:
            data = p[:, None]
        else:
            data = np.tile(p.reshape(1, -1), (mesh.nv, 1))
        self.psize = data.shape[1]
        self._data = np.ascontiguousarray(data, dtype=np.float64).copy()
        self._h = lib().mcu_field_create(mesh._h, self.psize)
        if not self._h:
            check(_lib.MCU_ERR_ARGS)
        check(lib().mcu_field_set(self._h, self._data.ctypes.data))

    def data(self):
        return self._data.copy()

    def assign(self, data):
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(self.mesh.nv, self.psize)
        self._data = data.copy()
        check(lib().mcu_field_set(self._h, self._data.ctypes.data))

    def close(self):
        if getattr(self, "_h", None):
            lib().mcu_field_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Selection:
    """A set of element ids per grade (selection.h:22-31).

    ``Selection(mesh, ids={2: [...]})``; ``Selection(mesh, fn)`` selects vertices where
    ``fn(x, y, z)`` is true (selection_selectwithfunction, selection.c:160-175);
    ``Selection(mesh, boundary=True)`` selects boundary facets and their vertices
    (selection_selectboundary, selection.c:211-232).
    """

    def __init__(self, mesh: Mesh, fn=None, ids=None, boundary=False):
        self.mesh = mesh
        self._ids = {}
        self._h = {}
        if fn is not None:
            v = mesh._vert
            mask = np.array([bool(fn(*x)) for x in v])
            self._ids[0] = np.nonzero(mask)[0].astype(np.int32)
        if ids:
            for g, a in ids.items():
                self._ids[int(g)] = np.unique(np.asarray(a, dtype=np.int32))
        if boundary:
            self._select_boundary()

    def _select_boundary(self):
        """selection_selectboundary (selection.c:211-232) on the device (mcu_mesh_boundary): elements of grade
        maxgrade-1 contained in exactly one top element, plus their vertices (selection_addgradelower)."""
        m = self.mesh
        mg = m.maxgrade()
        if mg < 1:
            return
        bnd = mg - 1
        sub = m.addgrade(bnd) if bnd > 0 else None
        cap = m.nv if bnd == 0 else sub.shape[0]
        ids = np.zeros(max(cap, 1), dtype=np.int32)
        g, n = C.c_int(), C.c_int()
        check(lib().mcu_mesh_boundary(m._h, C.byref(g), C.byref(n), ids.ctypes.data, cap))
        sel = ids[: n.value].copy()
        self._ids[bnd] = sel
        if bnd > 0:
            self._ids[0] = np.unique(sub[sel].ravel()).astype(np.int32)   # selection_addgradelower(sel, 0)

    def idlistforgrade(self, g):
        return self._ids.get(g, np.zeros(0, dtype=np.int32)).copy()

    def isselected(self, g, i):
        return bool(np.isin(i, self._ids.get(g, [])))

    def _handle(self, g):
        if g not in self._h:
            ids = np.ascontiguousarray(self._ids.get(g, np.zeros(0, dtype=np.int32)), dtype=np.int32)
            h = lib().mcu_selection_create(self.mesh._h, g, len(ids), ids.ctypes.data)
            if not h:
                check(_lib.MCU_ERR_ARGS)
            self._h[g] = h
        return self._h[g]

    def close(self):
        for h in self._h.values():
            lib().mcu_selection_destroy(h)
        self._h = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
