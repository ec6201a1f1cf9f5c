"""Synthetic mesh generators for parity tests and benchmarks.

All generators are deterministic and return plain numpy arrays in the layout the
reference stores:

* ``vert``  float64, shape ``(nv, dim)`` C-contiguous  ==  Morpho's column-major
  ``dim x nv`` vertex Matrix (``mesh->vert->elements[i*dim+k]``, matrix.h:59-73).
* ``conn``  int32, shape ``(nel, g+1)``, vertex ids **sorted ascending inside each
  element**, which is what the reference's DOK->CCS conversion produces
  (sparse.c:505-510).

Nothing here is on the measured path; meshes are inputs.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "IcoTopology", "icosphere", "perturb_radially", "radial_factors", "disk", "tet_cube", "polyline_loop",
    "sphere_points", "edges_from_faces", "write_mesh_file",
]

_PHI = (1.0 + 5.0 ** 0.5) / 2.0

_ICO_VERTS = np.array([
    [-1, _PHI, 0], [1, _PHI, 0], [-1, -_PHI, 0], [1, -_PHI, 0],
    [0, -1, _PHI], [0, 1, _PHI], [0, -1, -_PHI], [0, 1, -_PHI],
    [_PHI, 0, -1], [_PHI, 0, 1], [-_PHI, 0, -1], [-_PHI, 0, 1],
], dtype=np.float64)

_ICO_FACES = np.array([
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
], dtype=np.int64)


class IcoTopology:
    """Index arithmetic of the frequency-``f`` geodesic icosphere, usable face by face so that a rank
    can build only its own element range of a very large sphere (morpho_b200/partition.py).

    Vertex order (the documented "generator order"): the 12 icosahedron corners, then the interior
    points of the 30 edges (edge list sorted by (lo, hi) corner id, points ordered from the lower
    corner), then for each of the 20 faces its interior grid points row by row.  Element order: face by
    face, row by row, alternating up/down triangles along a row.
    """

    def __init__(self, f: int):
        if f < 1:
            raise ValueError("frequency must be >= 1")
        self.f = f = int(f)
        faces = _ICO_FACES
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
        self.edges = np.unique(np.sort(e, axis=1), axis=0)          # 30 x 2, sorted by (lo, hi)
        assert self.edges.shape[0] == 30
        self.edge_id = {(int(a), int(b)): k for k, (a, b) in enumerate(self.edges)}
        self.n_edge_int = f - 1
        self.n_face_int = (f - 1) * (f - 2) // 2
        self.base_edge = 12
        self.base_face = 12 + 30 * self.n_edge_int
        self.nv = self.base_face + 20 * self.n_face_int
        self.nel = 20 * f * f
        self.tri_per_face = f * f
        self.corners = _ICO_VERTS / np.linalg.norm(_ICO_VERTS, axis=1, keepdims=True)
        rows = np.arange(f + 1)
        row_len = np.clip(f - rows - 1, 0, None)
        row_len[0] = 0
        self.row_off = np.concatenate([[0], np.cumsum(row_len)])    # row_off[i] = interior points before row i

    def _edge_point(self, p, q, t):
        """global id of the point t/f of the way from corner p to corner q (0 < t < f)."""
        lo, hi = (p, q) if p < q else (q, p)
        tt = t if p < q else self.f - t
        return self.base_edge + self.edge_id[(lo, hi)] * self.n_edge_int + (tt - 1)

    def face_grid(self, fi: int) -> np.ndarray:
        """(f+1) x (f+1) array G[i, j] of global vertex ids of face ``fi`` (-1 outside i + j <= f);
        the grid point is (k A + i B + j C) / f with k = f - i - j."""
        f = self.f
        A, B, C = (int(x) for x in _ICO_FACES[fi])
        ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
        kk = f - ii - jj
        G = np.full((f + 1, f + 1), -1, dtype=np.int64)
        interior = (kk >= 1) & (ii >= 1) & (jj >= 1)
        G[interior] = self.base_face + fi * self.n_face_int + self.row_off[ii[interior]] + (jj[interior] - 1)
        G[0, 0], G[f, 0], G[0, f] = A, B, C
        if f > 1:
            t = np.arange(1, f)
            G[t, 0] = self._edge_point(A, B, t)          # j = 0 : A -> B, parameter i
            G[0, t] = self._edge_point(A, C, t)          # i = 0 : A -> C, parameter j
            G[f - t, t] = self._edge_point(B, C, t)      # k = 0 : B -> C, parameter j
        return G

    def face_triangles(self, fi: int) -> np.ndarray:
        """f^2 x 3 global vertex ids of the triangles of face ``fi`` in element order (unsorted ids)."""
        f = self.f
        G = self.face_grid(fi)
        ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
        iu, ju = np.nonzero((ii + jj) <= f - 1)
        up = np.stack([G[iu, ju], G[iu + 1, ju], G[iu, ju + 1]], axis=1)
        idn, jdn = np.nonzero((ii + jj) <= f - 2)
        dn = np.stack([G[idn + 1, jdn], G[idn + 1, jdn + 1], G[idn, jdn + 1]], axis=1)
        key = np.concatenate([iu * (2 * f + 2) + 2 * ju, idn * (2 * f + 2) + 2 * jdn + 1])
        return np.concatenate([up, dn])[np.argsort(key, kind="stable")]

    def positions(self, ids: np.ndarray) -> np.ndarray:
        """Unit-sphere positions of the given global vertex ids (any order, any subset)."""
        f = self.f
        ids = np.asarray(ids, dtype=np.int64)
        out = np.empty((ids.shape[0], 3), dtype=np.float64)
        c = self.corners
        is_corner = ids < self.base_edge
        out[is_corner] = c[ids[is_corner]]
        is_edge = (ids >= self.base_edge) & (ids < self.base_face)
        if is_edge.any():
            loc = ids[is_edge] - self.base_edge
            k, t = loc // self.n_edge_int, (loc % self.n_edge_int + 1).astype(np.float64)[:, None] / f
            out[is_edge] = (1.0 - t) * c[self.edges[k, 0]] + t * c[self.edges[k, 1]]
        is_face = ids >= self.base_face
        if is_face.any():
            loc = ids[is_face] - self.base_face
            fi, r = loc // self.n_face_int, loc % self.n_face_int
            i = np.searchsorted(self.row_off, r, side="right") - 1
            j = r - self.row_off[i] + 1
            w = np.stack([f - i - j, i, j], axis=1).astype(np.float64) / f
            tri = _ICO_FACES[fi]
            out[is_face] = w[:, 0:1] * c[tri[:, 0]] + w[:, 1:2] * c[tri[:, 1]] + w[:, 2:3] * c[tri[:, 2]]
        out /= np.linalg.norm(out, axis=1, keepdims=True)
        return out


def icosphere(f: int, sort_ids: bool = True):
    """Geodesic icosphere of frequency ``f`` on the unit sphere.

    ``20 f^2`` triangles, ``10 f^2 + 2`` vertices (BASELINE.json config 2 uses
    f = 707: 9 996 980 triangles, 4 998 492 vertices).  Ordering: see :class:`IcoTopology`.
    """
    topo = IcoTopology(f)
    conn = np.concatenate([topo.face_triangles(fi) for fi in range(20)]).astype(np.int32)
    assert conn.shape[0] == 20 * f * f and conn.min() >= 0
    vert = topo.positions(np.arange(topo.nv))
    if sort_ids:
        conn.sort(axis=1)
    return np.ascontiguousarray(vert), np.ascontiguousarray(conn)


def radial_factors(nv: int, amplitude: float = 0.01, seed: int = 12345) -> np.ndarray:
    """``1 + amplitude * u_i`` for all ``nv`` vertices of a mesh (shared by the partitioned builder)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return 1.0 + amplitude * rng.uniform(-1.0, 1.0, size=nv)


def perturb_radially(vert: np.ndarray, amplitude: float = 0.01, seed: int = 12345) -> np.ndarray:
    """``x -> x (1 + amplitude * u)``, ``u ~ U(-1, 1)``, PCG64(seed)."""
    return np.ascontiguousarray(vert * radial_factors(vert.shape[0], amplitude, seed)[:, None])


def disk(n: int, dim: int = 3):
    """Structured triangulated unit disk: ``n`` rings; ring r has 6 r vertices.

    ``6 n^2`` triangles, ``3 n (n+1) + 1`` vertices, z = 0 when ``dim == 3``
    (layout of examples/tactoid/disk.mesh: dim 3, planar).  Returns
    ``(vert, faces, boundary_edges)`` with sorted ids.
    """
    ring_start = [0] + [1 + 3 * r * (r - 1) for r in range(1, n + 2)]
    nv = 1 + 3 * n * (n + 1)
    vert = np.zeros((nv, dim), dtype=np.float64)
    for r in range(1, n + 1):
        m = 6 * r
        th = 2.0 * np.pi * np.arange(m) / m
        vert[ring_start[r]: ring_start[r] + m, 0] = (r / n) * np.cos(th)
        vert[ring_start[r]: ring_start[r] + m, 1] = (r / n) * np.sin(th)
    tris = []
    for r in range(1, n + 1):
        m_out, m_in = 6 * r, max(6 * (r - 1), 1)
        so, si = ring_start[r], ring_start[r - 1]
        k = np.arange(m_out)
        sector, pos = k // r, k % r
        a = so + k
        b = so + (k + 1) % m_out
        inner = si + ((sector * (r - 1) + pos) % m_in if r > 1 else 0 * k)
        tris.append(np.stack([a, b, inner], axis=1))
        if r > 1:
            sel = pos < r - 1
            k2 = k[sel]
            inner_a = si + (sector[sel] * (r - 1) + pos[sel]) % m_in
            inner_b = si + (sector[sel] * (r - 1) + pos[sel] + 1) % m_in
            tris.append(np.stack([so + (k2 + 1) % m_out, inner_b, inner_a], axis=1))
    faces = np.concatenate(tris).astype(np.int32)
    assert faces.shape[0] == 6 * n * n
    faces.sort(axis=1)
    m = 6 * n
    k = np.arange(m)
    bnd = np.stack([ring_start[n] + k, ring_start[n] + (k + 1) % m], axis=1).astype(np.int32)
    bnd.sort(axis=1)
    return vert, np.ascontiguousarray(faces), np.ascontiguousarray(bnd)


def tet_cube(n: int):
    """Unit cube split into ``n^3`` cells x 6 Kuhn tetrahedra; ``(n+1)^3`` vertices."""
    g = np.arange(n + 1, dtype=np.float64) / n
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    vert = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    s = n + 1

    def vid(i, j, k):
        return (i * s + j) * s + k

    I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for p in perms:
        d = np.zeros((3, 3), dtype=np.int64)
        for step, ax in enumerate(p):
            d[step:, ax] = 1
        v0 = vid(I, J, K)
        v1 = vid(I + d[0, 0], J + d[0, 1], K + d[0, 2])
        v2 = vid(I + d[1, 0], J + d[1, 1], K + d[1, 2])
        v3 = vid(I + 1, J + 1, K + 1)
        tets.append(np.stack([v0, v1, v2, v3], axis=1))
    # interleave the six tets of a cell so that element order follows cell order
    conn = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    conn.sort(axis=1)
    return np.ascontiguousarray(vert), np.ascontiguousarray(conn)


def polyline_loop(n: int, dim: int = 3, seed: int = 7, noise: float = 0.05):
    """Closed perturbed circle of ``n`` points; edges (i, i+1 mod n) with sorted ids."""
    rng = np.random.Generator(np.random.PCG64(seed))
    th = 2.0 * np.pi * np.arange(n) / n
    vert = np.zeros((n, dim), dtype=np.float64)
    vert[:, 0], vert[:, 1] = np.cos(th), np.sin(th)
    vert += noise * rng.uniform(-1, 1, size=vert.shape) * (np.arange(dim) < 3)[None, :]
    k = np.arange(n)
    edges = np.stack([k, (k + 1) % n], axis=1).astype(np.int32)
    edges.sort(axis=1)
    return vert, np.ascontiguousarray(edges)


def sphere_points(n: int, seed: int = 12345):
    """``n`` points uniform on S^2 (normalised Gaussians, PCG64(seed)); SURVEY §8d config 3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal((n, 3))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x)


def edges_from_faces(conn: np.ndarray) -> np.ndarray:
    """Unique edges of a simplicial complex in the order ``mesh_addgrade`` creates them.

    Reference: mesh.c:419-489 walks the higher-grade elements in id order and for
    each emits the (g+1 choose 2) vertex pairs in lexicographic counter order,
    keeping first occurrences only.
    """
    nvper = conn.shape[1]
    pairs = [(a, b) for a in range(nvper) for b in range(a + 1, nvper)]
    cand = np.stack([conn[:, [a, b]] for a, b in pairs], axis=1).reshape(-1, 2).astype(np.int64)
    key = cand[:, 0] * (int(conn.max()) + 1) + cand[:, 1]
    _, first = np.unique(key, return_index=True)
    first.sort()
    return np.ascontiguousarray(cand[first].astype(np.int32))


def write_mesh_file(path, vert, **grades):
    """Write Morpho's ``.mesh`` text format (mesh.c:910-1080): 1-based ids."""
    names = {1: "edges", 2: "faces", 3: "volumes"}
    with open(path, "w") as fh:
        fh.write("vertices\n\n")
        for i, x in enumerate(vert):
            fh.write(f"{i + 1} " + " ".join(repr(float(c)) for c in x) + "\n")
        for key, conn in grades.items():
            g = int(key[1:]) if key.startswith("g") else int(key)
            fh.write(f"\n{names[g]}\n\n")
            for i, el in enumerate(conn):
                fh.write(f"{i + 1} " + " ".join(str(int(c) + 1) for c in el) + "\n")
        fh.write("\n")      # the reference's loader repeats the last entry of a file that ends without a blank line
