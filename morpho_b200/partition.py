"""Element partition of a mesh across GPUs (one process per GPU) and the exchange of the gradient
rows of vertices shared between ranks.

Scheme = the reference's own work split: contiguous element-id ranges with the remainder spread over
the first bins (functional_binbounds, functional.c:863-879; the reference gives each *thread* a bin and a
private full-size gradient matrix which it merges with matrix_axpy, functional.c:1092-1135).  Here each
rank holds only its element range and the vertices those elements touch, renumbered compactly in
ascending global id; only the rows of vertices touched by MORE than one rank are communicated:

    local partial gradient  --pack shared rows-->  all_gather (NCCL over NVLink / gloo in CPU tests)
                            --sum over ranks in fixed rank order-->  unpack into the local gradient

so every rank ends with the complete gradient rows of all vertices it touches, bitwise identical on all
ranks that share a vertex (fixed summation order).  Indexing (ranges, local<->global maps, shared lists)
is integer-exact and tested against the oracle's restatement of functional_binbounds.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import meshgen as mg

__all__ = ["binbounds", "RankMesh", "build_rank_mesh", "partition_arrays", "shared_vertices", "halo_plan", "PeerExchange", "reduce_total"]


def binbounds(nel: int, nbins: int) -> np.ndarray:
    """functional_binbounds (functional.c:863-879): ``nel // nbins`` per bin, remainder to the first bins."""
    base, rem = divmod(int(nel), int(nbins))
    sizes = np.full(nbins, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


@dataclass
class RankMesh:
    rank: int
    world: int
    nel_global: int
    nv_global: int
    elem_range: tuple            # [lo, hi) of global element ids held by this rank
    gids: np.ndarray             # local vertex -> global vertex id (ascending)
    vert: np.ndarray             # (nv_local, dim)
    conn: np.ndarray             # (nel_local, g+1) local ids, ascending per element
    shared_local: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # local ids of shared vertices
    shared_slot: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))    # their row in the exchange buffer
    n_slots: int = 0             # rows of the exchange buffer (all vertices shared by >= 2 ranks)
    rims: list = field(default_factory=list)   # per rank: sorted global ids it may share (superset of its shared ids)

    @property
    def n_shared(self):
        return int(self.shared_local.shape[0])

    def make_exchange(self, torch, dist):
        """Return ``exchange(*grads)``: completes the shared-vertex rows of local gradient tensors
        (flat float64 tensors of nv_local*dim, on the device the process group works with)."""
        if self.world == 1 or self.n_slots == 0:
            return lambda *g: None
        dim = self.vert.shape[1]
        idx = None
        buf = None
        gathered = None

        def exchange(*grads):
            nonlocal idx, buf, gathered
            dev = grads[0].device
            k = len(grads)
            if idx is None or idx.device != dev or buf.shape[1] != k * dim:
                idx = torch.as_tensor(self.shared_local, dtype=torch.int64, device=dev)
                slot = torch.as_tensor(self.shared_slot, dtype=torch.int64, device=dev)
                buf = torch.zeros((self.n_slots, k * dim), dtype=torch.float64, device=dev)
                gathered = torch.empty((self.world, self.n_slots, k * dim), dtype=torch.float64, device=dev)
                exchange.slot = slot
            slot = exchange.slot
            rows = torch.cat([g.view(-1, dim).index_select(0, idx) for g in grads], dim=1)
            buf.zero_()
            buf.index_copy_(0, slot, rows)
            dist.all_gather_into_tensor(gathered.view(-1), buf.view(-1))
            total = gathered[0].clone()
            for r in range(1, self.world):          # fixed rank order: deterministic and identical everywhere
                total += gathered[r]
            mine = total.index_select(0, slot)
            for i, g in enumerate(grads):
                g.view(-1, dim).index_copy_(0, idx, mine[:, i * dim:(i + 1) * dim].contiguous())

        return exchange


def shared_vertices(touched_per_rank):
    """Given the sorted global ids touched by each rank, return (S_all, per-rank index arrays):
    S_all = sorted ids touched by >= 2 ranks; per rank the positions of its touched ids that are in S_all."""
    allids = np.concatenate(touched_per_rank)
    uniq, cnt = np.unique(allids, return_counts=True)
    s_all = uniq[cnt >= 2]
    return s_all


def _finish(rm: RankMesh, s_all: np.ndarray) -> RankMesh:
    pos = np.searchsorted(rm.gids, s_all)
    pos = np.clip(pos, 0, max(len(rm.gids) - 1, 0))
    hit = rm.gids[pos] == s_all if len(rm.gids) else np.zeros(len(s_all), bool)
    rm.shared_local = pos[hit].astype(np.int64)
    rm.shared_slot = np.nonzero(hit)[0].astype(np.int64)
    rm.n_slots = int(len(s_all))
    return rm


def partition_arrays(vert: np.ndarray, conn: np.ndarray, world: int):
    """Partition explicit arrays across ``world`` ranks (used by tests and small meshes).
    Returns the list of RankMesh for all ranks."""
    nel = conn.shape[0]
    b = binbounds(nel, world)
    out, touched = [], []
    for r in range(world):
        c = conn[b[r]: b[r + 1]]
        gids = np.unique(c)
        loc = np.searchsorted(gids, c).astype(np.int32)
        out.append(RankMesh(r, world, nel, vert.shape[0], (int(b[r]), int(b[r + 1])), gids.astype(np.int64),
                            np.ascontiguousarray(vert[gids]), np.ascontiguousarray(np.sort(loc, axis=1))))
        touched.append(gids)
    s_all = shared_vertices(touched)
    for rm in out:
        rm.rims = touched
    return [_finish(rm, s_all) for rm in out]


def build_rank_mesh(f: int, rank: int, world: int, amplitude: float = 0.01, seed: int = 12345, dist=None, torch=None) -> RankMesh:
    """This rank's element range of the radially perturbed icosphere of frequency ``f``, generated face by
    face (no rank materialises the global mesh).  The shared-vertex list is agreed with one all_gather
    of the candidate ids when ``world > 1`` (``torch.distributed`` must be initialised)."""
    topo = mg.IcoTopology(f)
    b = binbounds(topo.nel, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    tpf = topo.tri_per_face
    parts = []
    for fi in range(lo // tpf, min(19, (hi - 1) // tpf) + 1 if hi > lo else 0):
        tri = topo.face_triangles(fi)
        a, z = max(lo - fi * tpf, 0), min(hi - fi * tpf, tpf)
        parts.append(tri[a:z])
    cg = np.concatenate(parts) if parts else np.zeros((0, 3), np.int64)
    gids = np.unique(cg)
    vert = topo.positions(gids) * mg.radial_factors(topo.nv, amplitude, seed)[gids][:, None]
    conn = np.sort(np.searchsorted(gids, cg).astype(np.int32), axis=1)
    rm = RankMesh(rank, world, topo.nel, topo.nv, (lo, hi), gids.astype(np.int64), np.ascontiguousarray(vert),
                  np.ascontiguousarray(conn))
    if world == 1:
        return rm
    # Candidates for sharing: vertices of this rank that also occur in an element outside its range.  A vertex
    # of the icosphere is shared only if it lies on the border of the rank's element set; rather than deriving
    # that geometrically every rank publishes the ids on the rim of its range: vertices whose incident
    # element count inside the range is smaller than their valence on the closed sphere (6, or 5 at the 12 corners).
    if dist is None:
        import torch as _torch
        import torch.distributed as _dist
        torch, dist = _torch, _dist
    valence = np.where(gids < 12, 5, 6)
    inside = np.bincount(conn.ravel(), minlength=len(gids))
    rim = gids[inside < valence].astype(np.int64)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = torch.tensor([len(rim)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    nmax = int(max(int(s.item()) for s in sizes))
    pad = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    pad[: len(rim)] = torch.as_tensor(rim, device=dev)
    allr = [torch.empty(nmax, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allr, pad)
    rims = [a.cpu().numpy()[: int(s.item())] for a, s in zip(allr, sizes)]
    s_all = shared_vertices(rims)
    rm.rims = rims
    return _finish(rm, s_all)


def halo_plan(rm: RankMesh) -> dict:
    """Index arrays of the peer-memory exchange (include/morpho_cuda.h, mcu_halo_create) for one rank.

    Every pair of ranks (r, q) agrees on the list of vertices they share, in ascending global id; entry k of
    that list is where r's partial row lands in q's receive buffer and vice versa.  Per shared vertex the
    contributions are listed in ascending rank order (own partial at its place, k = -1): the summation order
    is the same on every rank that holds the vertex, so the completed rows are bitwise identical everywhere.
    All arrays are int32; everything is derived from the ranks' rim lists, which every rank knows."""
    world, r = rm.world, rm.rank
    order = np.argsort(rm.gids[rm.shared_local], kind="stable")
    shared_local = rm.shared_local[order].astype(np.int32)
    shared_gid = rm.gids[shared_local]
    rims = [np.asarray(x, dtype=np.int64) for x in rm.rims]
    common = [np.intersect1d(rims[r], rims[q]) if q != r else np.zeros(0, np.int64) for q in range(world)]
    peer_ptr = np.zeros(world + 1, dtype=np.int32)
    peer_pos = []
    for q in range(world):
        pos = np.searchsorted(shared_gid, common[q])
        assert np.array_equal(shared_gid[pos], common[q]) if len(pos) else True
        peer_pos.append(pos.astype(np.int32))
        peer_ptr[q + 1] = peer_ptr[q] + len(pos)
    max_m = 0
    for a in range(world):
        for b in range(a + 1, world):
            max_m = max(max_m, len(np.intersect1d(rims[a], rims[b])))
    # contributions per shared vertex: which ranks hold it, ascending, and where their row sits
    member = np.zeros((world, len(shared_gid)), dtype=bool)
    kidx = np.full((world, len(shared_gid)), -1, dtype=np.int64)
    for q in range(world):
        if q == r:
            member[q] = True
            continue
        pos = peer_pos[q]
        member[q, pos] = True
        kidx[q, pos] = np.arange(len(pos))
    counts = member.sum(axis=0)
    contrib_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    qq, ii = np.nonzero(member)                       # row-major: sorted by rank, then vertex
    o = np.lexsort((qq, ii))                          # by vertex, then ascending rank
    contrib_rank = qq[o].astype(np.int32)
    contrib_k = kidx[qq[o], ii[o]].astype(np.int32)
    return dict(shared_local=np.ascontiguousarray(shared_local), peer_ptr=peer_ptr,
                peer_pos=np.ascontiguousarray(np.concatenate(peer_pos) if peer_pos else np.zeros(0, np.int32)).astype(np.int32),
                max_m=int(max_m), contrib_ptr=contrib_ptr, contrib_rank=np.ascontiguousarray(contrib_rank),
                contrib_k=np.ascontiguousarray(contrib_k))


class PeerExchange:
    """Shared-vertex completion over NVLink peer memory (morpho_b200/csrc/mcu_halo.cu): one kernel per
    exchange, no collective call on the data path.  ``torch.distributed`` (NCCL) is used once, to hand the
    CUDA IPC handles of the receive blocks around."""

    def __init__(self, rm: RankMesh, torch, dist, ngrad: int = 2, dim: int = None):
        import ctypes as C
        from . import _lib
        self.L, self.C, self._lib = _lib.lib(), C, _lib
        # dim = doubles per vertex row of the tensors exchanged (mesh dimension for vertex gradients, psize for field gradients)
        self.rm, self.dim, self.ngrad = rm, (rm.vert.shape[1] if dim is None else int(dim)), ngrad
        pl = self.plan = halo_plan(rm)
        self.h = self.L.mcu_halo_create(rm.rank, rm.world, ngrad * self.dim, len(pl["shared_local"]),
                                        pl["shared_local"].ctypes.data, pl["peer_ptr"].ctypes.data, pl["peer_pos"].ctypes.data,
                                        pl["max_m"], pl["contrib_ptr"].ctypes.data, pl["contrib_rank"].ctypes.data,
                                        pl["contrib_k"].ctypes.data)
        if not self.h:
            raise RuntimeError("mcu_halo_create failed: " + self.L.mcu_last_error().decode())
        mine = np.zeros(64, dtype=np.uint8)
        _lib.check(self.L.mcu_halo_export(self.h, mine.ctypes.data))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(mine).to(dev)
        allh = [torch.empty_like(t) for _ in range(rm.world)]
        dist.all_gather(allh, t)
        handles = np.ascontiguousarray(np.concatenate([a.cpu().numpy() for a in allh]))
        _lib.check(self.L.mcu_halo_connect(self.h, handles.ctypes.data))
        dist.barrier()
        self._ptrs = (C.c_void_p * 4)()

    def __call__(self, *grads):
        for i, g in enumerate(grads):
            self._ptrs[i] = g.data_ptr()
        self._lib.check(self.L.mcu_halo_exchange(self.h, len(grads), self._ptrs, self.dim))

    def exchange_ptrs(self, *dev_ptrs):
        """The same for raw device pointers (gradients held in mcu_matrix objects)."""
        for i, p in enumerate(dev_ptrs):
            self._ptrs[i] = p
        self._lib.check(self.L.mcu_halo_exchange(self.h, len(dev_ptrs), self._ptrs, self.dim))

    def exchange_async(self, slot, dev_ptr):
        """One gradient (column block ``slot``) on the library's second stream: runs underneath the kernels queued next."""
        self._lib.check(self.L.mcu_halo_exchange_async(self.h, slot, dev_ptr, self.dim))

    def join(self):
        self._lib.check(self.L.mcu_halo_join(self.h))

    def attach(self, mesh):
        """Fused mode: the patch gradient kernel of ``mesh`` pushes the shared rows itself; call
        ``bind(slot)`` before evaluating the gradient that occupies column block ``slot``."""
        self._lib.check(self.L.mcu_halo_attach(self.h, mesh._h))

    def bind(self, slot):
        self._lib.check(self.L.mcu_halo_bind(self.h, slot))

    def status(self):
        return self.L.mcu_halo_status(self.h)

    def close(self):
        if self.h:
            self.L.mcu_halo_destroy(self.h)
            self.h = None


def reduce_total(local_total: float, torch, dist) -> float:
    """Energy of an element-partitioned mesh: every element belongs to exactly one rank, so the total is the sum of
    the ranks' totals.  The reference adds its per-thread partial sums in thread order (functional.c:962-996); here
    the per-rank totals are all-gathered (one double each) and added in ascending rank order on every rank —
    the same bits everywhere, independent of the collective's internal reduction order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_total)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([float(local_total)], dtype=torch.float64, device=dev)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    total = 0.0
    for p in parts:
        total += float(p.item())
    return total


def pairwise_partitioned(x: np.ndarray, mode: int, torch, dist, kind: int = 0, cutoff=None):
    """PairwisePotential on several GPUs: the N x N interaction matrix is cut into row blocks with the
    reference's own bin bounds (functional_binbounds), every rank holds all N positions and computes the rows of its
    block against all points (``mcu_pairwise_rows_device``).  No exchange on the data path: the blocks are
    concatenated with one all-gather at the end (integrand / gradient), the energy is the fixed-order sum of the
    per-rank partial energies (:func:`reduce_total`).  Returns the full result on every rank."""
    import ctypes as C
    from . import _lib
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n, dim = x.shape
    b = binbounds(n, world)
    r0, nr = int(b[rank]), int(b[rank + 1] - b[rank])
    L = _lib.lib()
    dx = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    ncol = {_lib.MCU_TOTAL: 0, _lib.MCU_INTEGRAND: 1, _lib.MCU_GRADIENT: dim}[mode]
    maxr = int(np.max(np.diff(b)))
    out = torch.zeros(max(maxr * max(ncol, 1), 1), dtype=torch.float64, device="cuda")
    _lib.check(L.mcu_pairwise_rows_device(n, dim, dx.data_ptr(), kind, None, C.c_double(float(cutoff) if cutoff else 0.0), mode, r0, nr, out.data_ptr()))
    L.mcu_sync()
    if mode == _lib.MCU_TOTAL:
        return reduce_total(float(out[0].item()), torch, dist)
    if world == 1:
        return out[: nr * ncol].cpu().numpy().reshape(n, ncol) if ncol > 1 else out[:nr].cpu().numpy()
    parts = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(parts, out)                       # equal-size padded blocks
    full = np.concatenate([parts[r][: int(b[r + 1] - b[r]) * ncol].cpu().numpy() for r in range(world)])
    return full.reshape(n, ncol) if ncol > 1 else full
