"""CPU: the C-ABI library loads without a GPU, exports every symbol include/morpho_cuda.h declares,
refuses to compute without a device (no CPU fallback), and its host-side logic (patch planner,
binning) is correct.  No compute entry point is exercised here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "morpho_b200", "lib", "libmorpho_cuda.so")
HDR = os.path.join(ROOT, "include", "morpho_cuda.h")


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    import morpho_b200 as mb
    return mb.lib()


def test_header_symbols_exported(L):
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mcu_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) >= 35
    exported = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    names = {l.split()[-1] for l in exported.splitlines() if l.strip()}
    missing = [d for d in declared if d not in names]
    assert not missing, f"declared but not exported: {missing}"
    for d in declared:
        assert hasattr(L, d)


def test_abi_version_and_error_ids(L):
    assert L.mcu_abi_version() == 1
    assert L.mcu_error_id(4) == b"VolEnclZero"
    assert L.mcu_error_id(2) == b"FnctlELNtFnd"
    assert L.mcu_error_id(0) == b""


def test_no_cpu_fallback_without_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import morpho_b200 as mb
    with pytest.raises(mb.MorphoError) as ei:
        mb.Mesh(np.zeros((3, 3)), faces=[[0, 1, 2]])
    assert ei.value.status == 6 and "no CPU fallback" in str(ei.value)


def test_binbounds_matches_reference_scheme(L, oracle):
    for nel, nb in ((10, 3), (9, 3), (2, 4), (9996980, 8), (0, 2)):
        b = np.zeros(nb + 1, dtype=np.int32)
        assert L.mcu_binbounds(nel, nb, b.ctypes.data) == 0
        np.testing.assert_array_equal(b, oracle.binbounds(nel, nb))


H_RUN_OFF, H_N_RUNS, H_TAB_OFF, H_TAB_BYTES, H_N_OWNED, H_N_SLOTS, H_N_SLICES, H_TX_BYTES = range(8)
ENT_START, ENT_SLOT = 0x8000, 0x03FF


def _plan(L, v, c, max_owned=512):
    counts = (C.c_long * 10)()
    rc = L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, counts, None, None, None)
    assert rc == 0
    npatch, nruns, ntab, hints = (int(x) for x in list(counts)[:4])
    assert hints == 8
    hdr = np.zeros((npatch, hints), dtype=np.int32)
    runs = np.zeros(nruns, dtype=np.dtype([("gid", "<i4"), ("slot", "<u2"), ("n", "<u2")]))
    tab = np.zeros(ntab, dtype=np.uint32)
    rc = L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, counts,
                                hdr.ctypes.data, runs.ctypes.data, tab.ctypes.data)
    assert rc == 0
    return dict(hdr=hdr, runs=runs, tab=tab, counts=[int(x) for x in counts])


def _decode_patch(P, h):
    """slot → global id map (-1 for alignment slots are real ids too: every slot of a run is a vertex row),
    owned ids per thread, and the entry columns of every owned vertex."""
    run_off, n_runs, tab_off, tab_bytes, n_owned, n_slots, n_slices, tx = (int(x) for x in h)
    runs = P["runs"][run_off: run_off + n_runs]
    slot_gid = np.full(n_slots, -1, dtype=np.int64)
    at = 0
    for r in runs:
        assert r["gid"] % 2 == 0 and r["n"] % 2 == 0 and r["n"] > 0 and r["slot"] == at   # 16-byte aligned bulk copies
        slot_gid[at: at + r["n"]] = np.arange(r["gid"], r["gid"] + r["n"])
        at += int(r["n"])
    assert at == n_slots and n_slots <= 1024
    assert tab_bytes % 16 == 0 and tx == 24 * n_slots + tab_bytes
    tab = P["tab"][4 * tab_off: 4 * tab_off + tab_bytes // 4]
    assert int(tab[1]) == n_slices and n_slices == (n_owned + 31) // 32
    patch_index = int(tab[0])
    gids = tab[4: 4 + 32 * n_slices].view(np.int32)
    noff = (n_slices + 4) & ~3
    offs_raw = tab[4 + 32 * n_slices: 4 + 32 * n_slices + noff].view(np.int32)[: n_slices + 1]
    offs = offs_raw & ~31                                      # bit 0 of an offset marks a CLEAN slice
    ents = tab[4 + 32 * n_slices + noff:].view(np.uint16)[: int(offs[-1])]
    assert int(tab[2]) == 0 and int(tab[3]) == 0              # push list of the multi-GPU exchange: none attached
    assert offs[0] == 0 and len(ents) == offs[-1] and np.all(np.diff(offs) % 32 == 0) and np.all(np.diff(offs) >= 128)
    for sl in range(n_slices):                                  # CLEAN: 32 real vertices, no START flag and no padding after entry 1
        blk = ents[offs[sl]: offs[sl + 1]].reshape(-1, 32)
        clean = bool(np.all(gids[32 * sl: 32 * sl + 32] >= 0)) and not np.any(blk[2:] & 0x8000)
        assert bool(offs_raw[sl] & 1) == clean, (sl, clean)
    return slot_gid, gids, offs, ents, patch_index


@pytest.mark.parametrize("f,max_owned", [(3, 512), (12, 512), (30, 200), (30, 512)])
def test_patch_plan_is_a_valid_owner_computes_decomposition(L, oracle, f, max_owned):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(f)
    v = mg.perturb_radially(v)
    P = _plan(L, v, c, max_owned)
    nv = v.shape[0]
    tptr, tidx = oracle.transpose(nv, c)
    owned_count = np.zeros(nv, dtype=np.int64)
    for hi, h in enumerate(P["hdr"]):
        n_owned, n_slices = int(h[H_N_OWNED]), int(h[H_N_SLICES])
        assert 0 < n_owned <= max_owned
        slot_gid, gids, offs, ents, pidx = _decode_patch(P, h)
        assert pidx == hi
        assert np.all(gids[n_owned:] == -1) and np.all(gids[:n_owned] >= 0)
        assert np.all(np.diff(gids[:n_owned]) > 0)          # threads in ascending id order → coalesced row writes
        owned_count[gids[:n_owned]] += 1
        # vertex-centric view: sliced ELL of 16-bit STRIPS.  Entry 0 is the owner's slot, then ring vertices;
        # consecutive ring entries (without the START bit on the second) span one triangle with the owner.  The
        # multiset of triangles decoded must be exactly the owner's incident elements.
        for i in range(32 * n_slices):
            s, lane = divmod(i, 32)
            column = ents[offs[s] + lane: offs[s + 1]: 32].astype(np.int64)
            assert np.all((column & ENT_SLOT) < int(h[H_N_SLOTS]))
            if i >= n_owned:
                assert np.all(column[1:] & ENT_START)       # idle lanes contribute nothing
                continue
            u = int(gids[i])
            assert slot_gid[column[0]] == u and not (column[0] & ENT_START)
            got = []
            for k2 in range(2, len(column)):
                if column[k2] & ENT_START:
                    continue
                a, b2 = int(slot_gid[column[k2 - 1] & ENT_SLOT]), int(slot_gid[column[k2] & ENT_SLOT])
                got.append(tuple(sorted((u, a, b2))))
            want = [tuple(int(x) for x in c[e]) for e in tidx[tptr[u]: tptr[u + 1]]]
            assert sorted(got) == sorted(want)
            if len(want) in (5, 6):
                assert int(np.sum((column[2:] & ENT_START) == 0)) == len(want)   # a closed fan is one strip
    assert np.all(owned_count == 1)
    assert P["counts"][9] > 0 and P["counts"][8] >= P["counts"][9]


def test_patch_plan_handles_isolated_vertices(L):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(6)
    v = np.concatenate([v, [[3.0, 3.0, 3.0], [-2.0, 0.1, 0.4]]])   # two vertices without elements
    P = _plan(L, np.ascontiguousarray(v), c)
    owned = []
    for h in P["hdr"]:
        slot_gid, gids, offs, ents, _ = _decode_patch(P, h)
        owned.extend(int(g) for g in gids if g >= 0)
    assert sorted(owned) == list(range(v.shape[0]))


@pytest.mark.parametrize("f,max_owned", [(3, 512), (12, 512), (30, 200)])
def test_patch_triangle_lists_cover_every_triangle_once(L, f, max_owned):
    """The totals walk per-patch triangle lists: every triangle of the mesh appears in exactly one list (the patch
    that owns its smallest vertex), with the slots of its three vertices in ascending global id."""
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(f)
    v = mg.perturb_radially(v)
    P = _plan(L, v, c, max_owned)
    nw = (C.c_long * 1)()
    assert L.mcu_debug_patch_triangles(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, nw, None, None) == 0
    th = np.zeros((len(P["hdr"]), 4), dtype=np.int32)
    tri = np.zeros(int(nw[0]), dtype=np.uint32)
    assert L.mcu_debug_patch_triangles(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, nw, th.ctypes.data, tri.ctypes.data) == 0
    seen = {}
    key = {tuple(t): i for i, t in enumerate(c.tolist())}
    for pi, h in enumerate(P["hdr"]):
        slot_gid, gids, offs, ents, patch_index = _decode_patch(P, h)
        off, nbytes, n_tri, _ = (int(x) for x in th[pi])
        words = tri[4 * off: 4 * off + nbytes // 4]
        assert int(words[0]) == pi and int(words[1]) == n_tri and nbytes % 16 == 0 and nbytes >= 16 + 8 * n_tri
        e = words[4: 4 + 2 * n_tri].view(np.uint16).reshape(n_tri, 4)
        owned = set(int(g) for g in gids if g >= 0)
        for s0, s1, s2, z in e.tolist():
            t = (int(slot_gid[s0]), int(slot_gid[s1]), int(slot_gid[s2]))
            assert z == 0 and t[0] < t[1] < t[2] and t[0] in owned
            assert t in key and key[t] not in seen
            seen[key[t]] = pi
    assert len(seen) == c.shape[0]


# ---------------------------------------------------------------------------------------------------------------
# Block decomposition of the generic owner-computes gradient kernel (morpho_b200/csrc/mcu_block.cu): host logic.
# ---------------------------------------------------------------------------------------------------------------

def _block_plan(L, v, conn, eids=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    e = None if eids is None else np.ascontiguousarray(eids, dtype=np.int32)
    n = conn.shape[0] if e is None else len(e)
    counts = (C.c_long * 5)()
    args = (v.shape[0], v.shape[1], n, conn.shape[1], conn.ctypes.data, None if e is None else e.ctypes.data, v.ctypes.data, counts)
    assert L.mcu_debug_block_plan(*args, None, None, None) == 0
    nb, nloc, nent, nact, allact = (int(x) for x in counts)
    hdr = np.zeros((nb, 14), dtype=np.int32)
    loc = np.zeros(max(nloc, 1), dtype=np.int32)
    ent = np.zeros(max(nent, 1), dtype=np.uint32)
    assert L.mcu_debug_block_plan(*args, hdr.ctypes.data, loc.ctypes.data, ent.ctypes.data) == 0
    return hdr, loc[:nloc], ent[:nent], nact, bool(allact)


def _emulate_block_gradient(name, v, hdr, loc, ent):
    """What k_block_grad computes, in numpy, straight from the plan: every owned vertex walks its entries."""
    nv, dim = v.shape
    x3 = np.concatenate([v, np.zeros((nv, 3 - dim))], axis=1)
    out = np.zeros((nv, 3))
    written = np.zeros(nv, dtype=int)
    ntets = np.zeros(nv, dtype=int)
    for h in hdr:
        loc_off, n_local, n_owned = int(h[0]), int(h[1]), int(h[2])
        ent_off = h[4:12].view(np.uint32)
        deg = h[12:14].view(np.uint8)
        assert n_owned <= 256 and n_local <= 1023 and n_owned <= n_local
        ids = loc[loc_off: loc_off + n_local]
        assert len(set(ids.tolist())) == n_local                        # every local vertex has one slot
        assert np.all(np.diff(ids[:n_owned]) > 0) and np.all(np.diff(ids[n_owned:]) > 0)
        X = x3[ids]
        for t in range(n_owned):
            w, lane = divmod(t, 32)
            p = X[t]
            acc = np.zeros(3)
            if name == "Volume":
                # tetrahedra: 16-bit entries walk the link of the vertex as triangle strips (mcu_block.cu link_strips)
                e16 = ent.view(np.uint16)
                a = b = c = np.zeros(3)
                ntri = 0
                assert int(ent_off[w]) % 128 == 0                            # 8-byte loads of four entries per lane
                for j in range(4 * int(deg[w])):                               # deg counts groups of four entries
                    e = int(e16[int(ent_off[w]) + ((j // 4) * 32 + lane) * 4 + (j & 3)])
                    op, s = e >> 10, e & 1023
                    a, b, c = (a if op == 1 else b), c, X[s]
                    if op < 2:
                        N = np.cross(b - a, c - a)
                        acc += np.sign((p - a) @ N) * N / 6
                        ntri += 1
                out[ids[t]] = acc
                written[ids[t]] += 1
                ntets[ids[t]] = ntri
                continue
            for j in range(int(deg[w])):
                e = int(ent[int(ent_off[w]) + 32 * j + lane])
                if e == 0xffffffff:
                    continue
                a, b, c = X[e & 1023], X[(e >> 10) & 1023], X[(e >> 20) & 1023]
                if name == "Length":
                    acc += (p - a) / np.linalg.norm(p - a)
                elif name == "Area":
                    N = np.cross(a - p, b - p)
                    acc += np.cross(N, (b - p) - (a - p)) / (2 * np.linalg.norm(N))
                elif name == "VolumeEnclosed":
                    cx = np.cross(a, b)
                    acc += np.sign(p @ cx) * cx / 6
                else:
                    N = np.cross(b - a, c - a)
                    acc += np.sign((p - a) @ N) * N / 6
            out[ids[t]] = acc
            written[ids[t]] += 1
    if name == "Volume":
        return out[:, :dim], written, ntets
    return out[:, :dim], written


@pytest.mark.parametrize("case", ["tets", "tets_selected", "triangles_selected", "lines_2d", "sphere_volencl"])
def test_block_plan_reproduces_the_oracle_gradient(L, case):
    """The plan alone determines the result: emulating the kernel from (headers, local lists, entries) gives the
    oracle's gradient (1e-12), every vertex with elements is owned exactly once, and vertices without (selected)
    elements are owned by nobody."""
    from oracle.oracle import FunctionalSpec, Oracle
    from morpho_b200 import meshgen as mg
    O = Oracle()
    rng = np.random.Generator(np.random.PCG64(7))
    sel = None
    if case in ("tets", "tets_selected"):
        v, c = mg.tet_cube(7)
        v = v + 0.02 * rng.uniform(-1, 1, size=v.shape)
        name, g = "Volume", 3
        if case == "tets_selected":
            sel = np.sort(rng.choice(c.shape[0], size=c.shape[0] // 2, replace=False)).astype(np.int32)
    elif case == "triangles_selected":
        v, c = mg.icosphere(9)
        v = mg.perturb_radially(v, 0.05, 3)
        sel = np.sort(rng.choice(c.shape[0], size=c.shape[0] // 3, replace=False)).astype(np.int32)
        name, g = "Area", 2
    elif case == "lines_2d":
        v, c = mg.polyline_loop(300, 2, 5)
        name, g = "Length", 1
    else:
        v, c = mg.icosphere(12)
        v = mg.perturb_radially(v, 0.05, 4)
        name, g = "VolumeEnclosed", 2
    hdr, loc, ent, nact, allact = _block_plan(L, v, c, sel)
    res = _emulate_block_gradient(name, v, hdr, loc, ent)
    got, written = res[0], res[1]
    if name == "Volume":                                   # every incident tetrahedron is visited exactly once by the strips
        cnt = np.bincount((c if sel is None else c[sel]).ravel(), minlength=v.shape[0])
        assert np.array_equal(res[2], cnt)
    used = np.zeros(v.shape[0], dtype=bool)
    used[np.unique(c if sel is None else c[sel])] = True
    assert np.array_equal(written == 1, used) and written.max() <= 1 and nact == used.sum() and allact == bool(used.all())
    want = O.gradient(v, {g: c}, FunctionalSpec(name), sel)
    err = np.linalg.norm(got - want, axis=1).max() / np.linalg.norm(want, axis=1).max()
    assert err <= 1e-12, err


def test_patch_plan_does_not_depend_on_the_number_of_threads(L):
    """The planner cuts the top of the bisection tree level by level and builds the groups of patches on all host cores
    (mcu_patch.cu); the pieces are concatenated in group order: headers, runs and tables are identical to the bit for
    1, 3 and the default number of threads, and to the digest recorded from the single-threaded planner of round 1."""
    import hashlib
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(60)
    v = np.ascontiguousarray(mg.perturb_radially(v))
    c = np.ascontiguousarray(np.sort(c, axis=1).astype(np.int32))
    digests = []
    try:
        for nt in (1, 3, 0):
            assert L.mcu_set_option(b"plan_threads", nt) == 0
            P = _plan(L, v, c, 0)
            h = hashlib.sha256()
            for a in (P["hdr"], P["runs"], P["tab"]):
                h.update(np.ascontiguousarray(a).tobytes())
            digests.append((h.hexdigest(), P["counts"]))
    finally:
        L.mcu_set_option(b"plan_threads", 0)
    assert digests[0] == digests[1] == digests[2], digests
    for pi, h in enumerate(P["hdr"]):                      # the patch index stored in every staged table is global
        assert int(P["tab"][4 * int(h[H_TAB_OFF])]) == pi
