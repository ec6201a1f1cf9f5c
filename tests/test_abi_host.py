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


H_VTX_OFF, H_N_OWNED, H_N_LOCAL, H_ELEM_OFF, H_N_ELEM, H_COLOR_OFF, H_N_COLORS, H_SLICE_OFF, H_N_SLICES, H_ENT_OFF, H_N_ENT = range(11)
ENT_NONE, ENT_START = 0xFFFF, 0x8000


def _plan(L, v, c, max_owned=768):
    counts = (C.c_long * 8)()
    rc = L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, counts,
                                None, None, None, None, None, None, None)
    assert rc == 0
    npatch, nvtx, nel, ncol, maxcol, hints, nslice, nent = list(counts)
    hdr = np.zeros((npatch, hints), dtype=np.int32)
    vtx = np.zeros(nvtx, dtype=np.int32)
    el = np.zeros(nel, dtype=np.uint32)
    eid = np.zeros(nel, dtype=np.int32)
    col = np.zeros(ncol, dtype=np.int32)
    sl = np.zeros(nslice, dtype=np.int32)
    ent = np.zeros(nent, dtype=np.uint16)
    rc = L.mcu_debug_patch_plan(v.shape[0], c.shape[0], c.ctypes.data, v.ctypes.data, max_owned, counts,
                                hdr.ctypes.data, vtx.ctypes.data, el.ctypes.data, eid.ctypes.data, col.ctypes.data,
                                sl.ctypes.data, ent.ctypes.data)
    assert rc == 0
    return dict(hdr=hdr, vtx=vtx, el=el, eid=eid, col=col, sl=sl, ent=ent, maxcol=maxcol)


@pytest.mark.parametrize("f,max_owned", [(3, 768), (12, 768), (30, 200), (30, 768)])
def test_patch_plan_is_a_valid_owner_computes_decomposition(L, oracle, f, max_owned):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(f)
    v = mg.perturb_radially(v)
    P = _plan(L, v, c, max_owned)
    hdr, vtx, el, eid, col = P["hdr"], P["vtx"], P["el"], P["eid"], P["col"]
    nv, nel = v.shape[0], c.shape[0]
    tptr, tidx = oracle.transpose(nv, c)
    owned_count = np.zeros(nv, dtype=np.int64)
    # every vertex→element incidence must be accumulated exactly once, by the owner of the vertex
    inc_seen = np.zeros((nel, 3), dtype=np.int64)
    for h in hdr:
        vtx_off, n_owned, n_local, elem_off, n_elem, color_off, n_colors = (int(x) for x in h[:7])
        assert 0 < n_owned <= max_owned and n_owned <= n_local <= 1024 and n_colors <= 64
        loc = vtx[vtx_off: vtx_off + n_local]
        assert len(np.unique(loc)) == n_local
        assert np.all(np.diff(loc[:n_owned]) > 0) and np.all(np.diff(loc[n_owned:]) > 0)   # sorted runs → coalesced staging
        owned_count[loc[:n_owned]] += 1
        # --- element-centric (coloured) view
        cp = col[color_off: color_off + n_colors + 1]
        assert cp[0] == 0 and cp[-1] == n_elem and np.all(np.diff(cp) > 0)
        packed = el[elem_off: elem_off + n_elem]
        idx = np.stack([packed & 1023, (packed >> 10) & 1023, packed >> 20], axis=1).astype(np.int64)
        assert idx.max() < n_local
        ge = eid[elem_off: elem_off + n_elem]
        np.testing.assert_array_equal(loc[idx], c[ge])            # local indices decode to the element's ids, in order
        assert len(np.unique(ge)) == n_elem
        for k in range(n_colors):                                  # no two triangles of a colour share a vertex
            used = idx[cp[k]: cp[k + 1]].ravel()
            assert len(np.unique(used)) == used.size
        own_mask = idx < n_owned
        assert np.all(own_mask.any(axis=1))                        # every listed triangle touches an owned vertex
        r, k = np.nonzero(own_mask)
        inc_seen[ge[r], k] += 1
        # --- vertex-centric view: sliced ELL of 16-bit ring STRIPS.  Column i lists ring vertices; consecutive
        #     entries (without the START bit on the second) span one triangle with the owner.  The multiset of
        #     triangles decoded must be exactly the owner's incident elements.
        slice_off, n_slices, ent_off = int(h[H_SLICE_OFF]), int(h[H_N_SLICES]), int(h[H_ENT_OFF])
        assert n_slices == (n_owned + 31) // 32 and ent_off % 32 == 0 and vtx_off % 4 == 0 and slice_off % 4 == 0
        sp = P["sl"][slice_off: slice_off + n_slices + 1]
        assert sp[0] == 0 and np.all(np.diff(sp) % 32 == 0) and sp[-1] == int(h[H_N_ENT])
        for i in range(n_owned):
            u = int(loc[i])
            s, lane = divmod(i, 32)
            column = P["ent"][ent_off + sp[s] + lane: ent_off + sp[s + 1]: 32].astype(np.int64)
            valid = column != ENT_NONE
            nvalid = int(valid.sum())
            assert np.all(valid[:nvalid]) and not valid[nvalid:].any()       # padding only at the end
            got = []
            for k2 in range(1, nvalid):
                if column[k2] & ENT_START:
                    continue
                a, b2 = int(loc[column[k2 - 1] & 1023]), int(loc[column[k2] & 1023])
                got.append(tuple(sorted((u, a, b2))))
            assert nvalid == 0 or (column[0] & ENT_START)
            want = [tuple(int(x) for x in c[e]) for e in tidx[tptr[u]: tptr[u + 1]]]
            assert sorted(got) == sorted(want)
            if len(want) == 6 or len(want) == 5:
                assert nvalid == len(want) + 1                              # a closed fan is one strip
    assert np.all(owned_count == 1)
    assert np.all(inc_seen == 1)
    assert P["maxcol"] <= 64


def test_patch_plan_handles_isolated_vertices(L):
    from morpho_b200 import meshgen as mg
    v, c = mg.icosphere(6)
    v = np.concatenate([v, [[3.0, 3.0, 3.0], [-2.0, 0.1, 0.4]]])   # two vertices without elements
    P = _plan(L, np.ascontiguousarray(v), c)
    owned = np.concatenate([P["vtx"][h[0]: h[0] + h[1]] for h in P["hdr"]])
    assert sorted(owned) == list(range(v.shape[0]))
