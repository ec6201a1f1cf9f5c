"""CPU: host-side logic of the multi-GPU path — element partition (functional_binbounds scheme), compact
local numbering, the shared-vertex exchange (world_size 2 and 3 over gloo), checked with the CPU oracle:
every rank's completed gradient rows must equal the global gradient on the vertices it touches."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morpho_b200 import meshgen as mg
from morpho_b200.partition import binbounds, build_rank_mesh, partition_arrays


def test_binbounds_matches_oracle(oracle):
    for nel, nb in ((10, 3), (20 * 707 * 707, 8), (7, 7), (5, 8), (0, 3)):
        np.testing.assert_array_equal(binbounds(nel, nb), oracle.binbounds(nel, nb))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_arrays_indexing_is_exact(world):
    v, c = mg.icosphere(9)
    v = mg.perturb_radially(v)
    parts = partition_arrays(v, c, world)
    seen = np.zeros(c.shape[0], dtype=int)
    touch = np.zeros(v.shape[0], dtype=int)
    for p in parts:
        lo, hi = p.elem_range
        seen[lo:hi] += 1
        np.testing.assert_array_equal(p.gids[p.conn], c[lo:hi])          # local → global round trip, bit exact
        np.testing.assert_array_equal(p.vert, v[p.gids])
        assert np.all(np.diff(p.gids) > 0)
        touch[p.gids] += 1
    assert np.all(seen == 1)
    shared_global = np.nonzero(touch >= 2)[0]
    for p in parts:
        np.testing.assert_array_equal(np.sort(p.gids[p.shared_local]), np.intersect1d(p.gids, shared_global))
        assert p.n_slots == len(shared_global)
        np.testing.assert_array_equal(shared_global[p.shared_slot], p.gids[p.shared_local])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, f, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import FunctionalSpec, Oracle
        O = Oracle()
        rm = build_rank_mesh(f, rank, world, dist=dist, torch=torch)
        v, c = mg.icosphere(f)
        v = mg.perturb_radially(v)
        # face-by-face generation reproduces the global generator exactly
        lo, hi = rm.elem_range
        assert np.array_equal(rm.gids[rm.conn], c[lo:hi]) and np.array_equal(rm.vert, v[rm.gids])
        grads, want = [], []
        for name in ("Area", "VolumeEnclosed"):
            spec = FunctionalSpec(name)
            grads.append(torch.from_numpy(O.gradient(rm.vert, {2: rm.conn}, spec).ravel().copy()))
            want.append(O.gradient(v, {2: c}, spec)[rm.gids])
        from morpho_b200.partition import reduce_total
        for name in ("Area", "VolumeEnclosed"):
            spec = FunctionalSpec(name)
            tot = reduce_total(O.total(rm.vert, {2: rm.conn}, spec), torch, dist)
            ref = O.total(v, {2: c}, spec)
            assert abs(tot - ref) <= 1e-13 * abs(ref), (name, tot, ref)
        rm.make_exchange(torch, dist)(*grads)
        errs = []
        for g, w in zip(grads, want):
            g = g.numpy().reshape(-1, 3)
            errs.append(np.linalg.norm(g - w, axis=1).max() / np.linalg.norm(w, axis=1).max())
        # rows of shared vertices must be bitwise identical on every rank that holds them
        sh = torch.zeros((rm.n_slots, 3), dtype=torch.float64)
        sh[torch.as_tensor(rm.shared_slot)] = grads[0].view(-1, 3)[torch.as_tensor(rm.shared_local)]
        allsh = [torch.zeros_like(sh) for _ in range(world)]
        dist.all_gather(allsh, sh)
        mask = [torch.zeros(rm.n_slots, dtype=torch.bool) for _ in range(world)]
        mine = torch.zeros(rm.n_slots, dtype=torch.bool)
        mine[torch.as_tensor(rm.shared_slot)] = True
        dist.all_gather(mask, mine)
        same = True
        for r in range(world):
            both = mask[r] & mine
            same &= bool(torch.equal(allsh[r][both], sh[both]))
        q.put((rank, max(errs), rm.n_shared, same))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_gradient_with_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nshared, same in res:
        assert err <= 1e-13, (rank, err)
        assert nshared > 0 and same


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_exchange_plan_emulated(oracle, world):
    """Index logic of the NVLink peer-memory exchange (mcu_halo.cu), emulated with numpy in one process:
    push every rank's partial rows into the peers' receive buffers as the plan says, combine in ascending
    rank order, and compare with the gradient of the undivided mesh.  Shared rows must come out bitwise
    identical on all ranks that hold them."""
    from oracle.oracle import FunctionalSpec
    from morpho_b200.partition import halo_plan
    v, c = mg.icosphere(10)
    v = mg.perturb_radially(v)
    parts = partition_arrays(v, c, world)
    spec = FunctionalSpec("Area")
    want = oracle.gradient(v, {2: c}, spec)
    plans = [halo_plan(p) for p in parts]
    max_m = plans[0]["max_m"]
    assert all(pl["max_m"] == max_m for pl in plans)
    partial = [oracle.gradient(p.vert, {2: p.conn}, spec) for p in parts]
    recv = [np.full((world, max(max_m, 1), 3), np.nan) for _ in range(world)]
    for r, (p, pl) in enumerate(zip(parts, plans)):
        assert np.all(np.diff(p.gids[pl["shared_local"]]) > 0)
        for q in range(world):
            pos = pl["peer_pos"][pl["peer_ptr"][q]: pl["peer_ptr"][q + 1]]
            if q == r:
                assert len(pos) == 0
                continue
            # both sides of the pair list the same vertices in the same order
            posq = plans[q]["peer_pos"][plans[q]["peer_ptr"][r]: plans[q]["peer_ptr"][r + 1]]
            np.testing.assert_array_equal(p.gids[pl["shared_local"][pos]], parts[q].gids[plans[q]["shared_local"][posq]])
            recv[q][r, : len(pos)] = partial[r][pl["shared_local"][pos]]
    done = []
    for r, (p, pl) in enumerate(zip(parts, plans)):
        g = partial[r].copy()
        for i, vloc in enumerate(pl["shared_local"]):
            s = np.zeros(3)
            ranks = pl["contrib_rank"][pl["contrib_ptr"][i]: pl["contrib_ptr"][i + 1]]
            ks = pl["contrib_k"][pl["contrib_ptr"][i]: pl["contrib_ptr"][i + 1]]
            assert np.all(np.diff(ranks) > 0) and r in ranks and len(ranks) >= 2
            for q, k in zip(ranks, ks):
                s = s + (partial[r][vloc] if k < 0 else recv[r][q, k])
            g[vloc] = s
        assert not np.isnan(g).any()
        err = np.linalg.norm(g - want[p.gids], axis=1).max() / np.linalg.norm(want, axis=1).max()
        assert err <= 1e-13
        done.append(g)
    glob = {}
    for r, (p, pl) in enumerate(zip(parts, plans)):
        for vloc in pl["shared_local"]:
            key = int(p.gids[vloc])
            if key in glob:
                assert np.array_equal(glob[key], done[r][vloc])      # bitwise identical on every sharing rank
            glob[key] = done[r][vloc]
