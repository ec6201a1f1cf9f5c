"""GPU, world_size 2 (needs two GPUs on the box; skipped otherwise): element-partitioned Area / VolumeEnclosed
gradients with the shared-vertex rows completed over NVLink peer memory (mcu_halo.cu), against the CPU oracle
on the undivided mesh."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, f, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import morpho_b200 as mb
        from morpho_b200 import _lib, meshgen as mg
        from morpho_b200.partition import PeerExchange, build_rank_mesh
        from oracle.oracle import FunctionalSpec, Oracle
        mb.init(rank)
        L = mb.lib()
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        _lib.check(L.mcu_set_stream(stream.cuda_stream))
        rm = build_rank_mesh(f, rank, world, dist=dist, torch=torch)
        mesh = mb.Mesh(rm.vert, faces=rm.conn)
        grads = [torch.empty(rm.vert.size, dtype=torch.float64, device="cuda") for _ in range(2)]
        ex = PeerExchange(rm, torch, dist, ngrad=2)
        v, c = mg.icosphere(f)
        v = mg.perturb_radially(v)
        O = Oracle()
        errs = []
        mb.set_option("gradient_path", _lib.MCU_PATH_PATCH)
        for rep in range(8):                           # several exchanges: alternating receive buffers, epochs
            mb.set_option("halo_merge", 0 if rep == 1 else 1)   # rep 1: one exchange launch per gradient; else both gradients in one
            if rep == 3:
                ex.attach(mesh)                        # fused mode: the gradient kernel pushes the shared rows itself
            for slot, (g, cls) in enumerate(zip(grads, (mb.Area, mb.VolumeEnclosed))):
                d = cls()._desc()
                if rep >= 3:
                    ex.bind(slot)
                _lib.check(L.mcu_eval_device(mesh._h, C.byref(d), None, _lib.MCU_GRADIENT, g.data_ptr()))
                if rep >= 4:                           # overlapped: this gradient's exchange runs on the second stream
                    ex.exchange_async(slot, g.data_ptr())
            if rep >= 4:
                ex.join()
            else:
                ex(*grads)
            assert ex.status() == 0
            for g, name in zip(grads, ("Area", "VolumeEnclosed")):
                want = O.gradient(v, {2: c}, FunctionalSpec(name))[rm.gids]
                got = g.cpu().numpy().reshape(-1, 3)
                errs.append(np.linalg.norm(got - want, axis=1).max() / np.linalg.norm(want, axis=1).max())
        # a gradient that does NOT come from the pushing patch kernel while the exchange is attached (AUTO on a mesh
        # below patch_min_elements takes the gather kernel; so do selections and other functionals): the exchange has to
        # notice that nothing was pushed and push the rows itself — synchronously and on the second stream
        mb.set_option("gradient_path", _lib.MCU_PATH_AUTO)
        mb.set_option("patch_min_elements", 10 ** 9)
        for mode in ("serial", "overlap"):
            for slot, (g, cls) in enumerate(zip(grads, (mb.Area, mb.VolumeEnclosed))):
                d = cls()._desc()
                ex.bind(slot)
                g.zero_()
                _lib.check(L.mcu_eval_device(mesh._h, C.byref(d), None, _lib.MCU_GRADIENT, g.data_ptr()))
                if mode == "overlap":
                    ex.exchange_async(slot, g.data_ptr())
            if mode == "overlap":
                ex.join()
            else:
                ex(*grads)
            assert ex.status() == 0
            for g, name in zip(grads, ("Area", "VolumeEnclosed")):
                want = O.gradient(v, {2: c}, FunctionalSpec(name))[rm.gids]
                got = g.cpu().numpy().reshape(-1, 3)
                errs.append(np.linalg.norm(got - want, axis=1).max() / np.linalg.norm(want, axis=1).max())
        mb.set_option("patch_min_elements", 4096)
        # energies: per-rank totals (patch kernel) summed in rank order
        from morpho_b200.partition import reduce_total
        for cls, name in ((mb.Area, "Area"), (mb.VolumeEnclosed, "VolumeEnclosed")):
            tot = reduce_total(cls().total(mesh), torch, dist)
            ref = O.total(v, {2: c}, FunctionalSpec(name))
            errs.append(abs(tot - ref) / abs(ref))
        # shared rows bitwise identical on both ranks
        sh = torch.zeros((rm.n_slots, 3), dtype=torch.float64, device="cuda")
        sh[torch.as_tensor(rm.shared_slot, device="cuda")] = grads[0].view(-1, 3)[torch.as_tensor(rm.shared_local, device="cuda")]
        allsh = [torch.zeros_like(sh) for _ in range(world)]
        dist.all_gather(allsh, sh)
        same = all(bool(torch.equal(allsh[r], sh)) for r in range(world)) if world == 2 else True
        ex.close()
        q.put((rank, max(errs), rm.n_shared, same))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_partitioned_gradient_over_peer_memory():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 160, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, nshared, same in res:
        assert err <= 1e-12, (rank, err)
        assert nshared > 0 and same


def _worker_tets(rank, world, port, n, q):
    """Config 5 in miniature: a tetrahedral cube partitioned by element ranges; Volume gradient per rank, shared rows
    completed by the (non-fused) peer-memory exchange; field gradient of GradSq likewise (5 doubles per vertex)."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import morpho_b200 as mb
        from morpho_b200 import _lib, meshgen as mg
        from morpho_b200.partition import PeerExchange, partition_arrays, reduce_total
        from oracle.oracle import FunctionalSpec, Oracle
        mb.init(rank)
        L = mb.lib()
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        _lib.check(L.mcu_set_stream(stream.cuda_stream))
        v, t = mg.tet_cube(n)
        rng = np.random.Generator(np.random.PCG64(5))
        v = v + 0.1 / n * rng.uniform(-1, 1, size=v.shape)
        qf = rng.standard_normal((v.shape[0], 5))
        rm = partition_arrays(v, t, world)[rank]
        mesh = mb.Mesh(rm.vert, volumes=rm.conn)
        O = Oracle()
        errs = []
        # vertex gradient of Volume: 3 doubles per shared vertex
        ex3 = PeerExchange(rm, torch, dist, ngrad=1)
        g = torch.from_numpy(mb.Volume().gradient(mesh).ravel().copy()).cuda()
        ex3(g)
        assert ex3.status() == 0
        want = O.gradient(v, {3: t}, FunctionalSpec("Volume"))[rm.gids]
        got = g.cpu().numpy().reshape(-1, 3)
        errs.append(np.linalg.norm(got - want, axis=1).max() / np.linalg.norm(want, axis=1).max())
        tot = reduce_total(mb.Volume().total(mesh), torch, dist)
        errs.append(abs(tot - O.total(v, {3: t}, FunctionalSpec("Volume"))) / tot)
        ex3.close()
        # field gradient of GradSq: rows of 5 doubles go through the same exchange (dim = 5 columns per vertex)
        fld = mb.Field(mesh, qf[rm.gids])
        fg = torch.from_numpy(mb.GradSq(fld).fieldgradient(fld, mesh).ravel().copy()).cuda()
        ex5 = PeerExchange(rm, torch, dist, ngrad=1, dim=5)
        ex5(fg)
        assert ex5.status() == 0
        wantf = O.fieldgradient(v, {3: t}, FunctionalSpec("GradSq", fields=[qf]), 0)[rm.gids]
        gotf = fg.cpu().numpy().reshape(-1, 5)
        errs.append(np.linalg.norm(gotf - wantf, axis=1).max() / np.linalg.norm(wantf, axis=1).max())
        ex5.close()
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_partitioned_tet_mesh_over_peer_memory():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_tets, args=(r, world, port, 10, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert errs[0] <= 1e-12 and errs[1] <= 1e-12, (rank, errs)     # analytic gradient and total
        assert errs[2] <= 2e-9, (rank, errs)                             # finite-difference field gradient


def _worker_pairwise(rank, world, port, n, q):
    """Config 3 in miniature on two GPUs: row blocks of the all-pairs interaction, no exchange until the final
    concatenation; against the oracle's i<j double loop on the whole point set."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import morpho_b200 as mb
        from morpho_b200 import _lib, meshgen as mg
        from morpho_b200.partition import pairwise_partitioned
        from oracle.oracle import Oracle
        mb.init(rank)
        x = mg.sphere_points(n, seed=77)
        O = Oracle()
        e_ref, g_ref = O.pairwise_coulomb(x, "integrand"), O.pairwise_coulomb(x, "gradient")
        e = pairwise_partitioned(x, _lib.MCU_INTEGRAND, torch, dist)
        g = pairwise_partitioned(x, _lib.MCU_GRADIENT, torch, dist)
        tot = pairwise_partitioned(x, _lib.MCU_TOTAL, torch, dist)
        errs = [np.max(np.abs(e - e_ref)) / np.max(np.abs(e_ref)),
                np.max(np.linalg.norm(g - g_ref, axis=1)) / np.max(np.linalg.norm(g_ref, axis=1)),
                abs(tot - e_ref.sum()) / abs(e_ref.sum())]
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_partitioned_pairwise_row_blocks():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pairwise, args=(r, world, port, 3001, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) <= 1e-12, (rank, errs)
