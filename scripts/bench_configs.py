#!/usr/bin/env python3
"""BASELINE.json configs 3, 4 and 5 measured on one B200 (kernel times with CUDA events on the library stream, data
resident), each with the roofline that bounds it.  bench.py appends the result as `"configs"` to its JSON line at N=1;
run standalone for a table:   python scripts/bench_configs.py [3 4 5]

Algorithmic bytes (SURVEY.md 8d / BASELINE.md 4), FP64 values, int32 ids:
  triangles (V = F/2):  gradient 36 B/element; total 24; integrand 32; a psize-p vertex field adds 4p B read per element
                        (and 4p written for a field gradient)
  tetrahedra (V = T/6): Volume gradient 16 + 4 + 4 = 24 B; total 16 + 4 = 20; psize-p field adds 8p/6 B read
                        (+ 8p/6 written for a field gradient)
  pairwise: 20 FP64 flop per ordered pair, N^2 pairs; peak = DFMA micro-benchmark on this device.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(L, launch, reps, warm=2):
    ms = C.c_float()
    for _ in range(warm):
        launch()
    L.mcu_sync()
    L.mcu_timer_start()
    for _ in range(reps):
        launch()
    L.mcu_timer_stop(C.byref(ms))
    return ms.value / reps


def config3(mb, torch, hbm_peak, n=200_000, reps=3):
    """examples/thomson scaled to 200k points: all-pairs Coulomb energy / gradient, FP64 compute-bound."""
    from morpho_b200 import _lib, meshgen as mg
    L = mb.lib()
    x = torch.from_numpy(mg.sphere_points(n)).cuda()
    out = torch.empty(n * 3, dtype=torch.float64, device="cuda")
    peak = C.c_double()
    _lib.check(L.mcu_debug_dfma_peak(C.byref(peak)))
    res = {"workload": f"{n} points on S^2 (normalised Gaussian, seed 12345), Coulomb f=1/r, no cut-off",
           "fp64_peak_tflops_dfma": peak.value, "flop_per_pair": 20, "kernels": {}}
    for mode, name in ((_lib.MCU_GRADIENT, "gradient"), (_lib.MCU_INTEGRAND, "integrand")):
        ms = _time(L, lambda: _lib.check(L.mcu_pairwise_device(n, 3, x.data_ptr(), 0, None, 0.0, mode, out.data_ptr())), reps, 1)
        tf = 20.0 * n * n / (ms * 1e-3) / 1e12
        res["kernels"][name] = {"kernel": "k_pairwise<Coulomb>", "ms": ms, "gpairs_per_s": n * float(n) / (ms * 1e-3) / 1e9,
                                "tflops": tf, "bound": "fp64", "frac": tf / peak.value}
    res["value"] = res["kernels"]["gradient"]["gpairs_per_s"]
    res["unit"] = "G ordered pairs/s (gradient)"
    res["frac"] = res["kernels"]["gradient"]["frac"]
    return res


def config4(mb, torch, hbm_peak, n=816, reps=5):
    """examples/tactoid at ~4M triangles: Nematic + Length on the boundary + Area + NormSq with a director field."""
    from morpho_b200 import _lib, meshgen as mg
    L = mb.lib()
    v, faces, bnd = mg.disk(n, dim=3)
    rng = np.random.Generator(np.random.PCG64(12345))
    th = 0.3 * rng.uniform(-1, 1, size=v.shape[0]) + 0.5 * v[:, 0]
    nn = np.stack([np.cos(th), np.sin(th), 0.1 * rng.uniform(-1, 1, size=v.shape[0])], axis=1)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    m = mb.Mesh(v, edges=bnd, faces=faces)
    fld = mb.Field(m, nn)
    nel = faces.shape[0]
    out = torch.empty(max(v.size, nel), dtype=torch.float64, device="cuda")
    res = {"workload": f"unit disk, {nel} triangles, {v.shape[0]} vertices, director field (3 components, seeded rotation noise)",
           "kernels": {}}
    nem = mb.Nematic(fld, ksplay=1.0, ktwist=0.6, kbend=1.4)
    # bytes per element: ids 12 + vertices 12 + field 12 (+ output)
    rows = ((nem, "Nematic.total", _lib.MCU_TOTAL, 36.0), (nem, "Nematic.integrand", _lib.MCU_INTEGRAND, 44.0),
            (nem, "Nematic.gradient", _lib.MCU_GRADIENT, 48.0), (nem, "Nematic.fieldgradient", _lib.MCU_FIELDGRADIENT, 48.0),
            (mb.Area(), "Area.gradient", _lib.MCU_GRADIENT, 36.0), (mb.GradSq(fld), "GradSq.fieldgradient", _lib.MCU_FIELDGRADIENT, 48.0))
    for f, name, mode, bpe in rows:
        d = f._desc(0)
        ms = _time(L, lambda: _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr())), reps)
        gbs = bpe * nel / (ms * 1e-3) / 1e9
        res["kernels"][name] = {"ms": ms, "gelem_per_s": nel / (ms * 1e-3) / 1e9, "bytes_per_element": bpe, "achieved_gbs": gbs,
                                "bound": "hbm", "frac": gbs / hbm_peak}
    res["value"] = res["kernels"]["Nematic.fieldgradient"]["gelem_per_s"]
    res["unit"] = "G element-evals/s (Nematic.fieldgradient)"
    res["frac"] = res["kernels"]["Nematic.fieldgradient"]["frac"]
    return res


def config5(mb, torch, hbm_peak, n=204, reps=3):
    """tetrahedral cube, 6 n^3 tets: Volume total/gradient, GradSq on a 5-component field, Nematic on a director."""
    from morpho_b200 import _lib, meshgen as mg
    L = mb.lib()
    t0 = time.time()
    v, tets = mg.tet_cube(n)
    rng = np.random.Generator(np.random.PCG64(12345))
    inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
    v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
    m = mb.Mesh(v, volumes=tets)
    q = np.stack([v[:, 0], 2 * v[:, 1], -v[:, 2], v[:, 0] + v[:, 1], 0 * v[:, 0] + 3.0], axis=1) + 0.01 * rng.standard_normal((v.shape[0], 5))
    nn = np.stack([np.cos(v[:, 0] + 0.3 * v[:, 1]), np.sin(v[:, 0] + 0.3 * v[:, 1]), 0.2 * np.sin(4 * v[:, 2])], axis=1)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    fq, fn = mb.Field(m, q), mb.Field(m, nn)
    nel = tets.shape[0]
    out = torch.empty(max(v.shape[0] * 5, nel), dtype=torch.float64, device="cuda")
    res = {"workload": f"unit cube, {nel} tetrahedra, {v.shape[0]} vertices ({n}^3 cells x 6), 5-component and director fields; "
                       f"built + mirrored in {time.time() - t0:.1f} s", "kernels": {}}
    rows = ((mb.Volume(), "Volume.total", _lib.MCU_TOTAL, 20.0), (mb.Volume(), "Volume.gradient", _lib.MCU_GRADIENT, 24.0),
            (mb.GradSq(fq), "GradSq(5).total", _lib.MCU_TOTAL, 20.0 + 40.0 / 6), (mb.GradSq(fq), "GradSq(5).fieldgradient", _lib.MCU_FIELDGRADIENT, 20.0 + 80.0 / 6),
            (mb.Nematic(fn, ksplay=1.0, ktwist=0.8, kbend=1.2), "Nematic.fieldgradient", _lib.MCU_FIELDGRADIENT, 20.0 + 48.0 / 6))
    for f, name, mode, bpe in rows:
        d = f._desc(0)
        ms = _time(L, lambda: _lib.check(L.mcu_eval_device(m._h, C.byref(d), None, mode, out.data_ptr())), reps, 1)
        gbs = bpe * nel / (ms * 1e-3) / 1e9
        res["kernels"][name] = {"ms": ms, "gelem_per_s": nel / (ms * 1e-3) / 1e9, "bytes_per_element": bpe, "achieved_gbs": gbs,
                                "bound": "hbm", "frac": gbs / hbm_peak}
    res["value"] = res["kernels"]["Volume.gradient"]["gelem_per_s"]
    res["unit"] = "G element-evals/s (Volume.gradient)"
    res["frac"] = res["kernels"]["Volume.gradient"]["frac"]
    return res


def config5_partitioned(mb, torch, dist, hbm_peak, n=204, reps=3, log=print):
    """BASELINE config 5 as north_star states it: the ~50M-tetrahedron cube ELEMENT-PARTITIONED over the ranks
    (contiguous element ranges, functional_binbounds functional.c:863-879): every rank evaluates Volume.gradient,
    GradSq(5).fieldgradient and Nematic.fieldgradient on its range and the rows of vertices shared between ranks are
    completed over NVLink peer memory (mcu_halo.cu; 3, 5 and 3 doubles per vertex) and summed in rank order.
    Strong scaling (total work fixed).  Before timing, every rank's completed rows are compared with the same
    evaluation of the UNDIVIDED mesh on rank 0 (the single-GPU path the parity tests pin to the oracle)."""
    from morpho_b200 import _lib, meshgen as mg
    from morpho_b200.partition import PeerExchange, partition_arrays
    L = mb.lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    t0 = time.time()
    v, tets = mg.tet_cube(n)
    rng = np.random.Generator(np.random.PCG64(12345))
    inner = np.all((v > 1e-9) & (v < 1 - 1e-9), axis=1)
    v = v + (0.2 / n) * rng.uniform(-1, 1, size=v.shape) * inner[:, None]
    q = np.stack([v[:, 0], 2 * v[:, 1], -v[:, 2], v[:, 0] + v[:, 1], 0 * v[:, 0] + 3.0], axis=1) + 0.01 * rng.standard_normal((v.shape[0], 5))
    nn = np.stack([np.cos(v[:, 0] + 0.3 * v[:, 1]), np.sin(v[:, 0] + 0.3 * v[:, 1]), 0.2 * np.sin(4 * v[:, 2])], axis=1)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    rm = partition_arrays(v, tets, world)[rank]
    nel = tets.shape[0]
    t_part = time.time() - t0

    def functionals(mesh, fq, fn):
        return ((mb.Volume(), "Volume.gradient", _lib.MCU_GRADIENT, 3, 24.0),
                (mb.GradSq(fq), "GradSq(5).fieldgradient", _lib.MCU_FIELDGRADIENT, 5, 20.0 + 80.0 / 6),
                (mb.Nematic(fn, ksplay=1.0, ktwist=0.8, kbend=1.2), "Nematic.fieldgradient", _lib.MCU_FIELDGRADIENT, 3, 20.0 + 48.0 / 6))

    # the undivided mesh on rank 0, results broadcast to everybody
    full = {}
    if rank == 0:
        mfull = mb.Mesh(v, volumes=tets)
        fqf, fnf = mb.Field(mfull, q), mb.Field(mfull, nn)
        for f, name, mode, dim, _ in functionals(mfull, fqf, fnf):
            out = torch.empty(v.shape[0] * dim, dtype=torch.float64, device="cuda")
            d = f._desc(0)
            _lib.check(L.mcu_eval_device(mfull._h, C.byref(d), None, mode, out.data_ptr()))
            L.mcu_sync()
            full[name] = out
        del fqf, fnf, mfull
    else:
        for name, dim in (("Volume.gradient", 3), ("GradSq(5).fieldgradient", 5), ("Nematic.fieldgradient", 3)):
            full[name] = torch.empty(v.shape[0] * dim, dtype=torch.float64, device="cuda")
    for name in ("Volume.gradient", "GradSq(5).fieldgradient", "Nematic.fieldgradient"):
        dist.broadcast(full[name], 0)

    mesh = mb.Mesh(rm.vert, volumes=rm.conn)
    fq, fn = mb.Field(mesh, q[rm.gids]), mb.Field(mesh, nn[rm.gids])
    gid_t = torch.as_tensor(rm.gids, dtype=torch.int64, device="cuda")
    res = {"workload": f"unit cube, {nel} tetrahedra, {v.shape[0]} vertices ({n}^3 cells x 6) partitioned by contiguous element ranges over {world} GPUs "
                       f"({rm.conn.shape[0]} tetrahedra, {rm.vert.shape[0]} vertices, {rm.n_shared} shared on rank 0); strong scaling",
           "kernels": {}, "parity_vs_undivided_mesh": {}}
    stream = torch.cuda.current_stream()
    for f, name, mode, dim, bpe in functionals(mesh, fq, fn):
        out = torch.empty(rm.vert.shape[0] * dim, dtype=torch.float64, device="cuda")
        ex = PeerExchange(rm, torch, dist, ngrad=1, dim=dim)
        d = f._desc(0)

        def step(evs=None):
            if evs: evs[0].record(stream)
            _lib.check(L.mcu_eval_device(mesh._h, C.byref(d), None, mode, out.data_ptr()))
            if evs: evs[1].record(stream)
            ex(out)
            if evs: evs[2].record(stream)

        step()
        L.mcu_sync()
        if ex.status() != 0:
            raise RuntimeError("halo exchange: a peer did not arrive in time")
        want = full[name].view(-1, dim).index_select(0, gid_t)
        got = out.view(-1, dim)
        err = torch.linalg.norm(got - want, dim=1).max() / torch.linalg.norm(full[name].view(-1, dim), dim=1).max()
        err_t = err.reshape(1).clone()
        dist.all_reduce(err_t, op=dist.ReduceOp.MAX)
        res["parity_vs_undivided_mesh"][name] = float(err_t.item())
        dist.barrier(); torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
        for k in range(reps):
            step(evs[k])
        dist.barrier(); torch.cuda.synchronize()
        ts = torch.tensor([float(np.mean([e[0].elapsed_time(e[2]) for e in evs])), float(np.mean([e[0].elapsed_time(e[1]) for e in evs])),
                           float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms, ms_k, ms_x = [float(x) for x in ts.cpu()]
        gbs = bpe * nel / (ms * 1e-3) / 1e9
        res["kernels"][name] = {"ms": ms, "kernel_ms": ms_k, "shared_row_exchange_ms": ms_x, "gelem_per_s": nel / (ms * 1e-3) / 1e9,
                                "bytes_per_element": bpe, "achieved_gbs_all_gpus": gbs, "frac_of_n_gpu_roofline": gbs / (hbm_peak * world)}
        ex.close()
    res["partition_s"] = t_part
    res["value"] = res["kernels"]["Volume.gradient"]["gelem_per_s"]
    res["unit"] = f"G element-evals/s (Volume.gradient, {world} GPUs, whole job)"
    res["frac"] = res["kernels"]["Volume.gradient"]["frac_of_n_gpu_roofline"]
    worst = max(res["parity_vs_undivided_mesh"].values())
    if worst > 1e-12:
        raise RuntimeError(f"partitioned config 5 disagrees with the undivided mesh: {res['parity_vs_undivided_mesh']}")
    return res


def run_all(mb, torch, hbm_peak, which=("3", "4", "5"), log=print):
    out = {}
    for key, fn in (("3", config3), ("4", config4), ("5", config5)):
        if key not in which:
            continue
        t0 = time.time()
        try:
            out[key] = fn(mb, torch, hbm_peak)
            out[key]["wall_s"] = time.time() - t0
        except Exception as e:                       # a failed side measurement must not take the headline line down
            out[key] = {"error": f"{type(e).__name__}: {e}"}
        log(f"[configs] {key}: {out[key].get('value')} {out[key].get('unit', '')} frac {out[key].get('frac')} ({time.time() - t0:.1f}s)")
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    import torch
    import morpho_b200 as mb
    mb.init(0)
    peak = 6549.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    which = tuple(sys.argv[1:]) or ("3", "4", "5")
    r = run_all(mb, torch, peak, which)
    print(json.dumps(r, indent=1))
