#!/usr/bin/env python3
"""bench.py — headline benchmark of the functional hot path (BASELINE.json):

    Area + VolumeEnclosed gradient throughput (M element-evals/s) on a ~10M-triangle sphere
    (geodesic icosphere f=707: 9 996 980 triangles, 4 998 492 vertices, radially perturbed,
    PCG64 seed 12345), on N B200s of one node, with the HBM roofline of the dominant kernel
    and the reference's CPU path timed on the same box.

A "step" is one pass of the hot path over the mesh: `Area.gradient(m)` followed by
`VolumeEnclosed.gradient(m)`  =  2 x n_triangles element-evals.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--freq F]

N > 1 (launched by torchrun, one rank per GPU): weak scaling — the global sphere has N x 10M
triangles (f = round(707 sqrt N)) and is partitioned by contiguous element ranges across the
ranks (the reference's own scheme, functional_binbounds functional.c:863-879); vertices shared
between ranks get their partial gradient rows pushed into the peers' receive buffers over NVLink peer
memory (CUDA IPC, morpho_b200/csrc/mcu_halo.cu) and summed in fixed rank order; NCCL only carries the
set-up and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "area_volume_gradient_throughput"
UNIT = "Melem/s"
BYTES_PER_ELEM_GRADIENT = 36.0   # 12 B ids + 12 B vertex (each read once) + 12 B gradient row


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def frequency_for(ngpu: int, base: int) -> int:
    return int(round(base * np.sqrt(ngpu)))


def workload_string(f: int, world: int) -> str:
    """The same text in both arms (the driver compares the arms' `config.workload`)."""
    nel, nv = 20 * f * f, 10 * f * f + 2
    per = -(-nel // world)
    return (f"icosphere f={f} ({nel} triangles, {nv} vertices global; {per} triangles per GPU on {world} GPU(s)), radial perturbation 1%, "
            "seed 12345; step = Area.gradient + VolumeEnclosed.gradient (2 element-evals per triangle)")


INPUTS = "larger than L2 (vertex matrix 120 MB + ring tables 101 MB + gradient 120 MB per functional, per GPU)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi absent
            log("clock sampler unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            t = [x.strip() for x in l.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); smax.append(float(t[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------

def reference_steps(vert, conn, steps, warmup, threads):
    """Time `steps` reference steps (Area.gradient + VolumeEnclosed.gradient builtins of the
    unmodified libmorpho via oracle/_ref) on the host; returns seconds per step list."""
    from oracle.oracle import FunctionalSpec, Reference
    ref = Reference(threads)
    rm = ref.mesh(vert, {2: conn})
    fa, fv = rm.functional(FunctionalSpec("Area")), rm.functional(FunctionalSpec("VolumeEnclosed"))
    out = []
    for i in range(warmup + steps):
        t = rm.time(fa, "gradient", 1) + rm.time(fv, "gradient", 1)
        if t < 0:
            raise RuntimeError("reference gradient failed: " + ref.err())
        if i >= warmup:
            out.append(t)
    rm.free()
    return out


def run_reference(args):
    """`--impl reference`: the UNMODIFIED reference (oracle/_ref, built from /root/reference/src) through its own
    builtins on the host cores.  Each step is a bounded sample of the workload: the first M triangles of the
    mesh (with the vertices they touch), M chosen from a calibration step so that warm-up + timed steps take
    about REFERENCE_BUDGET_S seconds; M = all triangles when that fits."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from morpho_b200.partition import build_rank_mesh
    world = max(int(args.gpus), 1)
    f = frequency_for(world, args.freq) if args.scaling != "strong" else args.freq
    t0 = time.time()
    part = build_rank_mesh(f, 0, 1) if world == 1 else _rank0_range(f, world)
    vert, conn = part.vert, part.conn
    nel = conn.shape[0]
    log(f"[reference] mesh f={f}: rank 0's {nel} of {part.nel_global} triangles built in {time.time() - t0:.1f}s")
    ncores = os.cpu_count() or 1
    # calibrate on the whole mesh (up to 10M triangles: ~2 s per step and setting).  "All the host threads it can use": the reference's threaded map does not always
    # win (per-thread full gradient matrices + serial merge, functional.c:1092-1135), so keep the faster setting.
    def submesh(m):
        c = conn[:m]
        gids = np.unique(c)
        return np.ascontiguousarray(vert[gids]), np.ascontiguousarray(np.searchsorted(gids, c).astype(np.int32))
    # (on a sub-mesh large enough to be representative: the threaded map allocates one full-size gradient matrix per
    #  thread and merges them serially, which wins at 1M triangles (26.7 against 14.9 M element-evals/s) and loses at 10M (10.3 against 13.8))
    mcal = min(nel, 10_000_000)
    vcal, ccal = submesh(mcal)
    cal = {}
    for th in (0, ncores):
        cal[th] = float(np.mean(reference_steps(vcal, ccal, 1, 1, th)))
        log(f"[reference] calibration ({mcal} triangles) threads={th}: {cal[th]:.3f} s/step")
    threads = min(cal, key=cal.get)
    budget = float(os.environ.get("REFERENCE_BUDGET_S", "120"))
    per_tri = cal[threads] / mcal
    m = int(min(nel, max(100_000, budget / max(args.steps + args.warmup, 1) / per_tri)))
    vs, cs = (vert, conn) if m == nel else submesh(m)
    ts = reference_steps(vs, cs, args.steps, args.warmup, threads)
    sec = float(np.mean(ts))
    value = 2.0 * m / sec / 1e6
    sample = (f"first {m} of {part.nel_global} triangles per step ({'the whole mesh' if m == part.nel_global else 'bounded sample'}), {args.steps} steps; "
              f"threads={threads} (0 = the reference's serial default path), the faster of serial and {ncores} threads; OPENBLAS_NUM_THREADS=1")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(f, world), "inputs": INPUTS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": max(threads, 1), "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def _rank0_range(f, world):
    """Rank 0's contiguous element range of the N-GPU workload, built without torch.distributed (no shared-vertex lists)."""
    from morpho_b200 import meshgen as mg
    from morpho_b200.partition import RankMesh, binbounds
    topo = mg.IcoTopology(f)
    hi = int(binbounds(topo.nel, world)[1])
    tpf = topo.tri_per_face
    parts = [topo.face_triangles(fi)[: min(hi - fi * tpf, tpf)] for fi in range((hi - 1) // tpf + 1)]
    cg = np.concatenate(parts)
    gids = np.unique(cg)
    vert = topo.positions(gids) * mg.radial_factors(topo.nv, 0.01, 12345)[gids][:, None]
    conn = np.sort(np.searchsorted(gids, cg).astype(np.int32), axis=1)
    return RankMesh(0, world, topo.nel, topo.nv, (0, hi), gids.astype(np.int64), np.ascontiguousarray(vert), np.ascontiguousarray(conn))


def cpu_baseline_sample(vert, conn):
    """Bounded sample of the same workload on the reference's serial path (its default, -w0): the full
    mesh, 2 steps (~10-20 s of CPU work)."""
    try:
        ts = reference_steps(vert, conn, 2, 1, 0)
        sec = float(np.mean(ts))
        return {"value": 2.0 * conn.shape[0] / sec / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "full 10M-triangle mesh, 1 warm-up + 2 timed steps, serial reference path (-w0, OPENBLAS_NUM_THREADS=1)"}
    except Exception as e:
        log("cpu baseline via oracle/_ref failed:", e)
    from oracle.oracle import FunctionalSpec, Oracle
    O = Oracle()
    sub = conn[: 2_000_000]
    t0 = time.time()
    O.gradient(vert, {2: sub}, FunctionalSpec("Area"))
    O.gradient(vert, {2: sub}, FunctionalSpec("VolumeEnclosed"))
    sec = time.time() - t0
    return {"value": 2.0 * sub.shape[0] / sec / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first 2M triangles of the mesh, 1 step, oracle/morpho_oracle.c (scalar C restatement)"}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def parity_check(part, vert, conn, step_device, g_area, g_vol, torch, dist, world, rank, barrier, nsample=10_000):
    """Before anything is timed: one step on the device, then this rank's COMPLETED gradient rows — every row of a
    vertex shared with another rank plus `nsample` seeded interior rows — against the CPU oracle
    (oracle/morpho_oracle.c, the restatement of functional_mapgradient + area/volumeenclosed gradients that
    tests/test_oracle_golden.py pins to the unmodified reference).  The oracle evaluates this rank's element range;
    the rows of shared vertices are completed on the host the way the reference merges its per-thread matrices
    (functional.c:1106-1119): partial rows all-gathered and added in ascending rank order.  Returns the max over ranks
    of the global measure max_v|d|/max_v|g| and of the per-vertex measure of tests/parity.py, and whether all ranks
    that share a vertex hold bitwise identical rows."""
    from oracle.oracle import FunctionalSpec, Oracle
    from tests.parity import contribution_norm_sum, EPS_FLOOR
    O = Oracle()
    step_device()
    barrier()
    dim = vert.shape[1]
    got = [g_area.cpu().numpy().reshape(-1, dim), g_vol.cpu().numpy().reshape(-1, dim)]
    want = [O.gradient(vert, {2: conn}, FunctionalSpec(n)) for n in ("Area", "VolumeEnclosed")]
    S = [contribution_norm_sum(n, vert, conn) for n in ("Area", "VolumeEnclosed")]
    bitwise = True
    sl, ss, ns = part.shared_local, part.shared_slot, part.n_slots
    if world > 1 and ns > 0:
        # oracle side: complete the shared rows (and their contribution sums) in ascending rank order
        buf = np.zeros((ns, 2 * dim + 2))
        buf[ss, :dim], buf[ss, dim:2 * dim] = want[0][sl], want[1][sl]
        buf[ss, 2 * dim], buf[ss, 2 * dim + 1] = S[0][sl], S[1][sl]
        t = torch.from_numpy(buf).cuda()
        allb = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allb, t)
        tot = allb[0].cpu().numpy().copy()
        for r in range(1, world):
            tot += allb[r].cpu().numpy()
        want[0][sl], want[1][sl] = tot[ss, :dim], tot[ss, dim:2 * dim]
        S[0][sl], S[1][sl] = tot[ss, 2 * dim], tot[ss, 2 * dim + 1]
        # device side: every rank that holds a shared vertex must hold the same bits
        mine = np.zeros((ns, 2 * dim + 1))
        mine[ss, :dim], mine[ss, dim:2 * dim], mine[ss, 2 * dim] = got[0][sl], got[1][sl], 1.0
        t = torch.from_numpy(mine).cuda()
        allm = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allm, t)
        allm = [a.cpu().numpy() for a in allm]
        for r in range(world):
            both = (allm[r][:, 2 * dim] > 0) & (mine[:, 2 * dim] > 0)
            if not np.array_equal(allm[r][both, :2 * dim].view(np.int64), mine[both, :2 * dim].view(np.int64)):
                bitwise = False
    rng = np.random.Generator(np.random.PCG64(4321 + rank))
    rows = np.unique(np.concatenate([sl, rng.integers(0, vert.shape[0], size=min(nsample, vert.shape[0]))])).astype(np.int64)
    worst = worst_pv = 0.0
    for a, b, s_ in zip(got, want, S):
        num = np.linalg.norm(a[rows] - b[rows], axis=1)
        scale = np.linalg.norm(b, axis=1).max()
        worst = max(worst, float(num.max() / scale))
        den = np.maximum(np.linalg.norm(b[rows], axis=1), EPS_FLOOR * s_[rows])
        worst_pv = max(worst_pv, float((num[den > 0] / den[den > 0]).max()))
    res = torch.tensor([worst, worst_pv, 0.0 if bitwise else 1.0, float(len(rows))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    res = res.cpu().numpy()
    return {"max_rel": float(res[0]), "max_rel_pervertex": float(res[1]), "shared_rows_bitwise_equal": bool(res[2] == 0.0),
            "rows_checked_per_rank": int(res[3]), "shared_rows": int(len(sl)), "against": "oracle/morpho_oracle.c (pinned to the unmodified reference)",
            "tolerance": 1e-12}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import morpho_b200 as mb
    from morpho_b200 import _lib, meshgen as mg
    from morpho_b200.partition import build_rank_mesh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL prints its version) must not get in front of the one JSON line: everything that writes to
    # file descriptor 1 goes to stderr, the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mb.init(local)
    L = mb.lib()
    # a dedicated (non-legacy) stream shared by torch/NCCL and the library, so that torch CUDA events
    # recorded on it bracket the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    _lib.check(L.mcu_set_stream(stream.cuda_stream))

    import ctypes as C
    from types import SimpleNamespace

    def build_workload(f):
        """This rank's element range of the icosphere of frequency f with its mesh mirror, the output rows, the
        shared-row exchange (N > 1) and the step itself."""
        t0 = time.time()
        part = build_rank_mesh(f, rank, world)       # this rank's element range of the global sphere (+ shared-vertex maps)
        vert, conn = part.vert, part.conn
        nel_local, nel_global = conn.shape[0], part.nel_global
        log(f"[rank {rank}] f={f}: {nel_local} of {nel_global} triangles, {vert.shape[0]} vertices, "
            f"{part.n_shared} shared ({time.time() - t0:.1f}s)")

        if args.patch_owned:
            mb.set_option("patch_max_owned", args.patch_owned)
        for kv in args.opt:
            k_, v_ = kv.split("=")
            mb.set_option(k_, int(v_))
        mesh = mb.Mesh(vert, faces=conn)
        area, vol = mb.Area()._desc(), mb.VolumeEnclosed()._desc()
        nout = vert.size
        g_area = torch.empty(nout, dtype=torch.float64, device="cuda")
        g_vol = torch.empty(nout, dtype=torch.float64, device="cuda")
        exch = None
        if world > 1:
            from morpho_b200.partition import PeerExchange
            exch = PeerExchange(part, torch, dist, ngrad=2)   # NVLink peer memory, one kernel per exchange, no collective call
            if args.fused_push and args.exchange != "nccl":
                exch.attach(mesh)                             # the gradient kernels store the shared rows into the peers' buffers themselves

        fused = exch is not None and args.fused_push and args.exchange != "nccl"

        nccl_state = {}

        def nccl_exchange():
            """Comparison arm (--exchange nccl): ONE ncclAllReduce of the packed rows of all shared vertices (zero where this
            rank does not hold the vertex), then unpack — the collective-library way to do what mcu_halo.cu does."""
            if not nccl_state:
                nccl_state["idx"] = torch.as_tensor(part.shared_local, dtype=torch.int64, device="cuda")
                nccl_state["slot"] = torch.as_tensor(part.shared_slot, dtype=torch.int64, device="cuda")
                nccl_state["buf"] = torch.zeros((part.n_slots, 6), dtype=torch.float64, device="cuda")
            idx, slot, buf = nccl_state["idx"], nccl_state["slot"], nccl_state["buf"]
            buf.zero_()
            buf.index_copy_(0, slot, torch.cat([g_area.view(-1, 3).index_select(0, idx), g_vol.view(-1, 3).index_select(0, idx)], dim=1))
            dist.all_reduce(buf)
            rows = buf.index_select(0, slot)
            g_area.view(-1, 3).index_copy_(0, idx, rows[:, :3].contiguous())
            g_vol.view(-1, 3).index_copy_(0, idx, rows[:, 3:].contiguous())

        def step_device(evs=None):
            if evs: evs[0].record(stream)
            if fused:
                exch.bind(0)
            _lib.check(L.mcu_eval_device(mesh._h, C.byref(area), None, _lib.MCU_GRADIENT, g_area.data_ptr()))
            if exch is not None and args.exchange == "overlap":
                exch.exchange_async(0, g_area.data_ptr())     # Area's shared rows complete underneath the next kernel
            if evs: evs[1].record(stream)
            if fused:
                exch.bind(1)
            _lib.check(L.mcu_eval_device(mesh._h, C.byref(vol), None, _lib.MCU_GRADIENT, g_vol.data_ptr()))
            if evs: evs[2].record(stream)
            if exch is not None:
                if args.exchange == "overlap":
                    exch.exchange_async(1, g_vol.data_ptr())
                    exch.join()
                elif args.exchange == "nccl":
                    nccl_exchange()
                else:
                    exch(g_area, g_vol)               # shared-vertex rows completed over NVLink peer memory (mcu_halo.cu)
            if evs: evs[3].record(stream)
        return SimpleNamespace(f=f, part=part, vert=vert, conn=conn, nel_local=nel_local, nel_global=nel_global, mesh=mesh, area=area, vol=vol,
                               nout=nout, g_area=g_area, g_vol=g_vol, exch=exch, fused=fused, step_device=step_device)

    t_run = time.time()
    if world > 1:
        # a rank that waits for a peer that never arrives must not hang the box: end the process group loudly instead
        def _watchdog():
            log(f"[rank {rank}] bench.py watchdog: no result after {args.watchdog} s, giving up")
            os._exit(4)
        wd = threading.Timer(float(args.watchdog), _watchdog)
        wd.daemon = True
        wd.start()
    f = frequency_for(world, args.freq) if args.scaling != "strong" else args.freq
    W = build_workload(f)
    part, vert, conn, nel_local, nel_global, mesh, area, vol = W.part, W.vert, W.conn, W.nel_local, W.nel_global, W.mesh, W.area, W.vol
    nout, g_area, g_vol, exch, fused, step_device = W.nout, W.g_area, W.g_vol, W.exch, W.fused, W.step_device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(name):
        if world > 1 and rank == 0:
            log(f"[phase +{time.time() - t_run:.1f}s] {name}")

    def first_step(w_):
        """One step, then the exchange's status word: a peer that never arrives must end the run with a message (and the
        library's dump of flags and epochs), not as a wrong parity figure or a hang."""
        if w_.exch is None:
            return
        w_.step_device()
        barrier()
        if w_.exch.status() != 0:
            raise RuntimeError("halo exchange: a peer did not arrive in time in the first step")

    first_step(W)
    # ---- parity of what is about to be timed (checker only: the CPU oracle, never the thing measured) ------
    parity = None
    if not args.no_parity:
        parity = parity_check(part, vert, conn, step_device, g_area, g_vol, torch, dist, world, rank, barrier)
        if rank == 0:
            log(f"[parity] {parity}")
        if parity["max_rel"] > 1e-12 or parity["max_rel_pervertex"] > 1e-12 or not parity["shared_rows_bitwise_equal"]:
            raise SystemExit(f"parity check failed: {parity}")

    # ---- device-resident timing -------------------------------------------------------------
    phase("warm-up")
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.lines.clear()
    launches0 = L.mcu_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    phase("timed steps")
    e_start.record(stream)
    sync_every = int(os.environ.get("BENCH_SYNC_EVERY", "0"))         # diagnostics only
    for k in range(args.steps):
        step_device(ev[k])
        if sync_every and (k + 1) % sync_every == 0:
            stream.synchronize()
    e_end.record(stream)
    barrier()
    launches = L.mcu_launch_count() - launches0
    ms_total = e_start.elapsed_time(e_end)
    ms_area = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]))
    ms_vol = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)]))
    ms_exch = float(np.mean([ev[k][2].elapsed_time(ev[k][3]) for k in range(args.steps)]))
    # nvidia-smi cannot sample faster than ~100 ms: when the timed region is shorter than that, keep the SAME
    # load running (untimed) until a few samples exist, so the clocks line describes the kernels under load
    # (every rank runs the same number of steps: the shared-vertex exchange pairs up launches across ranks)
    phase("clock samples under load")
    t_load = time.time()
    for _ in range(60):
        for _ in range(50):
            step_device()
        torch.cuda.synchronize()
        more = torch.tensor([1 if (rank == 0 and len(sampler.lines) < 6 and time.time() - t_load < 3.0) else 0], device="cuda")
        if world > 1:
            dist.broadcast(more, 0)
        if int(more.item()) == 0:
            break
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the C ABI -----------------------------------------------------------
    # (a) the reference's own arrangement of a step (round 1's figure, kept for comparison): host buffers, the vertex
    #     matrix up and both gradients down every step — PCIe-bound;
    # (b) the line-search step of optimize.morpho (energywithstepsize + totalforce, optimize.morpho:229-303, 461-486) with
    #     DEVICE-RESIDENT matrices (mcu_matrix_*): x.acc(-s, force); Area.total; VolumeEnclosed.total;
    #     Area.gradient; VolumeEnclosed.gradient — the calls the extension makes for a script's step.  Per step the
    #     host sends the step size (a kernel argument) and reads back the two energies and the status word.
    hv = torch.from_numpy(vert.copy()).pin_memory()
    ha = torch.empty(nout, dtype=torch.float64).pin_memory()
    hg = torch.empty(nout, dtype=torch.float64).pin_memory()

    def step_e2e_host():
        _lib.check(L.mcu_mesh_set_vertices(mesh._h, hv.data_ptr()))
        if exch is None:
            _lib.check(L.mcu_eval(mesh._h, C.byref(area), None, _lib.MCU_GRADIENT, ha.data_ptr()))
            _lib.check(L.mcu_eval(mesh._h, C.byref(vol), None, _lib.MCU_GRADIENT, hg.data_ptr()))
        else:                                             # partitioned: complete the shared rows before the read-back
            step_device()
            ha.copy_(g_area, non_blocking=True)
            hg.copy_(g_vol, non_blocking=True)
            stream.synchronize()

    def timed(fn, nsteps, nwarm):
        for _ in range(nwarm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(nsteps):
            fn()
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / nsteps

    e2e_steps = 1 if args.quick else max(3, min(args.steps, 10))
    phase("end to end, host buffers")
    ms_e2e_host = timed(step_e2e_host, e2e_steps, 0 if args.quick else 2)
    checksum = float(ha[:3].sum() + hg[:3].sum())

    phase("end to end, device-resident line-search step")
    dim = vert.shape[1]
    X = L.mcu_matrix_create(dim, vert.shape[0])
    F = L.mcu_matrix_create(dim, vert.shape[0])
    GA = L.mcu_matrix_create(dim, vert.shape[0])
    GV = L.mcu_matrix_create(dim, vert.shape[0])
    assert X and F and GA and GV, L.mcu_last_error()
    _lib.check(L.mcu_matrix_upload(X, hv.data_ptr()))
    _lib.check(L.mcu_mesh_bind_vertices(mesh._h, X))          # Mesh.setvertexmatrix: the mesh reads X in place
    if fused:
        exch.bind(0)
    _lib.check(L.mcu_eval_matrix(mesh._h, C.byref(area), None, _lib.MCU_GRADIENT, F))      # the force of the search direction
    if exch is not None:
        exch.exchange_ptrs(L.mcu_matrix_device(F))            # complete the shared rows of the search direction
    energies = (C.c_double * 2)()
    sgn = [1.0]

    def step_e2e_resident():
        s_ = 1e-7 * sgn[0]                                    # a trial step and its reverse: the mesh stays the headline mesh
        sgn[0] = -sgn[0]
        _lib.check(L.mcu_matrix_axpby(X, X, 1.0, F, -s_))                                               # x.acc(-s, force)
        _lib.check(L.mcu_eval(mesh._h, C.byref(area), None, _lib.MCU_TOTAL, C.byref(energies, 0)))       # 8 B D2H + status
        _lib.check(L.mcu_eval(mesh._h, C.byref(vol), None, _lib.MCU_TOTAL, C.byref(energies, 8)))
        if fused:
            exch.bind(0)
        _lib.check(L.mcu_eval_matrix(mesh._h, C.byref(area), None, _lib.MCU_GRADIENT, GA))
        if exch is not None and args.exchange == "overlap":
            exch.exchange_async(0, L.mcu_matrix_device(GA))
        if fused:
            exch.bind(1)
        _lib.check(L.mcu_eval_matrix(mesh._h, C.byref(vol), None, _lib.MCU_GRADIENT, GV))
        if exch is not None and args.exchange == "overlap":
            exch.exchange_async(1, L.mcu_matrix_device(GV))
            exch.join()
        elif exch is not None:
            exch.exchange_ptrs(L.mcu_matrix_device(GA), L.mcu_matrix_device(GV))
        _lib.check(L.mcu_mesh_status(mesh._h))                                                          # degenerate / VolEnclZero of the gradients

    res_steps = 2 if args.quick else max(10, min(args.steps, 100))
    ms_e2e = timed(step_e2e_resident, res_steps, 2)
    e_area, e_vol = float(energies[0]), float(energies[1])
    _lib.check(L.mcu_mesh_bind_vertices(mesh._h, None))
    for mtx in (X, F, GA, GV):
        L.mcu_matrix_destroy(mtx)

    if exch is not None and exch.status() != 0:
        raise RuntimeError("halo exchange: a peer did not arrive in time")

    # ---- strong scaling (N > 1): the SAME 10M-triangle sphere of BASELINE config 2 cut N ways ----------------
    strong = None
    if world > 1 and args.scaling == "both":
        phase("strong scaling")
        S = build_workload(args.freq)
        first_step(S)
        par_s = None
        if not args.no_parity:
            par_s = parity_check(S.part, S.vert, S.conn, S.step_device, S.g_area, S.g_vol, torch, dist, world, rank, barrier)
            if par_s["max_rel"] > 1e-12 or par_s["max_rel_pervertex"] > 1e-12 or not par_s["shared_rows_bitwise_equal"]:
                raise SystemExit(f"parity check failed (strong-scaling mesh): {par_s}")
        for _ in range(max(args.warmup, 3)):
            S.step_device()
        barrier()
        evs_s = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        for k in range(args.steps):
            S.step_device(evs_s[k])
        b_.record(stream)
        barrier()
        ts = torch.tensor([a_.elapsed_time(b_) / args.steps] +
                          [float(np.mean([evs_s[k][i].elapsed_time(evs_s[k][i + 1]) for k in range(args.steps)])) for i in range(3)],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ts = [float(x) for x in ts.cpu()]
        if S.exch.status() != 0:
            raise RuntimeError("halo exchange (strong-scaling mesh): a peer did not arrive in time")
        strong = {"workload": workload_string(args.freq, world), "value": 2.0 * S.nel_global / (ts[0] * 1e-3) / 1e6, "unit": UNIT,
                  "ms_per_step": ts[0], "steps": args.steps,
                  "kernel_ms": {"area_gradient": ts[1], "volumeenclosed_gradient": ts[2], "shared_row_exchange": ts[3]},
                  "shared_vertices_rank0": int(S.part.n_shared), "parity": par_s,
                  "note": "total work fixed: compare `value` with the N=1 line's (the driver's SCALE run); weak scaling is the headline `value`"}
        if rank == 0:
            log(f"[strong] {strong}")
    # ---- BASELINE config 5 partitioned over the N GPUs (N > 1; scripts/bench_configs.py) ------------------
    config5_multi = None
    if world > 1 and not args.no_configs:
        from scripts.bench_configs import config5_partitioned
        phase("config 5 partitioned")
        for w_ in (W, locals().get("S")):
            if w_ is not None:                          # release the sphere workloads first
                if w_.exch is not None:
                    w_.exch.close()
                w_.mesh = w_.g_area = w_.g_vol = None
        mesh = g_area = g_vol = None
        torch.cuda.empty_cache()
        t5 = time.time()
        config5_multi = config5_partitioned(mb, torch, dist, measured_peak()[0], n=args.tets, log=log)
        config5_multi["wall_s"] = time.time() - t5
        if rank == 0:
            log(f"[configs] 5 partitioned: {config5_multi}")

    # ---- max over ranks ---------------------------------------------------------------------
    times = torch.tensor([ms_total, ms_e2e, ms_area, ms_vol, ms_exch, ms_e2e_host], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_area, ms_vol, ms_exch, ms_e2e_host = [float(x) for x in times.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_step = ms_total / args.steps
    evals_per_step = 2.0 * nel_global
    value = evals_per_step / (ms_step * 1e-3) / 1e6
    e2e_value = evals_per_step / (ms_e2e * 1e-3) / 1e6
    peak, how = measured_peak()
    kname = "k_patch_ring"
    ms_dom, dom = (ms_area, kname + "<Area>") if ms_area >= ms_vol else (ms_vol, kname + "<VolumeEnclosed>")
    achieved = BYTES_PER_ELEM_GRADIENT * nel_local / (ms_dom * 1e-3) / 1e9
    # DRAM bytes per launch cannot be measured outside a profiler: the figure is the one of the committed
    # `ncu --set full` capture, and it is only reported while the kernel source is byte-identical to the captured one
    traffic, traffic_note = None, "no ncu capture of the current kernel source"
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(tp):
        try:
            import hashlib
            rec = json.load(open(tp))
            sha = hashlib.sha256(open(os.path.join(ROOT, "morpho_b200", "csrc", "mcu_patch.cu"), "rb").read()).hexdigest()[:16]
            if rec.get("source_sha16") == sha:
                traffic, traffic_note = rec.get("dram_bytes_per_launch"), f"ncu --set full capture {rec.get('capture')} of this kernel source"
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if (args.scaling == "strong" and world > 1) else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_string(f, world), "inputs": INPUTS,
                   "partition": ("contiguous element ranges per GPU; rows of shared vertices pushed into the peers' receive buffers over NVLink (CUDA IPC) "
                                 f"and summed in rank order; exchange mode: {args.exchange}, pushed by the {'gradient' if args.fused_push else 'exchange'} kernels") if world > 1 else "single GPU",
                   "gradient_path": "persistent patch kernel: bulk-async (TMA) staging of vertex runs into a 3-stage shared-memory ring, one consumer thread per owned vertex walks its ring (owner computes, no atomics)",
                   "kernel_ms": {"area_gradient": ms_area, "volumeenclosed_gradient": ms_vol, "shared_row_exchange": ms_exch if world > 1 else 0.0}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 2 * 8 + 3 * 4,
                "ms_per_step": ms_e2e, "steps": res_steps,
                "what": "one line-search step of optimize.morpho through the C ABI with device-resident matrices (mcu_matrix_*): "
                        "x.acc(-s, force) + Area.total + VolumeEnclosed.total + Area.gradient + VolumeEnclosed.gradient; the host sends the "
                        "step size as a kernel argument and reads back the two energies (8 B each, synchronising) and the status word; "
                        "only the 2 gradient element-evals per triangle are counted in `value`",
                "energies": [e_area, e_vol],
                "host_buffers": {"value": evals_per_step / (ms_e2e_host * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e_host,
                                 "h2d_bytes_per_step": int(vert.nbytes), "d2h_bytes_per_step": int(2 * nout * 8), "steps": e2e_steps,
                                 "what": "round 1's step: vertex matrix H2D from pinned memory, both gradients D2H, every step (PCIe-bound)",
                                 "checksum": checksum}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": f"of {how} (MEASURED_PEAKS.json hbm_gbs)" if how == "measured" else "of fallback",
                     "algorithmic_bytes_per_element": BYTES_PER_ELEM_GRADIENT, "elements_per_launch": nel_local,
                     "avg_launch_ms": ms_dom, "traffic": traffic, "traffic_source": traffic_note},
    }
    if parity is not None:
        line["parity"] = parity
    if strong is not None:
        line["strong_scaling"] = strong
    if config5_multi is not None:
        line["configs"] = {"5": config5_multi}
    if world == 1 and not args.no_configs:
        # the other single-GPU configs of BASELINE.json, measured in this same run (scripts/bench_configs.py)
        from scripts.bench_configs import run_all
        mesh = None
        line["configs"] = run_all(mb, torch, peak, log=log)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(vert, conn)
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--freq", type=int, default=707, help="icosphere frequency per GPU (707 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the gradients before timing")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3/4/5 (N=1 only)")
    ap.add_argument("--watchdog", type=int, default=900, help="N > 1: seconds after which a run that has not finished kills itself")
    ap.add_argument("--tets", type=int, default=204, help="N > 1: cells per edge of the config-5 tetrahedral cube (204 = BASELINE config 5)")
    ap.add_argument("--patch-owned", type=int, default=0, help="override the number of owned vertices per patch")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (mcu_set_option), repeatable")
    ap.add_argument("--quick", action="store_true", help="kernel timing only: skip the end-to-end leg")
    ap.add_argument("--exchange", default="serial", choices=["overlap", "serial", "nccl"],
                    help="multi-GPU completion of shared rows: ONE peer-memory exchange kernel for both gradients behind the gradient kernels "
                         "(default: 192.6 G at N = 2), one exchange per gradient on a second stream underneath the next kernel (188.9 G: the "
                         "co-running exchange slows VolumeEnclosed by 11 us), or one ncclAllReduce of the packed rows (159.0 G; comparison)")
    ap.add_argument("--scaling", default="both", choices=["both", "weak", "strong"],
                    help="N > 1: the headline line is weak scaling (N x 10M triangles); 'both' adds a strong-scaling leg (the 10M-triangle sphere "
                         "cut N ways) as `strong_scaling`; 'strong' makes the 10M-triangle sphere the only workload")
    ap.add_argument("--fused-push", action="store_true",
                    help="multi-GPU, experimental: the gradient kernels store the shared rows into the peers' buffers themselves and raise the flags "
                         "(default: the exchange kernel pushes them).  Off by default: at 10M triangles per GPU the pushing instantiation of the "
                         "VolumeEnclosed kernel stalled after 20-50 launches (scripts/halo_watch.py), not root-caused")
    ap.add_argument("--no-fused-push", action="store_true", help="accepted for compatibility (this is the default now)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
