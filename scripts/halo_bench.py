#!/usr/bin/env python3
"""Micro-benchmark of the peer-memory shared-row exchange alone (torchrun, one rank per GPU)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import morpho_b200 as mb
from morpho_b200 import _lib
from morpho_b200.partition import PeerExchange, build_rank_mesh
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mb.init(local); L = mb.lib()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); _lib.check(L.mcu_set_stream(stream.cuda_stream))
f = int(round(707 * np.sqrt(world)))
rm = build_rank_mesh(f, rank, world, dist=dist, torch=torch)
g = [torch.zeros(rm.vert.size, dtype=torch.float64, device="cuda") for _ in range(2)]
ex = PeerExchange(rm, torch, dist, ngrad=2)
for _ in range(20): ex(*g)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
n = 500
for _ in range(n): ex(*g)
e1.record(stream); torch.cuda.synchronize()
print(f"rank {rank}: {rm.n_shared} shared vertices, exchange alone {e0.elapsed_time(e1) / n * 1e3:.2f} us", file=sys.stderr)
# with a dummy kernel of ~100 us in front (ranks free-running, as in the bench)
a = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def work(): a.mul_(1.0001)
for _ in range(10): work(); ex(*g)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record(stream)
for _ in range(n): work()
e1.record(stream); torch.cuda.synchronize(); t_work = e0.elapsed_time(e1) / n
dist.barrier(); torch.cuda.synchronize()
e0.record(stream)
for _ in range(n): work(); ex(*g)
e1.record(stream); torch.cuda.synchronize(); t_both = e0.elapsed_time(e1) / n
print(f"rank {rank}: work {t_work * 1e3:.1f} us, work+exchange {t_both * 1e3:.1f} us -> exchange adds {(t_both - t_work) * 1e3:.1f} us", file=sys.stderr)
assert ex.status() == 0
ex.close(); dist.destroy_process_group()
