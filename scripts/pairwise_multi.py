#!/usr/bin/env python3
"""BASELINE config 3 on N GPUs: 200 000 points, all-pairs Coulomb gradient, row-block partition (every rank holds all
positions, computes its rows).  Launch: python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 scripts/pairwise_multi.py [n].  Prints the max over ranks of the device time per gradient."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import morpho_b200 as mb
from morpho_b200 import _lib, meshgen as mg
from morpho_b200.partition import binbounds

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mb.init(local)
L = mb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
x = torch.from_numpy(mg.sphere_points(n, seed=12345)).cuda()
fake = int(os.environ.get("PW_FAKE_WORLD", 0))      # time one rank's block of an N-way partition on a single GPU
b = binbounds(n, fake if fake else world)
r0, nr = int(b[rank]), int(b[rank + 1] - b[rank])
out = torch.zeros(nr * 3, dtype=torch.float64, device="cuda")
ms = C.c_float()
for mode, name in ((_lib.MCU_GRADIENT, "gradient"), (_lib.MCU_TOTAL, "energy")):
    for _ in range(2):
        _lib.check(L.mcu_pairwise_rows_device(n, 3, x.data_ptr(), 0, None, C.c_double(0.0), mode, r0, nr, out.data_ptr()))
    L.mcu_sync()
    if world > 1:
        dist.barrier()
    L.mcu_timer_start()
    reps = 3
    for _ in range(reps):
        _lib.check(L.mcu_pairwise_rows_device(n, 3, x.data_ptr(), 0, None, C.c_double(0.0), mode, r0, nr, out.data_ptr()))
    L.mcu_timer_stop(C.byref(ms))
    t = torch.tensor([ms.value / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t.item()) * 1e-3
        print(f"pairwise {name:8s} n={n} on {world} GPU(s): {sec * 1e3:8.2f} ms  {float(n) * n / sec / 1e9:8.1f} G ordered pairs/s")
if world > 1:
    dist.destroy_process_group()
