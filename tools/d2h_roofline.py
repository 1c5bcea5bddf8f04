#!/usr/bin/env python
"""PCIe / host-memory floor of the e2e number, at N ranks (python tools/d2h_roofline.py, or under torchrun with N processes):
every rank copies ITS share of one image's planes (553 MB at 4096^2, 1/N per rank) from its GPU's memory into page-locked host memory,
all ranks at the same time, nothing else running -- the floor of any host-plane call at that N.  Two destinations: each rank's own
cudaHostAlloc'd planes, and the ONE shared image (POSIX shm + cudaHostRegister) that bench.py's e2e leg fills.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 4096
rows = n // world
planes = [(torch.float64, 8)] * 4 + [(torch.uint8, 1)]
dev = [torch.zeros((rows, n), dtype=dt, device="cuda") for dt, _ in planes]
own = [torch.empty((rows, n), dtype=dt).pin_memory() for dt, _ in planes]
nbytes = sum(t.numel() * t.element_size() for t in dev)
from multiprocessing import shared_memory, resource_tracker  # noqa: E402
names = [None]
segs = []
if rank == 0:
    segs = [shared_memory.SharedMemory(create=True, size=n * n * sz) for _, sz in planes]
    names = [[s.name for s in segs]]
if world > 1:
    dist.broadcast_object_list(names, src=0)
if rank != 0:
    segs = [shared_memory.SharedMemory(name=nm) for nm in names[0]]
    for s in segs:
        try:
            resource_tracker.unregister(s._name, "shared_memory")
        except Exception:
            pass
cudart = torch.cuda.cudart()
shared = []
for (dt, sz), s in zip(planes, segs):
    full = torch.frombuffer(s.buf, dtype=dt, count=n * n).view(n, n)
    err = cudart.cudaHostRegister(full.data_ptr(), n * n * sz, 1)       # 1 = cudaHostRegisterPortable
    assert int(err) == 0, "cudaHostRegister failed: %r" % (err,)
    shared.append(full[rank * rows:(rank + 1) * rows])
    shared[-1].zero_()           # first touch by the owner


def fence():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(dst):
    best = 1e30
    for _ in range(5):
        fence()
        t0 = time.perf_counter()
        for d, h in zip(dev, dst):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


t_own = timed(own)
t_shared = timed(shared)
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_rank": nbytes, "bytes_total": nbytes * world,
                      "own_pinned_ms": round(t_own * 1e3, 3), "own_pinned_gb_s_aggregate": round(nbytes * world / t_own / 1e9, 2),
                      "shared_registered_ms": round(t_shared * 1e3, 3), "shared_registered_gb_s_aggregate": round(nbytes * world / t_shared / 1e9, 2),
                      "e2e_floor_rays_per_s": round(n * n / min(t_own, t_shared), 0)}))
fence()
del shared
if world > 1:
    dist.barrier()
for s in segs:
    try:
        s.close()
    except BufferError:
        pass
if rank == 0:
    for s in segs:
        try:
            s.unlink()
        except Exception:
            pass
if world > 1:
    dist.destroy_process_group()
