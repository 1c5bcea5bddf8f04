"""The N>1 path on CPU: world_size-2 `gloo` run of the row-split / gather / assemble logic of
sim5_b200/dist.py.  Each rank traces its interleaved row blocks with a CPU checker standing in for its GPU
(test infrastructure), the planes are gathered on rank 0 and must equal the single-process image."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

import harness as H
from sim5_b200 import abi, dist as sdist


def test_split_plan_covers_every_row_once():
    for ny, world, br in ((4096, 8, 32), (128, 4, 8), (64, 2, 32), (96, 3, 4)):
        rows = [sdist.local_rows(ny, r, world, br) for r in range(world)]
        allrows = sorted(sum(rows, []))
        assert allrows == list(range(ny))
        parts = [np.array(r, dtype=np.float64)[:, None] * np.ones((1, 3)) for r in rows]
        full = sdist.assemble(parts, world, br)
        assert np.array_equal(full[:, 0], np.arange(ny))
    with pytest.raises(ValueError):
        sdist.check_split(100, 8, 32)
    assert sdist.lattice_range(2048, 7, 8) == (1792, 2048)
    assert sdist.lattice_range(5, 3, 4) == (5, 5)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = abi.default_params(1, n)
        rows = sdist.apply_split(p, rank, world, 8)
        run = H.run_ref if H.have_ref() else H.run_hostsim
        planes, _, _ = run(p, 1)                      # full-image indexing on the host: pick my rows
        mine = sdist.local_rows(n, rank, world, 8)
        out = {}
        for k in ("r", "g", "flux"):
            loc = torch.from_numpy(planes.image(k)[mine].copy())
            assert loc.shape == (rows, n)
            gl = [torch.empty_like(loc) for _ in range(world)] if rank == 0 else None
            tdist.gather(loc, gl, dst=0)
            if rank == 0:
                out[k] = sdist.assemble(gl, world, 8).numpy()
        if rank == 0:
            q.put(out)
    finally:
        tdist.destroy_process_group()


def test_two_rank_gloo_gather_reassembles_the_image():
    n, world = 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    out = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = abi.default_params(1, n)
    run = H.run_ref if H.have_ref() else H.run_hostsim
    full, _, _ = run(p, 1)
    for k in out:
        assert np.array_equal(out[k], full.image(k)), k
