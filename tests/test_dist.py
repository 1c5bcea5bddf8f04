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


def _hist_images(p, images):
    """the lattice histogram with only `images` filled (CPU checker standing in for the rank's GPU)"""
    import ctypes as C
    nb = p.n_bins
    acc = np.zeros(p.n_spin * p.n_incl * nb)
    for img in images:
        q = abi.ImageParams.from_buffer_copy(p)
        q.lattice_begin, q.lattice_end = img, img + 1
        h = np.zeros_like(acc)
        hp = h.ctypes.data_as(C.POINTER(C.c_double))
        if H.have_ref():
            assert H.load_ref().ref_trace_histogram(C.byref(q), hp, 1, 1) >= 0
        else:
            assert H.load_oracle().orc_trace_histogram(C.byref(q), hp, 1) >= 0
        acc[img * nb:(img + 1) * nb] = h[img * nb:(img + 1) * nb]
    return acc


def _hist_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = abi.default_params(5, 24)
        p.n_spin, p.n_incl, p.n_bins = 3, 2, 16
        mine = sdist.lattice_images(p.n_spin * p.n_incl, rank, world)
        t = torch.from_numpy(_hist_images(p, mine))
        tdist.reduce(t, dst=0, op=tdist.ReduceOp.SUM)          # the one collective of cfg 5
        if rank == 0:
            q.put((t.numpy().copy(), mine))
    finally:
        tdist.destroy_process_group()


@pytest.mark.skipif(not (H.have_ref() or H.have_oracle()), reason="no CPU checker built")
def test_two_rank_gloo_reduce_of_the_interleaved_lattice():
    """cfg 5 at N > 1: image i of the lattice goes to rank i mod W, the per-rank histograms (zero outside the rank's images) are
    reduced with one sum -- the host-side logic of `bench.py --config 5 --gpus N` and of sim5_trace_image_multi."""
    assert sdist.lattice_images(7, 1, 3) == [1, 4] and sorted(sum((sdist.lattice_images(2048, r, 8) for r in range(8)), [])) == list(range(2048))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_hist_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    red, mine0 = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert mine0 == [0, 2, 4]
    p = abi.default_params(5, 24)
    p.n_spin, p.n_incl, p.n_bins = 3, 2, 16
    full = _hist_images(p, range(6))
    assert np.array_equal(red, full)        # disjoint images: the sum adds zeros, bit for bit
    assert np.all(red.reshape(6, 16).sum(axis=1) > 0)
