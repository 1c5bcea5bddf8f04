#!/usr/bin/env python
"""bench.py -- benchmarks of the SIM5 photon hot path on B200, one line of JSON per run.

  python bench.py [--config C] --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference [--config C] --gpus N --steps K ...   the reference's own CPU code on the host cores

--config picks one of BASELINE.json's five configurations at its SURVEY.md 8(d) size (default 2, the one the metric is quoted on):
  1  thin-disk image 512^2, a=0.9, i=70 deg: r, g, F g^4, status                           (example 04)
  2  thin-disk image 4096^2, a=0.998, i=75 deg: r, phi, g, F g^4, status                   (headline)
  3  polarized image 2048^2, a=0.94: + chi (Walker-Penrose), delta (Chandrasekhar)
  4  stepwise raytrace() through the torus, 1024^2: I, tau, steps, status
  5  transfer-function lattice 64 spins x 32 inclinations x 1024^2 rays -> [64][32][256] g-histogram, reduced over the GPUs
One STEP = one pass of the hot path over the whole configuration (one image / one lattice).  At N > 1 the rows (lattice images) are
dealt to the ranks interleaved; images are assembled in rank 0's HBM by peer stores over NVLink, the histogram is NCCL-reduced to
rank 0 inside the step.  Total work per step is fixed: strong scaling.

Metric: geodesics (rays) per second.
  value : device-resident -- results stay in HBM; CUDA events on the launch stream, max over ranks
  e2e   : through the public host API with HOST planes: parameter upload and device->host copy of every plane inside the timed
          region, every step; at N > 1 every rank copies its rows into ONE shared pinned host image that rank 0 reads
Roofline: FP64 non-tensor pipe (no dense contraction; <= 44 B written and 0 B read per ray -> compute bound);
  achieved = F_alg (frozen) x units per launch / CUDA-event time of the dominant kernel,
  peak = live DFMA-chain microbenchmark (MEASURED_PEAKS.json has no FP64 entry).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "geodesics_per_second"
UNIT = "rays/s"
# algorithmic flop per unit, frozen in BASELINE.md section 4 (SURVEY.md 8d convention: add/mul 1, FMA 2, / 15, sqrt 13, libm ~60)
F_ALG_TRACE = 3.7e3       # ray through phase A (roots, K(mm), cn^-1, crossing, r, g, F g^4): 3 rf + 1 sncndn + roots + glue
F_ALG_POL = 0.6e3         # + emitter tetrad, emission angle, Walker-Penrose constant, chi per polarized hit
F_ALG_AZ_RR = 13.1e3      # disk hit through the bit-faithful azimuth: 3 rf + 6 rj + 2 sncndn + glue
F_ALG_AZ_FAST = 2.9e3     # disk hit through the tolerance-mode azimuth (default): tools/count_azimuth_ops.py
F_ALG_STEP = 1.3e3        # one raytrace() step

CONFIGS = {
    1: {"planes": ("r", "g", "flux", "status"), "label": "BASELINE configs[0]: Novikov-Thorne thin-disk image %dx%d, a=0.9, i=70deg, rmax=r_ms+8, outputs r/g/F*g^4/status, crossing orders 0-1"},
    2: {"planes": ("r", "phi", "g", "flux", "status"), "label": "BASELINE configs[1]: Novikov-Thorne thin-disk image %dx%d, a=0.998, i=75deg, rmax=r_ms+20, outputs r/phi/g/F*g^4/status, crossing orders 0-1"},
    3: {"planes": ("r", "phi", "g", "flux", "chi", "delta", "status"), "label": "BASELINE configs[2]: polarized disk image %dx%d, a=0.94, i=75deg, rmax=r_ms+20, Walker-Penrose chi + Chandrasekhar delta, outputs r/phi/g/flux/chi/delta/status"},
    4: {"planes": ("intensity", "tau", "steps", "status"), "label": "BASELINE configs[3]: stepwise raytrace() through the optically thin torus, %dx%d, a=0.9, i=60deg, r0=50, precision_factor=0.01, outputs I/tau/steps/status"},
    5: {"planes": (), "label": "BASELINE configs[4]: transfer-function lattice 64 spins x 32 inclinations x %dx%d rays, [64][32][256] g-histogram weighted by F*g^4*dA"},
}
# bounded CPU samples (reference arm and cpu_baseline leg): cfg 1-3 are timed at FULL size; cfg 4 on every 8th 32-row block of the
# same image, cfg 5 on every 128th image of the same lattice (about 4-6 s per step on 16 threads instead of 35 s / 12 min)
CPU_ROW_STRIDE = {4: 8}
CPU_LATTICE_STRIDE = 128


def workload_params(abi, cfg, n=None):
    p = abi.default_params(cfg)
    if n:
        p.nx = p.ny = n
    return p


def bytes_per_ray(abi, p):
    return sum(C.sizeof(ct) for name, bit, ct in abi.PLANES if p.outputs & bit)


def rays_of(p, abi):
    if p.mode == abi.MODE_HISTOGRAM:
        return p.n_spin * p.n_incl * p.nx * p.ny
    return p.nx * p.ny


def config_dict(abi, cfg, p, n_gpus, gather="peer", extra=None):
    n = p.nx
    bpr = bytes_per_ray(abi, p)
    if cfg == 5:
        par = "one GPU, whole lattice" if n_gpus == 1 else "lattice images dealt to %d GPUs one by one (image i -> rank i mod N); the 4 MB histograms are NCCL-reduced (sum) to rank 0 inside the step" % n_gpus
        l2 = "no input reads (rays are generated from the pixel index); no per-pixel output, the 4 MB histogram is accumulated with atomics in L2"
    else:
        par = ("one GPU, whole image" if n_gpus == 1 else
               "rows interleaved over %d GPUs in 32-row blocks; %s" % (n_gpus, "every rank's rows land in rank 0's image planes over NVLink peer memory (CUDA IPC; copy-engine transfers of the finished row blocks under the next step's kernels), barrier at the end"
                                                                       if gather == "peer" else "NCCL gather of compact planes to rank 0 + re-assembly"))
        mb = n * n * bpr // 1000000
        l2 = "no input reads (rays are generated from the pixel index); %d MB of output planes per step %s 126 MB L2, written with streaming stores" % (mb, ">" if mb > 126 else "<")
    d = {"workload": CONFIGS[cfg]["label"] % (n, n), "baseline_config": cfg, "rays_per_step": rays_of(p, abi), "parallelism": par, "l2": l2}
    if extra:
        d.update(extra)
    return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe), one sample every 20 ms."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.proc = None
        self.marks = {}

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def mark(self, name):
        self.marks[name] = time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)

        def digest(rows):
            sm, smax, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); smax.append(float(r[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    continue
            return sm, smax, reasons
        sm, smax, reasons = digest(self.rows)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
        t0, t1 = self.marks.get("timed_begin"), self.marks.get("timed_end")
        if t0 and t1:
            inside, _, r_in = digest([x for x in self.rows if t0 - 0.02 <= x[0] <= t1 + 0.02])
            out["samples_in_device_timed_region"] = len(inside)
            if inside:
                out["sm_mhz_in_device_timed_region"] = statistics.median(inside)
                out["reasons_in_device_timed_region"] = sorted(r_in)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own CPU code (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_runner(cfg, n=None):
    """Returns (kind, cores, sample description, rays per pass, fn() -> seconds of one pass over the CPU sample of config cfg)."""
    import harness as H
    import numpy as np
    from sim5_b200 import abi
    try:
        cores = len(os.sched_getaffinity(0))      # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    except AttributeError:
        cores = os.cpu_count() or 1
    p = workload_params(abi, cfg, n)
    if H.have_ref():
        kind = "reference"
    elif H.have_oracle() and cfg in (1, 2, 3, 5):
        kind = "port"
    else:
        return None
    if cfg == 5:
        nimg = p.n_spin * p.n_incl
        imgs = list(range(CPU_LATTICE_STRIDE // 2, nimg, CPU_LATTICE_STRIDE)) if nimg >= CPU_LATTICE_STRIDE else list(range(nimg))
        hist = np.zeros(nimg * p.n_bins)
        hp = hist.ctypes.data_as(C.POINTER(C.c_double))
        lib = H.load_ref() if kind == "reference" else H.load_oracle()

        def run():
            t = 0.0
            for img in imgs:
                p.lattice_begin, p.lattice_end = img, img + 1
                dt = lib.ref_trace_histogram(C.byref(p), hp, cores, 1) if kind == "reference" else lib.orc_trace_histogram(C.byref(p), hp, cores)
                assert dt >= 0
                t += dt
            return t
        rays = len(imgs) * p.nx * p.ny
        sample = "%d of the %d lattice images at full %dx%d size (every %dth image of the same lattice: all spins' range, all inclinations' range)" % (len(imgs), nimg, p.nx, p.ny, CPU_LATTICE_STRIDE)
        return kind, cores, sample, rays, run
    stride = CPU_ROW_STRIDE.get(cfg, 1)
    if stride > 1 and p.ny % (stride * 32) == 0:
        p.split_count, p.split_index, p.split_rows = stride, stride // 2, 32
        rays = p.nx * p.ny // stride
        sample = "every %dth 32-row block of the SAME %dx%d image (rows interleaved over the whole image, %d rays per pass)" % (stride, p.nx, p.ny, rays)
    else:
        rays = p.nx * p.ny
        sample = "the full %dx%d image, nothing reduced" % (p.nx, p.ny)
    pl = H.Planes(p)
    st = abi.TraceStats()
    lib = H.load_ref() if kind == "reference" else H.load_oracle()

    def run():
        dt = lib.ref_trace_image(C.byref(p), C.byref(pl.out), cores, 1, C.byref(st)) if kind == "reference" else lib.orc_trace_image(C.byref(p), C.byref(pl.out), cores, C.byref(st))
        assert dt >= 0
        return dt
    return kind, cores, sample, rays, run


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from sim5_b200 import abi
    r = cpu_runner(args.config, args.size)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libsim5ref.so nor oracle/libsim5oracle.so is built"}))
        return 0
    kind, cores, sample, rays, run = r
    for _ in range(args.warmup):
        run()
    t = 0.0
    for _ in range(args.steps):
        t += run()
    v = rays * args.steps / t
    p = workload_params(abi, args.config, args.size)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(abi, args.config, p, args.gpus, args.gather,
                              {"rays_timed_per_step": rays, "cpu_sample": sample,
                               "note": "ms_per_step is the time of one pass over rays_timed_per_step rays; rays/s is an intensive rate"}),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs next to its GPU (nvidia-smi topo), so that its pinned staging and the pages of the
    shared host image it first touches live on that NUMA node.  Returns the description of what was done."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            cols = line.split()
            if cols and cols[0] == "GPU%d" % local:
                for c in cols[1:]:
                    if c and c[0].isdigit() and ("-" in c or "," in c) and all(ch.isdigit() or ch in "-," for ch in c):
                        cpus = set()
                        for part in c.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                        cpus &= os.sched_getaffinity(0)
                        if cpus:
                            os.sched_setaffinity(0, cpus)
                            return "cpus %s" % c
        return "no affinity column for GPU%d" % local
    except Exception as e:
        return "unbound (%s)" % e


def run_ours(args):
    # stdout carries exactly ONE line, the JSON: everything native libraries print on the way (NCCL's version banner, ...) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from sim5_b200 import abi, api, dist as sdist

    cfg = args.config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1")
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    api.init(local)
    L = api.lib()
    stream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: the library launches on it, torch events time it
    torch.cuda.set_stream(stream)
    api.check(L.sim5_set_stream(C.c_void_p(stream.cuda_stream)), "sim5_set_stream")

    p = workload_params(abi, cfg, args.size)
    n = p.nx
    rays_step = rays_of(p, abi)
    bpr = bytes_per_ray(abi, p)
    names = CONFIGS[cfg]["planes"]
    st = abi.TraceStats()
    hist_mode = (cfg == 5)
    two_phase = "phi" in names
    rows = n
    if world > 1:
        if hist_mode:
            p.split_count, p.split_index = world, rank
        else:
            rows = sdist.apply_split(p, rank, world)
    rays_local = (len(range(rank, p.n_spin * p.n_incl, world)) * n * n) if hist_mode else rows * n
    p.flags = abi.FLAG_DEVICE_PTRS | (0 if hist_mode else abi.FLAG_ASYNC) | (abi.FLAG_EXACT_AZIMUTH if args.exact_azimuth else 0) | (abi.FLAG_ROW_MAJOR if args.row_major else 0)
    # a train of images: the redo wave of step k can run beside the kernels of step k+1 (SIM5_FLAG_DEFER_REDO); it pays when the wave is a
    # large part of the step (multi-GPU split), on one GPU it is neutral to slightly negative -> by default for N > 1 only
    defer = two_phase and ((world > 1) if args.defer_redo == "auto" else (args.defer_redo == "on"))
    if defer:
        p.flags |= abi.FLAG_DEFER_REDO | (abi.FLAG_ALT_STREAMS if args.alt_streams else 0)
    peer = world > 1 and args.gather == "peer" and not hist_mode
    gath, image, loc, hist = None, None, None, None
    out = abi.ImageOut()
    if hist_mode:
        hist = torch.zeros(p.n_spin * p.n_incl * p.n_bins, dtype=torch.float64, device=dev)
        out.hist = hist.data_ptr()
    elif peer:
        # the image lives in rank 0's HBM; every rank maps those planes (CUDA IPC, peer access over NVLink/NVSwitch) and its kernels store
        # their interleaved row blocks straight into the final image: the transfer rides under the FP64 work, no gather, no re-assembly
        p.flags |= abi.FLAG_FULL_INDEX
        if rank != 0 and (args.peer_copy == "dma" or (args.peer_copy == "auto" and world >= 8)):
            # ... by DMA: the rank traces into local compact planes and the copy engine moves the finished row blocks into rank 0's image
            # under the next step's kernels (8-byte stores from 7 kernels at once saturate rank 0's NVLink ingress and stretch phase A by 50 %)
            p.flags |= abi.FLAG_STAGE_COPY
        if rank == 0:
            image = api.DevicePlanes(p, names=names)
            blob = [image.handles()]
        else:
            blob = [None]
        dist.broadcast_object_list(blob, src=0)
        if rank != 0:
            image = api.DevicePlanes.from_handles(p, blob[0])
        out = image.out
    else:
        tdt = {"steps": torch.int32, "status": torch.uint8}
        loc = {k: torch.empty((rows, n), dtype=tdt.get(k, torch.float64), device=dev) for k in names}
        for k, t in loc.items():
            setattr(out, k, t.data_ptr())
        if world > 1 and rank == 0:
            gath = {k: [torch.empty_like(t) for _ in range(world)] for k, t in loc.items()}
    kev = []

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        api.check(L.sim5_trace_image(C.byref(p), C.byref(out), C.byref(st)), "sim5_trace_image")
        if timed:
            e1.record()
            kev.append((e0, e1))
        if hist_mode and world > 1:
            dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)                    # NCCL over NVLink: the one collective of the path, inside the step
        if world > 1 and not peer and not hist_mode:
            api.check(L.sim5_join(), "sim5_join")      # the gather reads phi: this step's redo passes must be behind it on the stream
            for k, t in loc.items():
                dist.gather(t, gath[k] if rank == 0 else None, dst=0)
                if rank == 0:
                    sdist.assemble(gath[k], world)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf = api.fp64_peak_tflops(local, 8192) if rank == 0 else None
    api.check(L.sim5_set_stream(C.c_void_p(stream.cuda_stream)), "sim5_set_stream")

    # one synchronous call: the counters of the workload (disk hits, raytrace() steps: the same every step)
    q = abi.ImageParams.from_buffer_copy(p)
    q.flags &= ~(abi.FLAG_ASYNC | abi.FLAG_DEFER_REDO)
    st_sync = abi.TraceStats()
    api.check(L.sim5_trace_image(C.byref(q), C.byref(out), C.byref(st_sync)), "sim5_trace_image")
    total_steps_local = int(st_sync.total_steps)
    hits_local = int(st_sync.class_count[0] + st_sync.class_count[1] + st_sync.class_count[5])
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(False)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    fence()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark("timed_begin")
    ev0.record()
    for _ in range(args.steps):
        step(True)                 # SIM5_FLAG_ASYNC: the K steps are enqueued back to back, no host sync in between
    api.check(L.sim5_join(), "sim5_join")      # the launch stream waits for the last steps' deferred redo passes: they are inside the timed region
    ev1.record()
    fence()
    if sampler:
        sampler.mark("timed_end")
    phase_ms, phase_items, launches = [0.0, 0.0, 0.0], [0, 0], 0
    if hist_mode:
        launches = args.steps
    else:
        # per-kernel CUDA events the library recorded on the launch stream for each of the timed calls (it keeps the last 63),
        # read back after the timed region so that the steps above run without a synchronisation per step
        nread = min(args.steps, 63)
        for back in range(nread):
            pm, nk = api.phase_history(back)
            for i, v in enumerate(pm):
                phase_ms[i] += v * (args.steps / nread)
            launches += nk
        launches = launches * args.steps // nread
        if two_phase:
            _, items, _ = api.last_phase_ms()
            phase_items[0], phase_items[1] = items
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    # device time of the kernels of one step: CUDA events around each call; a deferred train leaves its redo passes outside those
    # brackets (they run beside the next call), so there the per-step kernel time is the whole timed region / K
    kms = torch.tensor([(ev0.elapsed_time(ev1) / args.steps) if defer else (sum(a.elapsed_time(b) for a, b in kev) / len(kev))], dtype=torch.float64, device=dev)
    cnt = torch.tensor([launches, total_steps_local, hits_local], dtype=torch.float64, device=dev)
    # per-rank table: this rank's device time of the timed region and of its kernels per step, its units of work
    mine_t = torch.tensor([ev0.elapsed_time(ev1) / args.steps, phase_ms[0] / args.steps, phase_ms[1] / args.steps, phase_ms[2] / args.steps,
                           float(rays_local), float(total_steps_local), float(hits_local)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine_t) for _ in range(world)] if world > 1 else [mine_t]
    if world > 1:
        dist.all_gather(per_rank, mine_t)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    kernel_ms = float(kms.item())
    launches_all, total_steps_all = int(cnt[0].item()), int(cnt[1].item())
    value = rays_step * args.steps / (total_ms * 1e-3)
    hist_dev = hist.cpu().numpy().copy() if (hist_mode and rank == 0) else None

    # ---- e2e: the public host API with HOST buffers; H2D of the parameters + D2H of every result byte inside the timed region
    api.check(L.sim5_set_stream(None), "sim5_set_stream")
    ph = workload_params(abi, cfg, args.size)
    if args.exact_azimuth:
        ph.flags |= abi.FLAG_EXACT_AZIMUTH
    e2e_note, image_check, shm = None, None, None
    if hist_mode:
        if world > 1:
            ph.split_count, ph.split_index = world, rank
        hh = torch.empty(p.n_spin * p.n_incl * p.n_bins, dtype=torch.float64).pin_memory()
        if world == 1:
            oh = abi.ImageOut(); oh.hist = hh.data_ptr()
            e2e_note = "sim5_trace_image with a HOST histogram: 4 MB D2H per step"

            def e2e_step():
                api.check(L.sim5_trace_image(C.byref(ph), C.byref(oh), C.byref(st)), "sim5_trace_image")
        else:
            ph.flags |= abi.FLAG_DEVICE_PTRS
            e2e_note = "every rank traces its images, NCCL reduce to rank 0, rank 0 copies the 4 MB histogram to pinned host memory, every step"

            def e2e_step():
                api.check(L.sim5_trace_image(C.byref(ph), C.byref(out), C.byref(st)), "sim5_trace_image")
                dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    hh.copy_(hist, non_blocking=False)
        d2h = p.n_spin * p.n_incl * p.n_bins * 8
    else:
        if world > 1:
            sdist.apply_split(ph, rank, world)
            # ONE host image for the whole job: a POSIX shared-memory segment per plane, mapped and page-locked (cudaHostRegister)
            # by every rank; each rank's chunks go over its own PCIe link straight into the rows it owns, rank 0 reads the image
            try:
                shm = SharedHostImage(api, abi, ph, rank, world, dist)
                hp = shm.planes
                e2e_note = "every rank copies its rows (chunks under its kernels, own PCIe link) into ONE shared %s host image (POSIX shm + cudaHostRegister) that rank 0 reads; ranks bound to their GPU's NUMA node (%s)" % ("pageable (registration failed)" if shm.unregistered else "page-locked", numa)
            except RuntimeError:
                shm = None
                hp = api.HostPlanes(ph, pinned=True)
                e2e_note = "no shared-memory segment available: every rank copies its own rows to pinned host planes of its own (no process holds the whole host image)"
        else:
            hp = api.HostPlanes(ph, pinned=True)
            e2e_note = "sim5_trace_image with pinned HOST planes: chunks of ~2^21 rays, the copy of chunk k under the kernels of chunk k+1"
        if world > 1:
            mine = np.asarray(sdist.local_rows(ph.ny, rank, world))
            for a in hp.arrays.values():
                a.reshape(ph.ny, ph.nx)[mine] = 0      # first touch by the rank that owns the rows: the pages land on its NUMA node
            dist.barrier()
        else:
            for a in hp.arrays.values():
                a[...] = 0

        def e2e_step():
            api.trace_image(ph, hp)
        d2h = (rays_step // world) * bpr
    for _ in range(2):
        e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    fence()
    clocks = sampler.stop() if sampler else None
    e2e_value = rays_step * args.steps / float(e2e_s.item())
    checksum = None
    if hist_mode:
        if rank == 0:
            hsum, dsum = float(hh.numpy().sum()), float(hist_dev.sum())
            image_check = {"sum_hist_host_e2e": hsum, "sum_hist_device_resident": dsum, "rel_diff": abs(hsum - dsum) / max(abs(dsum), 1e-300)}
            assert image_check["rel_diff"] < 1e-9, image_check
            checksum = hsum
    else:
        key = "g" if "g" in names else "intensity"
        if rank == 0:
            checksum = float(np.nansum(hp[key]))                 # the WHOLE host image, as rank 0 holds it
        if peer and rank == 0 and (world == 1 or shm is not None):
            # the image the ranks assembled in rank 0's HBM through peer stores, against the host image all ranks filled
            dsum = float(np.nansum(image.to_host(key)))
            image_check = {"plane": key, "sum_device_image": dsum, "sum_host_image": checksum, "rel_diff": abs(dsum - checksum) / max(abs(checksum), 1e-300)}
            if two_phase:
                dphi, hphi = float(np.nansum(np.abs(image.to_host("phi")))), float(np.nansum(np.abs(hp["phi"])))
                image_check.update({"sum_abs_phi_device_image": dphi, "sum_abs_phi_host_image": hphi, "rel_diff_phi": abs(dphi - hphi) / max(abs(hphi), 1e-300)})
                assert image_check["rel_diff_phi"] < 1e-12, image_check
            assert image_check["rel_diff"] < 1e-12, image_check
            # plane by plane, bit for bit: the one host image all ranks filled over their PCIe links == the one device image they assembled over NVLink
            image_check["host_image_equals_device_image"] = bool(all(np.array_equal(image.to_host(k), hp[k], equal_nan=True) for k in names))
            assert image_check["host_image_equals_device_image"], image_check

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            # the reference's own CPU code beside it (N = 1 only), bounded sample: 1 timed pass after 1 warm-up
            try:
                r = cpu_runner(cfg, args.size)
                kind, cores, sample, crays, run = r
                run()
                dt = run()
                cpu = {"value": crays / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + "; all host threads, 1 timed pass after 1 warm-up"}
            except Exception as e:  # the checker is optional for the benchmark itself
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        ph_ms = [v / args.steps for v in phase_ms]
        roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tf,
                "peak_source": "live FP64 DFMA-chain microbenchmark (sim5_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 entry",
                "achieved_is": "ALGORITHMIC flop (BASELINE.md section 4, a model count frozen there) / CUDA-event time of the kernel; not a counter reading"}
        if hist_mode:
            kname, units, per_unit, what, k_ms = "k_trace_histogram", rays_local, F_ALG_TRACE, "rays", kernel_ms
            flop_step = F_ALG_TRACE * rays_local
        elif cfg == 4:
            kname, units, per_unit, what, k_ms = "k_trace_lanes<StepwiseProg>", total_steps_local, F_ALG_STEP, "raytrace() steps", ph_ms[0] or kernel_ms
            flop_step = F_ALG_STEP * total_steps_local
            roof["steps_per_ray"] = total_steps_all / float(rays_step)
        else:
            f_a = F_ALG_TRACE + (F_ALG_POL * hits_local / max(rays_local, 1) if cfg == 3 else 0.0)
            f_rr = F_ALG_AZ_RR if args.exact_azimuth else F_ALG_AZ_FAST
            nhit = phase_items[0] + phase_items[1]
            flop_step = f_a * rays_local + (f_rr * nhit if two_phase else 0.0)
            knames = (("k_trace_eqplane<DEFER>", "k_azimuth<RR>", "k_azimuth<RC>") if args.exact_azimuth else ("k_trace_eqplane<DEFER>", "k_azimuth_fast", "k_azimuth<RR|RC> redo")) \
                if two_phase else ("k_trace_eqplane", "-", "-")
            dom = max(range(3), key=lambda i: ph_ms[i]) if two_phase else 0
            if dom == 0:
                kname, units, per_unit, what = knames[0], rays_local, f_a, "rays"
            elif dom == 1:
                kname, units, per_unit, what = knames[1], (phase_items[0] if args.exact_azimuth else nhit), f_rr, "disk hits"
            else:
                kname, units, per_unit, what = knames[2], phase_items[1], F_ALG_AZ_RR, "RC disk hits"
            k_ms = ph_ms[dom] or kernel_ms
            roof["kernels_ms"] = dict(zip(knames, ph_ms))
            if two_phase:
                roof["azimuth_items"] = {"rr": phase_items[0], "rc": phase_items[1]}
        achieved = per_unit * units / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        achieved_step = flop_step / (kernel_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                traffic = json.load(fh).get(kname)
        except Exception:
            pass
        roof.update({"achieved": achieved, "frac": achieved / peak_tf if (peak_tf and achieved) else None, "traffic": traffic,
                     "kernel": kname, "kernel_ms": k_ms, "units_per_launch": units, "unit_kind": what, "flop_per_unit": per_unit,
                     "step": {"achieved": achieved_step, "frac": achieved_step / peak_tf if peak_tf else None,
                              "flop_per_ray": flop_step / max(rays_local, 1), "kernels_ms_total": kernel_ms,
                              "note": "the flops of the algorithm actually run, all kernels of the step, per rank"},
                     "hbm_written_bytes_per_launch": (rays_local * bpr) if not hist_mode else p.n_spin * p.n_incl * p.n_bins * 8})
        roof["per_rank"] = [dict(zip(("ms_per_step", "phase_a_ms", "azimuth_ms", "redo_ms", "rays", "raytrace_steps", "disk_hits"), [round(float(v), 4) for v in t.tolist()]))
                            for t in per_rank]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(abi, cfg, p, world, args.gather, {"e2e_note": e2e_note}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(abi.ImageParams), "d2h_bytes_per_step": d2h,
                    "checksum": checksum, "image_check": image_check},
            "gpu_launches": launches_all,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        fence()
        if shm is not None:
            shm.close(dist)
        if peer:
            fence()
            if rank != 0:
                image.close()
            fence()
            if rank == 0:
                image.close()
        dist.destroy_process_group()
    return 0


class HostImageView:
    """what api.trace_image needs of a set of host planes: .out (sim5_image_out), .arrays, [name]"""

    def __init__(self, out, shape):
        self.out, self.shape, self.arrays = out, shape, {}

    def __getitem__(self, k):
        return self.arrays[k]


class SharedHostImage:
    """The output planes of one image in POSIX shared memory, mapped by every rank of the node and page-locked for CUDA in each
    (sim5_host_register): `planes` looks like api.HostPlanes.  Rank 0 creates the segments, the others attach."""

    def __init__(self, api, abi, p, rank, world, dist):
        import numpy as np
        from multiprocessing import shared_memory
        self.api, self.rank = api, rank
        self.segs, self.ptrs, self.unregistered = [], [], False
        n = p.nx * p.ny
        names = [None]
        specs = [(name, ct) for name, bit, ct in abi.PLANES if p.outputs & bit]
        if rank == 0:
            made = []
            try:
                need = sum(n * C.sizeof(ct) for _, ct in specs)
                vfs = os.statvfs("/dev/shm")              # tmpfs hands out pages lazily: a segment that does not fit fails at first touch (SIGBUS), so ask first
                if vfs.f_bavail * vfs.f_frsize < need + (64 << 20):
                    raise OSError("/dev/shm has %d MB free, %d MB needed" % (vfs.f_bavail * vfs.f_frsize >> 20, need >> 20))
                for name, ct in specs:
                    made.append(shared_memory.SharedMemory(create=True, size=n * C.sizeof(ct)))
                names = [[s.name for s in made]]
            except Exception as e:      # e.g. /dev/shm too small: every rank falls back to planes of its own
                for sgm in made:
                    sgm.close(); sgm.unlink()
                made, names = [], [None]
                sys.stderr.write("bench.py: no shared host image (%s); every rank keeps its rows in pinned planes of its own\n" % e)
            self.segs = made
        dist.broadcast_object_list(names, src=0)
        if names[0] is None:
            raise RuntimeError("shared host image unavailable")
        if rank != 0:
            self.segs = [shared_memory.SharedMemory(name=nm) for nm in names[0]]
            try:      # the creator unlinks; an attaching process must not let its resource tracker "clean up" the segment at exit (Python < 3.13)
                from multiprocessing import resource_tracker
                for sgm in self.segs:
                    resource_tracker.unregister(sgm._name, "shared_memory")
            except Exception:
                pass
        npdt = {C.c_double: np.float64, C.c_int32: np.int32, C.c_uint8: np.uint8}
        self.planes = HostImageView(abi.ImageOut(), (p.ny, p.nx))
        for (name, ct), seg in zip(specs, self.segs):
            a = np.ndarray((n,), dtype=npdt[ct], buffer=seg.buf)
            addr = a.ctypes.data
            if api.lib().sim5_host_register(C.c_void_p(addr), C.c_size_t(a.nbytes)) == abi.OK:      # (unregistered memory still works: pageable copies)
                self.ptrs.append(addr)
            else:
                self.unregistered = True
            self.planes.arrays[name] = a
            setattr(self.planes.out, name, addr)
        dist.barrier()

    def close(self, dist):
        for addr in self.ptrs:
            self.api.lib().sim5_host_unregister(C.c_void_p(addr))
        self.planes.arrays = {}
        dist.barrier()
        for s in self.segs:
            try:
                s.close()
            except BufferError:
                pass
        if self.rank == 0:
            for s in self.segs:
                try:
                    s.unlink()
                except Exception:
                    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (default 2: the one the metric is quoted on)")
    ap.add_argument("--size", type=int, default=None, help="image side (default: the BASELINE size of the configuration)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = every rank stores its rows into rank 0's image over NVLink peer memory (default); "
                         "'nccl' = compact planes + torch.distributed gather + re-assembly on rank 0 (A/B)")
    ap.add_argument("--peer-copy", default="auto", choices=["auto", "dma", "stores"],
                    help="N>1 with --gather peer: 'dma' = ranks 1..N-1 trace into local planes and the copy engine moves their row blocks into rank 0's image "
                         "under the next step's kernels; 'stores' = their kernels store straight into rank 0's planes over NVLink; 'auto' (default) = dma from "
                         "8 GPUs on, where the stores of 7 kernels saturate rank 0's NVLink ingress (8 GPUs: 0.915 vs 1.095 ms per step; 4 GPUs: 1.644 vs 1.633, "
                         "2 GPUs: 3.189 vs 3.176 -- profiles/r05o_*, r05n_*)")
    ap.add_argument("--alt-streams", action="store_true", help="A/B: let the calls of the deferred train alternate between two launch streams (SIM5_FLAG_ALT_STREAMS; measured slower)")
    ap.add_argument("--row-major", action="store_true", help="A/B, config 4: rows top to bottom instead of from the middle outwards (SIM5_FLAG_ROW_MAJOR)")
    ap.add_argument("--defer-redo", default="auto", choices=["auto", "on", "off"], help="SIM5_FLAG_DEFER_REDO for the timed train (auto: only with more than one GPU)")
    ap.add_argument("--exact-azimuth", action="store_true", help="A/B: bit-faithful azimuth kernels (SIM5_FLAG_EXACT_AZIMUTH) instead of the tolerance-mode default")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
