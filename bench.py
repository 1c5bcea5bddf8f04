#!/usr/bin/env python
"""bench.py -- headline benchmark of the SIM5 photon hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation (host cores)

Workload (BASELINE.json configs[1]): thin-disk image 4096x4096, a=0.998, i=75 deg, outputs r, phi, g,
F*g^4 and the status byte; one STEP = one full image.  At N>1 the image rows are dealt to the ranks
in interleaved 32-row blocks (sim5_b200/dist.py), every rank traces its rows, and the planes are gathered
on rank 0 with NCCL at the end of the step (strong scaling: total work per step is fixed).

Metric: geodesics (rays) per second.
  value : device-resident -- planes stay in HBM (gathered to rank 0's HBM at N>1); CUDA events on the launch stream
  e2e   : through the public host API (sim5_trace_image with pinned HOST planes): the timed region holds
          the parameter upload and the device->host copy of every plane, every step
Roofline: FP64 non-tensor pipe (no dense contraction, 44 B written and 0 B read per ray -> compute bound);
  achieved = F_alg (16.8 kflop per ray with azimuth) * rays / kernel time,
  peak = live DFMA-chain microbenchmark (MEASURED_PEAKS.json has no FP64 entry).
"""
import argparse
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
F_ALG_PHI = 16.8e3        # flop per ray, (r, phi, g, F): SURVEY.md 8(d) non-redundant algorithm, / = 15, sqrt = 13 flop
F_ALG_TRACE = 3.7e3       # of which phase A (roots, crossing, r, g, F): 3 rf + 1 sncndn + roots + glue, per ray
F_ALG_AZ_RR = 13.1e3      # phase B per RR disk hit, the reference's non-redundant algorithm (bit-faithful kernels): 3 rf + 6 rj + 2 sncndn + glue
F_ALG_AZ_FAST = 2.9e3     # phase B per RR disk hit, tolerance-mode algorithm (the default): 3 shared duplication sequences x 3.80 steps,
                          # 18.4 R_C series, 5 R_J + 3 R_F tails, one AGM complete Pi, glue -- counted by tools/count_azimuth_ops.py on a host
                          # build with -DS5_COUNT_ITERS; 3.3e3 before the complete Pi went from duplication to AGM
BYTES_PER_RAY = 4 * 8 + 1
SAMPLE_N = 1024           # CPU sample: the same camera at 1024x1024 (1/16 of the rays of the 4096^2 image)


def workload_params(abi, n=None):
    p = abi.default_params(2)
    if n:
        p.nx = p.ny = n
    return p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref, the unmodified sources + OpenMP pixel loop) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import harness as H
    from sim5_b200 import abi
    p = workload_params(abi, SAMPLE_N)
    # all the host threads this process may use -- stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if H.have_ref():
        kind, run = "reference", (lambda q: H.run_ref(q, nthreads=cores))
    elif H.have_oracle():
        kind, run = "port", (lambda q: H.run_oracle(q, nthreads=cores))
    else:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libsim5ref.so nor oracle/libsim5oracle.so is built"}))
        return 0
    for _ in range(args.warmup):
        run(p)
    t = 0.0
    for _ in range(args.steps):
        _, st, dt = run(p)
        t += dt
    rays = p.nx * p.ny * args.steps
    v = rays / t
    sample = "same camera (a=0.998, i=75deg, rmax=r_ms+20, r/phi/g/flux/status) at %dx%d = 1/16 of the rays per step" % (SAMPLE_N, SAMPLE_N)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args.gpus, 4096, args.gather),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def config_dict(n_gpus, n, gather="peer"):
    return {"workload": "BASELINE configs[1]: Novikov-Thorne thin-disk image %dx%d, a=0.998, i=75deg, rmax=r_ms+20, "
                        "outputs r/phi/g/F*g^4/status, crossing orders 0-1" % (n, n),
            "rays_per_step": n * n,
            "parallelism": ("one GPU, whole image" if n_gpus == 1 else
                            "rows interleaved over %d GPUs in 32-row blocks; %s" % (n_gpus, "every rank stores its rows into rank 0's image planes over NVLink peer memory (CUDA IPC), barrier at the end"
                                                                                    if gather == "peer" else "NCCL gather of compact planes to rank 0 + re-assembly")),
            "l2": "no input reads (rays are generated from the pixel index); %d MB of output planes per step > 126 MB L2" % (n * n * BYTES_PER_RAY // 1000000),
            "e2e_note": "each rank copies its own rows to its own pinned host planes" if n_gpus > 1 else "pinned host planes"}


def run_ours(args):
    # NCCL prints its version / debug lines to stdout by default; stdout carries exactly one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import ctypes as C
    from sim5_b200 import abi, api, dist as sdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    api.init(local)
    L = api.lib()
    stream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: the library launches on it, torch events time it
    torch.cuda.set_stream(stream)
    api.check(L.sim5_set_stream(C.c_void_p(stream.cuda_stream)), "sim5_set_stream")

    n = args.size
    p = workload_params(abi, n)
    p.device = local
    rows = sdist.apply_split(p, rank, world) if world > 1 else n
    p.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC | (abi.FLAG_EXACT_AZIMUTH if args.exact_azimuth else 0)
    # a train of images: the redo wave of step k can run beside the kernels of step k+1 (SIM5_FLAG_DEFER_REDO).  It pays when the wave is a
    # large part of the step (multi-GPU split: 8 GPUs 1.24 -> 1.12 ms); on one GPU it is neutral to slightly negative (the wave's 32 CTAs
    # take SMs from the 2.7 ms azimuth kernel: 7.30 vs 7.32-7.35 ms), so by default it is used for N > 1 only
    defer = (world > 1) if args.defer_redo == "auto" else (args.defer_redo == "on")
    if defer and not args.no_defer_redo:
        p.flags |= abi.FLAG_DEFER_REDO
    names = ("r", "phi", "g", "flux")
    peer = world > 1 and args.gather == "peer"
    st = abi.TraceStats()
    gath = None
    if peer:
        # the image lives in rank 0's HBM; every rank maps those planes (CUDA IPC, peer access over NVLink/NVSwitch) and its
        # kernels store their interleaved row blocks straight into the final image: the transfer rides under the FP64 work,
        # there is no gather and no re-assembly pass
        p.flags |= abi.FLAG_FULL_INDEX
        if rank == 0:
            image = api.DevicePlanes(p)
            blob = [image.handles()]
        else:
            image, blob = None, [None]
        dist.broadcast_object_list(blob, src=0)
        if rank != 0:
            image = api.DevicePlanes.from_handles(p, blob[0])
        out = image.out
        loc = None
    else:
        loc = {k: torch.empty((rows, n), dtype=torch.float64, device=dev) for k in names}
        loc["status"] = torch.empty((rows, n), dtype=torch.uint8, device=dev)
        out = abi.ImageOut()
        for k, t in loc.items():
            setattr(out, k, t.data_ptr())
        if world > 1 and rank == 0:
            gath = {k: [torch.empty_like(t) for _ in range(world)] for k, t in loc.items()}
    kev = []
    phase_ms = [0.0, 0.0, 0.0]      # summed device time of k_trace_eqplane, k_azimuth<RR>, k_azimuth<RC> over the timed steps
    phase_items = [0, 0]            # RR / RC disk hits integrated by the azimuth kernels (per step)
    launches = [0]

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        api.check(L.sim5_trace_image(C.byref(p), C.byref(out), C.byref(st)), "sim5_trace_image")
        if timed:
            e1.record()
            kev.append((e0, e1))
        if world > 1 and not peer:
            api.check(L.sim5_join(), "sim5_join")      # the gather reads phi: this step's redo passes must be behind it on the stream
            full = None
            for k, t in loc.items():
                dist.gather(t, gath[k] if rank == 0 else None, dst=0)
                if rank == 0:
                    full = sdist.assemble(gath[k], world)
            return full
        return None

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf = api.fp64_peak_tflops(local, 8192) if rank == 0 else None
    api.check(L.sim5_set_stream(C.c_void_p(stream.cuda_stream)), "sim5_set_stream")

    for _ in range(max(args.warmup, 3)):
        step(False)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    fence()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(True)                 # SIM5_FLAG_ASYNC: the K steps are enqueued back to back, no host sync in between
    api.check(L.sim5_join(), "sim5_join")      # the launch stream waits for the last steps' deferred redo passes: they are inside the timed region
    ev1.record()
    fence()
    # per-kernel CUDA events the library recorded on the launch stream for each of the timed calls (it keeps the last 63),
    # read back after the timed region so that the steps above run without a synchronisation per step
    nread = min(args.steps, 63)
    for back in range(nread):
        pm, nk = api.phase_history(back)
        for i, v in enumerate(pm):
            phase_ms[i] += v * (args.steps / nread)
        launches[0] += nk
    launches[0] = launches[0] * args.steps // nread
    _, items, _ = api.last_phase_ms()
    phase_items[0], phase_items[1] = items
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / len(kev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kernel_ms = float(kms.item())
    rays_step = n * n
    value = rays_step * args.steps / (total_ms * 1e-3)

    # ---- e2e: public host API, pinned host planes, H2D of the parameters + D2H of every plane inside the timed region
    api.check(L.sim5_set_stream(None), "sim5_set_stream")
    ph = workload_params(abi, n)
    ph.device = local
    if args.exact_azimuth:
        ph.flags |= abi.FLAG_EXACT_AZIMUTH
    if world > 1:
        sdist.apply_split(ph, rank, world)
    hp = api.HostPlanes(ph, pinned=True)
    for a in hp.arrays.values():
        a[...] = 0                      # at N>1 a rank fills only its own rows
    for _ in range(2):
        api.trace_image(ph, hp)
    fence()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, hst = api.trace_image(ph, hp)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    fence()
    clocks = sampler.stop() if sampler else None
    e2e_value = rays_step * args.steps / float(e2e_s.item())
    import numpy as _np
    cs = torch.tensor([float(hp["g"].sum()), float(_np.nansum(_np.abs(hp["phi"])))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)      # every rank's own rows -> checksums of the whole image
    checksum, checksum_phi = float(cs[0].item()), float(cs[1].item())
    image_check = None
    if peer and rank == 0:
        # the image the ranks assembled in rank 0's HBM through peer stores, against the host-API result of all ranks
        dev_sum = float(image.to_host("g").sum())
        dev_phi = float(_np.nansum(_np.abs(image.to_host("phi"))))      # phi is completed by the (deferred) redo passes of the timed train
        image_check = {"sum_g_device_image": dev_sum, "sum_g_host_api": checksum, "rel_diff": abs(dev_sum - checksum) / max(abs(checksum), 1e-300),
                       "sum_abs_phi_device_image": dev_phi, "sum_abs_phi_host_api": checksum_phi,
                       "rel_diff_phi": abs(dev_phi - checksum_phi) / max(abs(checksum_phi), 1e-300)}
        assert image_check["rel_diff"] < 1e-12 and image_check["rel_diff_phi"] < 1e-12, image_check

    if rank == 0:
        # CPU baseline beside it (N=1 only): the unmodified reference on the host cores, bounded sample
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                import harness as H
                ps = workload_params(abi, SAMPLE_N)
                if H.have_ref():
                    kind, run, cores = "reference", H.run_ref, H.load_ref().ref_max_threads()
                else:
                    kind, run, cores = "port", H.run_oracle, os.cpu_count()
                run(ps)
                _, _, dt = run(ps)
                cpu = {"value": ps.nx * ps.ny / dt, "unit": UNIT, "cores": cores, "kind": kind,
                       "sample": "same camera at %dx%d (1/16 of the rays), all host threads, 1 timed pass after 1 warm-up" % (SAMPLE_N, SAMPLE_N)}
            except Exception as e:  # the checker is optional for the benchmark itself
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        f_rr = F_ALG_AZ_RR if args.exact_azimuth else F_ALG_AZ_FAST
        flop_step = F_ALG_TRACE * (rays_step / world) + f_rr * (phase_items[0] + (0 if args.exact_azimuth else phase_items[1])) \
            + (F_ALG_AZ_RR * phase_items[1] if args.exact_azimuth else 0)      # work of the algorithm actually run
        achieved_step = flop_step / (kernel_ms * 1e-3) / 1e12
        ph = [v / args.steps for v in phase_ms]
        dom = max(range(3), key=lambda i: ph[i])
        knames = ("k_trace_eqplane<DEFER>", "k_azimuth<RR>", "k_azimuth<RC>") if args.exact_azimuth else ("k_trace_eqplane<DEFER>", "k_azimuth_fast", "k_azimuth<RR|RC> redo")
        if dom == 1:
            units, per_unit, what = (phase_items[0], f_rr, "RR disk hits") if args.exact_azimuth else (phase_items[0] + phase_items[1], f_rr, "RR+RC disk hits")
        elif dom == 0:
            units, per_unit, what = rays_step // world, F_ALG_TRACE, "rays"
        else:
            units, per_unit, what = phase_items[1], F_ALG_AZ_RR, "RC disk hits"
        achieved = per_unit * units / (ph[dom] * 1e-3) / 1e12 if ph[dom] > 0 else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                traffic = json.load(fh).get(knames[dom])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(world, n, args.gather),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(abi.ImageParams),
                    "d2h_bytes_per_step": (rays_step // world) * BYTES_PER_RAY, "checksum_g": checksum, "peer_image_check": image_check},
            "gpu_launches": launches[0] * world,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if (peak_tf and achieved) else None, "traffic": traffic,
                         "peak_source": "live FP64 DFMA-chain microbenchmark (sim5_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 entry",
                         "kernel": knames[dom], "kernel_ms": ph[dom], "units_per_launch": units, "unit_kind": what, "flop_per_unit": per_unit,
                         "kernels_ms": dict(zip(knames, ph)), "azimuth_items": {"rr": phase_items[0], "rc": phase_items[1]},
                         "step": {"achieved": achieved_step, "frac": achieved_step / peak_tf if peak_tf else None,
                                  "flop_per_ray": flop_step / (rays_step / world), "kernels_ms_total": kernel_ms,
                                  "reference_algorithm_equiv_tflops": F_ALG_PHI * (rays_step / world) / (kernel_ms * 1e-3) / 1e12,
                                  "note": "achieved counts the flops of the algorithm actually run (tolerance-mode azimuth: 2.9 kflop per RR hit); "
                                          "reference_algorithm_equiv uses SURVEY.md 8(d)'s 16.8 kflop per ray of the reference's non-redundant algorithm"},
                         "hbm_written_bytes_per_launch": (rays_step // world) * BYTES_PER_RAY},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        if peer:
            fence()
            if rank != 0:
                image.close()
            fence()
            if rank == 0:
                image.close()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="image side (default: the BASELINE 4096)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = every rank stores its rows into rank 0's image over NVLink peer memory (default); "
                         "'nccl' = compact planes + torch.distributed gather + re-assembly on rank 0 (A/B)")
    ap.add_argument("--defer-redo", default="auto", choices=["auto", "on", "off"], help="SIM5_FLAG_DEFER_REDO for the timed train (auto: only with more than one GPU)")
    ap.add_argument("--no-defer-redo", action="store_true", help="A/B: join the azimuth redo passes inside every step instead of letting them run beside the next step's tracing kernel")
    ap.add_argument("--exact-azimuth", action="store_true", help="A/B: bit-faithful azimuth kernels (SIM5_FLAG_EXACT_AZIMUTH) instead of the tolerance-mode default")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
