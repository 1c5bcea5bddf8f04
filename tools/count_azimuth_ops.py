#!/usr/bin/env python
"""Algorithmic work of the tolerance-mode azimuth per disk hit, counted by a host build of the device headers with
-DS5_COUNT_ITERS (one thread): duplication sequences, their steps, R_C series and their steps, per hit of the bench camera.

  g++ -O2 -fPIC -std=c++17 -x c++ -ffp-contract=off -fopenmp -DS5_COUNT_ITERS -I include -shared -o /tmp/libhostsim_count.so tests/hostsim/hostsim.cpp
  python tools/count_azimuth_ops.py /tmp/libhostsim_count.so
"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from sim5_b200 import abi  # noqa: E402

lib = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else "/tmp/libhostsim_count.so")
p = abi.default_params(2, 256)
cnt = (C.c_long * 4)()
out = (C.c_longlong * 4)()
lib.hs_hi_counts(out, 1)
lib.hs_fast_azimuth_coverage(C.byref(p), cnt)        # runs azimuth_fast_rr / _rc on every hit (single-threaded build: OMP_NUM_THREADS=1)
lib.hs_hi_counts(out, 0)
hits = cnt[0] + cnt[2]
res = {"hits": hits, "sequences_per_hit": out[0] / hits, "steps_per_sequence": out[1] / out[0], "rc_series_per_hit": out[2] / hits,
       "rc_steps_per_series": out[3] / max(out[2], 1)}
# flop convention of SURVEY.md 8(d): add/mul 1, fma 2, div 15, sqrt 13
seq_step = 3 * 13 + 2 + 2 + 2 + 1 + 2 + 4 * 2 + 1            # 3 sqrt, syz, lam (fma), ssum, sprod, scale w, x y z updates, convergence test adds
rj_step = 2 + 1 + 3 + 2 + 1                                  # v (fma), pl, the two R_C arguments, acc (fma), pt
rc_series = 3 + 13 + 3 + 6 * 2 + 4                           # A, rsqrt, s, Horner, final fma
rc_step = 13 + 2 + 2 + 3                                     # sqrt, lam, two updates, test
res["flop_convention"] = {"sequence_step": seq_step, "rj_per_step": rj_step, "rc_series": rc_series, "rc_step": rc_step}
print(json.dumps(res, indent=1))
