#!/usr/bin/env python
"""Micro-benchmark: how fast do the device routines run in isolation (small code, no i-cache pressure)?"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sim5_b200 import api
api.init(0)
L = api.lib()
L.sim5_micro_bench.restype = C.c_double
L.sim5_micro_bench.argtypes = [C.c_int, C.c_int64, C.c_int]
n, reps = 148 * 128 * 16 * 8, 16
names = ["rf", "rj", "rc", "sncndn", "sincos", "log", "atan2", "pow_third", "4 div", "4 sqrt", "rj complete"]
for w, name in enumerate(names):
    ms = L.sim5_micro_bench(w, n, reps)
    calls = n * reps
    print("%-12s %.3f ms for %d calls -> %.3e calls/s, %.2f ns*SM/call" % (name, ms, calls, calls / ms * 1e3, ms * 1e6 / calls * 148))
