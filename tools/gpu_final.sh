#!/bin/bash
# Final 1-GPU session of a round: all GPU tests, bench of the five configs (both arms), ncu launch list and full captures of every kernel family.
# usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
for c in 1 3 4 5; do
  K=10; [ $c = 5 ] && K=3; [ $c = 4 ] && K=3
  python bench.py --config $c --steps $K --warmup 3 > gpurun_out/${TAG}_bench_cfg$c.json 2>> gpurun_out/${TAG}_bench.err
  python bench.py --impl reference --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_cfg$c.json 2>> gpurun_out/${TAG}_bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth_fast" -s 12 -c 3 -f -o gpurun_out/${TAG}_prof_eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_step \
    python bench.py --config 4 --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_surface \
    python tools/surface_bench.py > gpurun_out/${TAG}_ncu_surface.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_histogram" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_hist \
    python tools/perf_sweep.py > gpurun_out/${TAG}_ncu_hist.log 2>&1
timeout 300 python tools/surface_bench.py > gpurun_out/${TAG}_surface.json 2> gpurun_out/${TAG}_surface.err
timeout 300 python tools/spectrum_bench.py > gpurun_out/${TAG}_spectrum.json 2> gpurun_out/${TAG}_spectrum.err
python tools/d2h_roofline.py > gpurun_out/${TAG}_d2h_n1.json 2> gpurun_out/${TAG}_d2h.err
for f in gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench_cfg*.json; do cut -c1-260 $f; done
tail -n 5 gpurun_out/${TAG}_bench.err
