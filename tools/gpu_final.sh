#!/bin/bash
# Final 1-GPU session of a round: all GPU tests, bench of the five configs, ncu launch list and full captures of every kernel family,
# condensed ON THE BOX (gpurun copies back at most 64 MiB: the .ncu-rep files stay there).  usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
for c in 1 3 4 5; do
  K=10; [ $c = 5 ] && K=3; [ $c = 4 ] && K=3
  python bench.py --config $c --steps $K --warmup 3 > gpurun_out/${TAG}_bench_cfg$c.json 2>> gpurun_out/${TAG}_bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
R=/tmp/${TAG}_rep; mkdir -p $R
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth_fast" -s 12 -c 3 -f -o $R/eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $R/eqplane.ncu-rep gpurun_out/${TAG}_ncu_summary.csv > /dev/null 2>&1
ncu -i $R/eqplane.ncu-rep --page source --csv --kernel-name regex:k_trace_eqplane 2>/dev/null | gzip > gpurun_out/${TAG}_src_eqplane.csv.gz
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes" -s 2 -c 1 -f -o $R/step \
    python bench.py --config 4 --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_step.log 2>&1
python tools/ncu_summary.py $R/step.ncu-rep gpurun_out/${TAG}_ncu_summary_stepwise.csv > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes" -s 2 -c 1 -f -o $R/surface \
    python tools/surface_bench.py > gpurun_out/${TAG}_ncu_surface.log 2>&1
python tools/ncu_summary.py $R/surface.ncu-rep gpurun_out/${TAG}_ncu_summary_surface.csv > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_histogram" -s 1 -c 1 -f -o $R/hist \
    python bench.py --config 5 --size 256 --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_hist.log 2>&1
python tools/ncu_summary.py $R/hist.ncu-rep gpurun_out/${TAG}_ncu_summary_histogram.csv > /dev/null 2>&1
timeout 300 python tools/surface_bench.py > gpurun_out/${TAG}_surface.json 2> gpurun_out/${TAG}_surface.err
timeout 300 python tools/spectrum_bench.py > gpurun_out/${TAG}_spectrum.json 2> gpurun_out/${TAG}_spectrum.err
python tools/d2h_roofline.py > gpurun_out/${TAG}_d2h_n1.json 2> gpurun_out/${TAG}_d2h.err
for f in gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench_cfg*.json; do cut -c1-200 $f; done
head -c 600 gpurun_out/${TAG}_ncu_summary_surface.csv; ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
tail -n 3 gpurun_out/${TAG}_bench.err
