#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench (both arms), peaks, launch list, one full capture of each hot
# kernel.  Outputs -> gpurun_out/.   Usage: bash profiles/run_profile.sh [tag]
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cat gpurun_out/${TAG}_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cat gpurun_out/${TAG}_bench_reference.json
python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
python profiles/profile_sweep.py --int8 > gpurun_out/${TAG}_int8_peak.json 2>&1
cat gpurun_out/${TAG}_int8_peak.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/profile_sweep.py --sweeps 2 > gpurun_out/${TAG}_launches.log 2>&1
for k in ${KERNELS:-gram_tc_kernel pg_pick_kernel pg_ig_small_kernel spike_slab activation_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k \
      python profiles/profile_sweep.py --sweeps 2 > gpurun_out/${TAG}_prof_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page details > gpurun_out/${TAG}_ncu_$k.txt 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$k.csv 2>&1
  # gpurun_out/ is capped at 64 MiB: keep a report only when it is small
  sz=$(stat -c %s gpurun_out/${TAG}_prof_$k.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/${TAG}_prof_$k.ncu-rep; fi
done
ls -la gpurun_out
