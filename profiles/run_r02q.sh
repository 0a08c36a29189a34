#!/bin/bash
# Round-2 validation on one B200 (under gpurun): full GPU tests, both bench arms, launch list, full ncu captures of the hot
# kernels, compute-sanitizer memcheck / racecheck on small shapes.  Outputs -> gpurun_out/.   bash profiles/run_r02q.sh [tag]
TAG=${1:-r02q}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cut -c1-300 gpurun_out/${TAG}_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
python bench.py --config cfg2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg2.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.json 2> gpurun_out/${TAG}_launches.err
for k in gram_tcm_kernel spike_slab_fast_kernel pg_pick_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page details > gpurun_out/${TAG}_ncu_$k.txt 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$k.csv 2>&1
  rm -f gpurun_out/${TAG}_prof_$k.ncu-rep
done
# compute-sanitizer on small shapes: tcgen05 Gram (resident, multicast and streaming), cluster kernels (generate, scan)
SEL='gram_tc_integer_sums or spike_slab_golden or all_inactive or (gram_tc_time_slabs)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "generate_replays and 12" >> gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck generate rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "spike_slab_golden or all_inactive" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/${TAG}_sanitizer_racecheck.log
ls -la gpurun_out | tail -30
