TAG=r01m
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k pg 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cat gpurun_out/${TAG}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.json 2> gpurun_out/${TAG}_bench_under_ncu.err
for k in pg_pick_kernel pg_ig_small_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k python profiles/profile_sweep.py --sweeps 2 > gpurun_out/${TAG}_prof_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page details > gpurun_out/${TAG}_ncu_$k.txt 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_$k.csv 2>&1
  rm -f gpurun_out/${TAG}_prof_$k.ncu-rep
done
ls gpurun_out | grep r01m
