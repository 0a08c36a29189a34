TAG=r01n
python examples/block_network.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_example_block_network.log
for k in pg_pick_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k python profiles/profile_sweep.py --sweeps 2 > gpurun_out/${TAG}_prof_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page details > gpurun_out/${TAG}_ncu_$k.txt 2>&1
  rm -f gpurun_out/${TAG}_prof_$k.ncu-rep
done
grep -E "Duration|Issue Slots Busy|Avg. Active Threads" gpurun_out/${TAG}_ncu_pg_pick_kernel.txt
