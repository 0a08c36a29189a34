#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full capture of each hot kernel.  Outputs -> gpurun_out/.
set -x
mkdir -p gpurun_out
python profiles/profile_sweep.py --dgemm > gpurun_out/fp64_peak.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python profiles/profile_sweep.py --sweeps 1 > gpurun_out/launches.log 2>&1
for k in gram_kernel pg_draw_kernel spike_slab_kernel activation_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
      python profiles/profile_sweep.py --sweeps 1 > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
