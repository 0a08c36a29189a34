python profiles/probe_overlap_timeline.py > gpurun_out/r02z_timeline_scan_first.log 2>&1; tail -25 gpurun_out/r02z_timeline_scan_first.log
echo ---- aug_first
PYGLM_OVERLAP_ORDER=aug_first python profiles/probe_overlap_timeline.py > gpurun_out/r02z_timeline_aug_first.log 2>&1; tail -25 gpurun_out/r02z_timeline_aug_first.log
echo ---- aug_first, 32 connections
CUDA_DEVICE_MAX_CONNECTIONS=32 PYGLM_OVERLAP_ORDER=aug_first python profiles/probe_overlap_timeline.py > gpurun_out/r02z_timeline_aug_first_c32.log 2>&1; tail -25 gpurun_out/r02z_timeline_aug_first_c32.log
