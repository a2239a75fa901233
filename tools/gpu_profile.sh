#!/bin/bash
# full profiling round of the current code: launch list + `ncu --set full` of every kernel of a C2 frame (isotropic) and
# of the two anisotropic march kernels.  usage: tools/gpu_profile.sh r02k
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python tools/prof_step.py C2 3 > gpurun_out/${TAG}_prof.log 2>&1
tail -1 gpurun_out/${TAG}_prof.log | cut -c1-200
for k in aabb_params key_count scan_flags scatter cell_order depth_clear depth_seed depth_gate depth_bounds depth_coarse depth_cull depth_splat classify march_first march_long; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f python tools/prof_step.py C2 3 > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
for k in march_first march_long; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 1 -c 1 -o gpurun_out/${TAG}_aniso_$k -f python tools/prof_step_aniso.py C2 2 > gpurun_out/${TAG}_ncu_aniso_$k.log 2>&1
done
ls gpurun_out/${TAG}_*.ncu-rep | wc -l; du -sh gpurun_out | tail -1
