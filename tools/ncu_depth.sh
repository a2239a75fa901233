for k in depth_bounds depth_cull depth_seed classify; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 2 -c 1 -o gpurun_out/r01_final_$k -f python tools/prof_step.py C2 3 > gpurun_out/r01_final_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
