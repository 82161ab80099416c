mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_active"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -c 8 -o /tmp/r2_mlp_f16 python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2_ncu_mlp.log 2>&1
ncu -i /tmp/r2_mlp_f16.ncu-rep --page raw --csv > gpurun_out/r2_mlp_f16_raw.csv 2>/dev/null
ncu -i /tmp/r2_mlp_f16.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r2_mlp_f16_src_geo.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k "regex:knn_rays_kernel|bound_rays_kernel|knn_lists_kernel" -c 4 -o /tmp/r2_walk python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2_ncu_walk.log 2>&1
ncu -i /tmp/r2_walk.ncu-rep --page raw --csv > gpurun_out/r2_walk_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
