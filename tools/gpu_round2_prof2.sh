mkdir -p gpurun_out
# 2. full captures (reports stay in /tmp: too large to bring back), exported as raw CSV
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -c 8 -o /tmp/r2f_mlp python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2f_ncu_mlp.log 2>&1
ncu -i /tmp/r2f_mlp.ncu-rep --page raw --csv > gpurun_out/r2f_mlp_raw.csv 2>/dev/null
timeout 500 ncu --set full --clock-control none -k "regex:knn_rays_kernel|bound_dir_kernel|knn_lists_kernel" -c 8 -o /tmp/r2f_walk python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2f_ncu_walk.log 2>&1
ncu -i /tmp/r2f_walk.ncu-rep --page raw --csv > gpurun_out/r2f_walk_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | grep r2f
