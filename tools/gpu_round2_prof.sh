mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -c 8 -o gpurun_out/r2_mlp_f16 python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2_ncu_mlp.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:knn_rays_kernel|bound_rays_kernel|knn_lists_kernel" -c 4 -o gpurun_out/r2_walk python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2_ncu_walk.log 2>&1
tail -2 gpurun_out/r2_ncu_mlp.log gpurun_out/r2_ncu_walk.log; ls -la gpurun_out/*.ncu-rep
