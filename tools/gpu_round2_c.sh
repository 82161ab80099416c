mkdir -p gpurun_out
timeout 200 python tools/smoke_diag.py > gpurun_out/r2c_smoke_diag.txt 2>&1
NMB_TC_PROFILE=1 timeout 200 python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2c_tcprof.txt 2>&1
grep -v Warn gpurun_out/r2c_smoke_diag.txt | tail -30; grep tc-prof gpurun_out/r2c_tcprof.txt | head -12
