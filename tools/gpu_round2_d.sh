mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "test_field_vs_oracle or test_render_teacher_forced or test_full_size_properties or test_config3 or test_field_vs_reference_golden" > gpurun_out/r2d_mlp_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2d_mlp_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/r2d_smoke.txt
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-rays 0 > gpurun_out/r2d_bench_f16.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-rays 0 --engine tcgen05 > gpurun_out/r2d_bench_tf32.txt 2>&1
NMB_TC_PROFILE=1 timeout 200 python tools/prof_driver.py 200000 tcgen05_f16 > gpurun_out/r2d_tcprof.txt 2>&1
grep -E "passed|failed|rc=" gpurun_out/r2d_mlp_tests.txt | tail -3; tail -2 gpurun_out/r2d_smoke.txt; tail -c 300 gpurun_out/r2d_bench_f16.txt; grep tc-prof gpurun_out/r2d_tcprof.txt | head -3
