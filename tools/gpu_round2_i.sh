mkdir -p gpurun_out
timeout 300 python bench.py --simulate-world 8 --steps 3 --warmup 2 > gpurun_out/r2i_sim8.txt 2>&1
timeout 200 python tools/gemm_probe.py > gpurun_out/r2i_gemm_probe.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "perturb" -s > gpurun_out/r2i_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.txt
grep '^{' gpurun_out/r2i_sim8.txt; grep -v Warn gpurun_out/r2i_gemm_probe.txt | tail -5; grep -E "passed|failed|perturb:" gpurun_out/r2i_tests.txt
