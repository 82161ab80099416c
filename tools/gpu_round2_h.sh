mkdir -p gpurun_out
timeout 300 python bench.py --simulate-world 8 --steps 3 --warmup 2 > gpurun_out/r2h_sim8.txt 2>&1
timeout 300 python bench.py --simulate-world 4 --steps 3 --warmup 2 > gpurun_out/r2h_sim4.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_full_size_properties or perturb" -s > gpurun_out/r2h_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.txt
grep '^{' gpurun_out/r2h_sim8.txt; grep '^{' gpurun_out/r2h_sim4.txt; grep -E "passed|failed|perturb:" gpurun_out/r2h_tests.txt
