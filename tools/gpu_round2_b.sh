mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -s -m gpu > gpurun_out/r2b_all_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_all_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/r2b_smoke.txt
timeout 400 python bench.py > gpurun_out/r2b_bench_default.txt 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference.txt 2>&1
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2b_bench_train.txt 2>&1
timeout 300 python bench.py --workload codes256 --steps 2 --warmup 3 --cpu-rays 0 > gpurun_out/r2b_bench_codes256.txt 2>&1
timeout 300 python bench.py --workload scan63_full --steps 2 --warmup 3 --cpu-rays 0 > gpurun_out/r2b_bench_scan63_full.txt 2>&1
timeout 400 python bench.py --workload big --image 2048 --steps 1 --warmup 3 --cpu-rays 0 > gpurun_out/r2b_bench_big2048.txt 2>&1
grep -E "passed|failed" gpurun_out/r2b_all_tests.txt | tail -2; tail -2 gpurun_out/r2b_smoke.txt; for f in default train codes256 scan63_full big2048; do echo $f; tail -c 400 gpurun_out/r2b_bench_$f.txt; echo; done
