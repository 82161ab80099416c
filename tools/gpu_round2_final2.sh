mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "teacher_forced" > gpurun_out/r2y_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2y_tests.txt
timeout -s KILL 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2y_bench_train.txt 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --cpu-rays 0 > gpurun_out/r2f_launch_bench.log 2>&1
grep -E "passed|failed|rc=|per-sample" gpurun_out/r2y_tests.txt | cut -c1-250 | tail -6
grep '^{' gpurun_out/r2y_bench_train.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['clocks'])"
wc -l gpurun_out/r2f_launches.csv
