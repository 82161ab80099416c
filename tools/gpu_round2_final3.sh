mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -q -m gpu -k "repack or train or field_vs_oracle or teacher or neus" > gpurun_out/r2x_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2x_tests.txt
for i in 1 2; do timeout -s KILL 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2x_bench_train_$i.txt 2>&1; done
timeout -s KILL 200 python bench.py --steps 2 --warmup 2 --cpu-rays 0 > gpurun_out/r2x_bench.txt 2>&1
grep -E "passed|failed|rc=" gpurun_out/r2x_tests.txt | tail -3
for i in 1 2; do grep '^{' gpurun_out/r2x_bench_train_$i.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['clocks'])"; done
grep '^{' gpurun_out/r2x_bench.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],1))"
