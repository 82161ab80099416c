mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -s -m gpu > gpurun_out/r2z_all_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2z_all_tests.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/r2z_smoke.txt
timeout -s KILL 400 python bench.py > gpurun_out/r2z_bench_default.txt 2>&1
timeout -s KILL 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2z_bench_train.txt 2>&1
timeout -s KILL 300 python bench.py --workload codes256 --steps 2 --warmup 3 --cpu-rays 0 > gpurun_out/r2z_bench_codes256.txt 2>&1
timeout -s KILL 300 python bench.py --workload scan63_full --steps 2 --warmup 3 --cpu-rays 0 > gpurun_out/r2z_bench_scan63_full.txt 2>&1
timeout -s KILL 400 python bench.py --workload big --image 2048 --steps 1 --warmup 3 --cpu-rays 0 > gpurun_out/r2z_bench_big2048.txt 2>&1
grep -E "passed|failed|rc=" gpurun_out/r2z_all_tests.txt | tail -2; tail -2 gpurun_out/r2z_smoke.txt | cut -c1-300
for f in default train codes256 scan63_full big2048; do echo $f; grep '^{' gpurun_out/r2z_bench_$f.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],1), 'e2e', round(j['e2e']['value']), (j.get('cpu_baseline') or {}).get('value'))
bc=(j.get('roofline') or {}).get('by_class') or {}
print({k:round(v['ms_per_step'],1) for k,v in bc.items()})"; done
