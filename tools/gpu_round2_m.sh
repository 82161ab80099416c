mkdir -p gpurun_out
timeout -s KILL 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"
tail -2 gpurun_out/r2m_smoke.txt | cut -c1-400
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv; exit 0; fi
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.txt
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --cpu-rays 0 > gpurun_out/r2m_bench.txt 2>&1
NMB_TC_PROFILE=1 timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --cpu-rays 0 2>&1 | grep tc-prof | tail -12 > gpurun_out/r2m_tcprof.txt
grep -E "passed|failed|rc=" gpurun_out/r2m_tests.txt | tail -3
grep '^{' gpurun_out/r2m_bench.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); bc=j['roofline']['by_class']; print(round(j['value']), round(j['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in bc.items()})"
cut -c1-330 gpurun_out/r2m_tcprof.txt | tail -6
