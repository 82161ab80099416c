mkdir -p gpurun_out
for e in 0 1; do
NMB_TC_EXP=$e NMB_TC_PROFILE=1 timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --cpu-rays 0 2>&1 | grep tc-prof | tail -5 | cut -c1-330 > gpurun_out/r2n_tcprof_$e.txt
NMB_TC_EXP=$e timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --cpu-rays 0 2>/dev/null | grep '^{' | python -c "
import json,sys
j=json.loads(sys.stdin.read()); bc=j['roofline']['by_class']; print(round(j['value']), round(j['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in bc.items()})"
cat gpurun_out/r2n_tcprof_$e.txt
done
