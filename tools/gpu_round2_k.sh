mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "full_size or config5 or teacher_forced or frame_parity or kwargs or detailed" > gpurun_out/r2k_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.txt
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-rays 0 > gpurun_out/r2k_bench.txt 2>&1
grep -E "passed|failed|rc=" gpurun_out/r2k_tests.txt | tail -3
grep '^{' gpurun_out/r2k_bench.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); bc=j['roofline']['by_class']; print(round(j['value']), round(j['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in bc.items()})"
