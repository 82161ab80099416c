mkdir -p gpurun_out
for G in 128 192 256; do
  NMB_SHELL_G=$G timeout 300 python bench.py --steps 3 --warmup 3 --cpu-rays 0 > gpurun_out/r2j_bench_G$G.txt 2>&1
done
timeout 200 python examples/train_distill.py --steps 12 > gpurun_out/r2j_train_distill.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shell_certificate or full_size" > gpurun_out/r2j_tests.txt 2>&1
for G in 128 192 256; do grep '^{' gpurun_out/r2j_bench_G$G.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read()); bc=j['roofline']['by_class']; print($G, round(j['value']), round(j['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in bc.items()})"; done
grep -v Warn gpurun_out/r2j_train_distill.txt | tail -5; tail -2 gpurun_out/r2j_tests.txt
