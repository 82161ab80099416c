mkdir -p gpurun_out
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
export -f run; export N
timeout 300 bash -c 'run --steps 3 --warmup 3' > gpurun_out/r2g_bench_${N}gpu_frames.txt 2>&1
timeout 300 bash -c 'run --steps 3 --warmup 3 --frames-per-step 1' > gpurun_out/r2g_bench_${N}gpu_single.txt 2>&1
timeout 300 bash -c 'run --workload train --steps 10 --warmup 3' > gpurun_out/r2g_bench_${N}gpu_train.txt 2>&1
timeout 500 bash -c 'run --workload big --steps 1 --warmup 1 --frames-per-step 1 --cpu-rays 0' > gpurun_out/r2g_bench_${N}gpu_big4096.txt 2>&1
for f in frames single train big4096; do echo $f; grep '^{' gpurun_out/r2g_bench_${N}gpu_$f.txt | cut -c1-260; tail -2 gpurun_out/r2g_bench_${N}gpu_$f.txt | cut -c1-200; done
