mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); run 2 --steps 3 --warmup 3" > gpurun_out/r2e_bench_2gpu_frames.txt 2>&1
timeout 300 bash -c "$(declare -f run); run 2 --steps 3 --warmup 3 --frames-per-step 1" > gpurun_out/r2e_bench_2gpu_single.txt 2>&1
timeout 300 bash -c "$(declare -f run); run 2 --workload train --steps 10 --warmup 3" > gpurun_out/r2e_bench_2gpu_train.txt 2>&1
timeout 200 bash -c "$(declare -f run); run 2 --impl reference --steps 1 --warmup 1" > gpurun_out/r2e_bench_2gpu_reference.txt 2>&1
for f in frames single train reference; do echo $f; grep '^{' gpurun_out/r2e_bench_2gpu_$f.txt | cut -c1-330; done
