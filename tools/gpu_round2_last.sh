mkdir -p gpurun_out
timeout -s KILL 50 python -m pytest tests -q -x -m gpu -k "frame_parity or perturb" > gpurun_out/r2v_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2v_tests.txt
tail -3 gpurun_out/r2v_tests.txt
