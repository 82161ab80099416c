mkdir -p gpurun_out
T="timeout 200"
$T python -m pytest tests/test_train_ops.py -x -q -s -m gpu > gpurun_out/r2a_train_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_train_tests.txt
$T python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or mesh_distance" > gpurun_out/r2a_knn_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_knn_tests.txt
NMB_KNN_NO_DIR=1 $T python tools/knn_ab.py dump nodir > gpurun_out/r2a_ab_nodir.txt 2>&1
$T python tools/knn_ab.py dump dir > gpurun_out/r2a_ab_dir.txt 2>&1
python tools/knn_ab.py compare nodir dir > gpurun_out/r2a_ab_cmp.txt 2>&1
rm -f gpurun_out/knn_ab_*.pt
NMB_KNN_NO_DIR=1 $T python bench.py --steps 2 --warmup 3 --cpu-rays 0 --engine tcgen05_f16 > gpurun_out/r2a_bench_nodir.txt 2>&1
$T python bench.py --steps 2 --warmup 3 --cpu-rays 0 --engine tcgen05_f16 > gpurun_out/r2a_bench_dir.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_texture_edit.py tests/test_train_path.py -q -s -m gpu > gpurun_out/r2a_all_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_all_tests.txt
tail -4 gpurun_out/r2a_train_tests.txt; tail -2 gpurun_out/r2a_knn_tests.txt; grep AB gpurun_out/r2a_ab_cmp.txt; tail -3 gpurun_out/r2a_all_tests.txt
