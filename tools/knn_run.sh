mkdir -p gpurun_out
T="timeout 150"
$T python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or mesh_distance or bounded_near_far" > gpurun_out/r2_knn_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_knn_tests.txt
NMB_KNN_LEGACY=1 $T python tools/knn_ab.py dump legacy > gpurun_out/r2_ab_legacy.txt 2>&1
NMB_KNN_NO_DIR=1 $T python tools/knn_ab.py dump nodir > gpurun_out/r2_ab_nodir.txt 2>&1
$T python tools/knn_ab.py dump coop > gpurun_out/r2_ab_coop.txt 2>&1
python tools/knn_ab.py compare legacy nodir > gpurun_out/r2_ab_cmp.txt 2>&1
python tools/knn_ab.py compare legacy coop >> gpurun_out/r2_ab_cmp.txt 2>&1
rm -f gpurun_out/knn_ab_*.pt
NMB_KNN_NO_DIR=1 $T python bench.py --steps 2 --warmup 3 --cpu-rays 0 --engine tcgen05_f16 > gpurun_out/r2_bench_nodir.txt 2>&1
$T python bench.py --steps 2 --warmup 3 --cpu-rays 0 --engine tcgen05_f16 > gpurun_out/r2_bench_coop.txt 2>&1
tail -3 gpurun_out/r2_knn_tests.txt; cat gpurun_out/r2_ab_cmp.txt; tail -c 300 gpurun_out/r2_bench_coop.txt
