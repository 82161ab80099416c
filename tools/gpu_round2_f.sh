mkdir -p gpurun_out
T="timeout 200"
$T python tools/knn_ab.py dump sort > gpurun_out/r2f_ab_sort.txt 2>&1
NMB_KNN_ORDER=1 $T python tools/knn_ab.py dump nearest > gpurun_out/r2f_ab_nearest.txt 2>&1
python tools/knn_ab.py compare sort nearest > gpurun_out/r2f_ab_cmp.txt 2>&1
rm -f gpurun_out/knn_ab_*.pt
NMB_KNN_ORDER=1 $T python bench.py --steps 3 --warmup 3 --cpu-rays 0 > gpurun_out/r2f_bench_nearest.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --cpu-rays 0 > gpurun_out/r2f_bench_under_ncu.txt 2>&1
grep AB gpurun_out/r2f_ab_cmp.txt; tail -c 300 gpurun_out/r2f_bench_nearest.txt; wc -l gpurun_out/r2_launches_bench.csv
