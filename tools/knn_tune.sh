#!/bin/bash
# GPU experiment: KNN kernel time per frame for a few occupancy / leaf-size variants (single process each)
for cfg in "8 16" "10 16" "12 16" "8 8" "8 32" "10 32"; do
  set -- $cfg
  echo "MINB=$1 LEAF=$2" >> gpurun_out/knn_tune.txt
  NMB_KNN_MINB=$1 NMB_LEAF_MAX=$2 timeout 120 python bench.py --simulate-world 1 --steps 2 --warmup 2 --cpu-rays 0 --tune 2>/dev/null | grep diagnostic >> gpurun_out/knn_tune.txt
done
