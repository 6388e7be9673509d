#!/bin/bash
# Same-box A/B of two builds of libdurf_b200.so (the pool's boxes differ by +-2-3 %, more than most kernel changes).
#
#   1. build the baseline, keep it:   make -C durf_b200/csrc && cp durf_b200/libdurf_b200.so durf_b200/_ab_old.so
#   2. change the kernel, rebuild:    make -C durf_b200/csrc
#   3. one GPU call, alternating:     gpurun --timeout 900 -- 'bash tools/ab_run.sh durf_b200/_ab_old.so'
#
# DURF_B200_LIB (durf_b200/_lib.py) selects the library a process loads; *.so files travel to the GPU box with the snapshot
# and are git-ignored.  Prints, for new / old / new / old: the render rate (rays resident in HBM) and the train step.
old=${1:?path of the baseline library}
for v in new old new old; do
  if [ "$v" = old ]; then export DURF_B200_LIB=$(realpath "$old"); else unset DURF_B200_LIB; fi
  echo "== $v"
  timeout 200 python bench.py --no-train --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
  timeout 120 python bench.py --train-only --train-steps 20 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
