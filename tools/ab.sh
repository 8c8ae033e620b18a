#!/bin/bash
# A/B of library variants: tools/ab.sh lib0 libA ...  (files under dfine_b200/_C/variants/)
for v in "$@"; do
  DFINE_B200_LIB=$PWD/d-fine-seg_b200/dfine_b200/_C/variants/$v.so python bench.py --no-cpu-baseline --no-eager --no-secondary --steps 20 > gpurun_out/ab_$v.log 2>&1
  echo "== $v"; python tools/bench_summary.py gpurun_out/ab_$v.log
done
