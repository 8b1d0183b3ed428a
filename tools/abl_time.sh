#!/bin/bash
# developer tool: time the dense stages at N=1024 with the product library and every ablation build present
for lib in fk_mc_b200/lib fk_mc_b200/lib_*; do
  [ -f $lib/libfkmc_b200.so ] || continue
  echo "== $lib"
  FKMC_LIB=$lib/libfkmc_b200.so timeout 180 python tools/dense_time.py 32 592 2>&1 | head -1
done
