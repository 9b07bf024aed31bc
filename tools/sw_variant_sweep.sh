#!/bin/bash
# developer sweep over Smith-Waterman kernel build variants (lorikeet_b200/_variants/lib_sw_*.so) on the cfg#4 generator
N=${1:-100000}
for lib in default $(ls lorikeet_b200/_variants/ 2>/dev/null | grep '^lib_sw_' | sed 's/^lib_sw_//; s/\.so$//'); do
  if [ "$lib" = default ]; then unset LK_SO; else export LK_SO=$PWD/lorikeet_b200/_variants/lib_sw_$lib.so; fi
  echo "== variant $lib"
  timeout 120 python tools/quick_bench_sw.py $N 2>&1 | grep GCUPS
done
