#!/bin/bash
# developer sweep over kernel build variants (lorikeet_b200/_variants/lib_*.so) and K
N=${1:-3000}
for lib in default u2 u8 pf u2pf; do
  if [ "$lib" = default ]; then unset LK_SO; else export LK_SO=$PWD/lorikeet_b200/_variants/lib_$lib.so; fi
  echo "== variant $lib"
  python tools/quick_bench.py $N s10,s15,s19 2>&1 | grep GCUPS
done
