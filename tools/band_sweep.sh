#!/bin/bash
# developer sweep: band kernel (long reads) build variants lorikeet_b200/_variants/lib_band_*.so
N=${1:-111}
mkdir -p gpurun_out
echo "== default" ; python tools/quick_bench_long.py $N 10000 12000 check 2>&1 | tail -2
for lib in lorikeet_b200/_variants/lib_band_*.so; do
  echo "== $lib"
  LK_SO=$PWD/$lib python tools/quick_bench_long.py $N 10000 12000 2>&1 | tail -1
done
