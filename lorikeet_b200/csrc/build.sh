#!/bin/bash
# Build liblorikeet_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=${LK_OUT:-../liblorikeet_b200.so}
BUILD=${LK_BUILD_DIR:-../_build}
NVCC=${NVCC:-nvcc}
FLAGS="${LK_EXTRA_FLAGS} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function,-Wno-unknown-pragmas"
mkdir -p $BUILD
pids=()
for f in pairhmm_kernels.cu sw_kernels.cu engine_kernels.cu lk_api.cu lk_sw_api.cu tools_kernels.cu; do
  [ -f "$f" ] || continue
  o=$BUILD/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o" -print -quit)" ] || [ ../../include/lorikeet_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" & pids+=($!)
  fi
done
for f in host_model.cpp host_prep.cpp host_preops.cpp host_engine.cpp host_cigar.cpp host_batcher.cpp host_pool.cpp; do
  [ -f "$f" ] || continue
  o=$BUILD/${f%.cpp}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.h' -newer "$o" -print -quit)" ] || [ ../../include/lorikeet_b200.h -nt "$o" ]; then
    g++ -O3 -std=c++17 -fPIC -Wall -Wno-unknown-pragmas -I/usr/local/cuda/include -c "$f" -o "$o" & pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
rm -f $BUILD/host_synth.o  # generators moved to their own library
$NVCC -shared -o $OUT $BUILD/*.o -Xcompiler -fPIC -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
# synthetic-workload generators (bench / tests only): a separate library with no CUDA dependency, so the reference arm
# of bench.py can generate inputs without mapping the product library
SYNTH_OUT=${LK_SYNTH_OUT:-../liblorikeet_b200_synth.so}
if [ ! -f "$SYNTH_OUT" ] || [ host_synth.cpp -nt "$SYNTH_OUT" ] || [ host_preops.cpp -nt "$SYNTH_OUT" ] || [ ../../include/lorikeet_b200_synth.h -nt "$SYNTH_OUT" ]; then
  g++ -O3 -std=c++17 -fPIC -shared -Wall -Wno-unknown-pragmas -o "$SYNTH_OUT" host_synth.cpp host_preops.cpp -lpthread
  echo "built $SYNTH_OUT"
fi
# native harness of the drop-in call pattern (tools/native/bench_batcher.cpp); bench.py runs it for other_configs.batcher
BB=../../tools/native/bench_batcher
if [ -z "$LK_OUT" ] && { [ ! -f "$BB" ] || [ ../../tools/native/bench_batcher.cpp -nt "$BB" ] || [ "$OUT" -nt "$BB" ]; }; then
  g++ -O2 -std=c++17 -I../../include ../../tools/native/bench_batcher.cpp -o "$BB" -L"$(dirname "$OUT")" -l:"$(basename "$OUT")" -l:"$(basename "$SYNTH_OUT")" -lpthread -Wl,-rpath,'$ORIGIN/../../lorikeet_b200' && echo "built $BB"
fi
