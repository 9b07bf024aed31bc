python -m pytest tests/test_gpu_pool.py tests/test_gpu_batcher.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for d in 1 2; do tools/native/bench_batcher 4000 64 512 200 1 $d | tail -3; tools/native/bench_batcher 4000 16 512 200 16 $d | tail -1; done
