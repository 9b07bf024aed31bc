#!/bin/bash
# Round-end evidence on an 8-GPU box: the bench line at N = 8 (weak headline + cfg5 strong, gather verified), the in-process
# pool (lk_pool) through the native harness at 1/2/4/8 devices, and smoke() with every device visible.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/final_gpu_n8.sh'
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err
{
for d in 1 2 4 8; do
  [ $d -le $N ] || continue
  tools/native/bench_batcher 8000 64 512 200 1 $d | tail -1
  tools/native/bench_batcher 8000 256 512 200 1 $d | tail -1
  tools/native/bench_batcher 8000 32 512 200 16 $d | tail -1
  tools/native/bench_batcher 16000 64 512 200 64 $d | tail -1
done
} > gpurun_out/r02_pool_batcher_n$N.jsonl 2>&1
cat gpurun_out/r02_pool_batcher_n$N.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02_smoke_n$N.txt
