set -x
python bench.py --regions 500 --cfg5-regions 3000 --cfg3-regions 8 --steps 2 --warmup 1 --cpu-sample-regions 16 > gpurun_out/bench_small_n1.json 2> gpurun_out/bench_small_n1.err; tail -3 gpurun_out/bench_small_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --regions 500 --cfg5-regions 3000 --steps 2 --warmup 1 > gpurun_out/bench_small_n2.json 2> gpurun_out/bench_small_n2.err; tail -3 gpurun_out/bench_small_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config cfg5 --cfg5-regions 3000 --steps 2 --warmup 1 > gpurun_out/bench_small_n2_cfg5.json 2> gpurun_out/bench_small_n2_cfg5.err; tail -3 gpurun_out/bench_small_n2_cfg5.err
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
