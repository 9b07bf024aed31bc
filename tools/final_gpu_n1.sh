#!/bin/bash
# Round-end evidence on ONE B200: GPU test suite, the bench line, both reference arms, launch list and ncu captures.
# Usage: gpurun --timeout 1500 -- 'bash tools/final_gpu_n1.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_scalar.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 --ref-kind vector > gpurun_out/r02_bench_reference_vector.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_2000regions.csv python bench.py --steps 2 --warmup 1 --regions 2000 --no-extras --no-cpu-baseline > /dev/null 2>&1
R=/tmp/ncu; mkdir -p $R
ncu --set full --clock-control none --import-source on -k regex:pairhmm_band_kernel -c 1 -o $R/band -f python tools/quick_bench_long.py 111 10000 12000 > /dev/null 2>&1
ncu -i $R/band.ncu-rep --page details > gpurun_out/r02_pairhmm_band_k16_ncu_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairhmm_fast_kernel -c 1 -o $R/fast -f python tools/quick_bench.py child 1000 > /dev/null 2>&1
ncu -i $R/fast.ncu-rep --page details > gpurun_out/r02_pairhmm_fast_k19_ncu_details.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_kernel -c 1 -o $R/sw -f python tools/quick_bench_sw.py 20000 > /dev/null 2>&1
ncu -i $R/sw.ncu-rep --page details > gpurun_out/r02_sw_kernel_k17_ncu_details.txt 2>&1
ls -la gpurun_out | tail -12
