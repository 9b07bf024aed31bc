ncu --set full --clock-control none --import-source on -k regex:pairhmm_band_kernel -c 1 -o gpurun_out/r02_band_k15_v2 -f python tools/quick_bench_long.py 111 10000 12000 > gpurun_out/ncu_band.log 2>&1
ncu -i gpurun_out/r02_band_k15_v2.ncu-rep --page details > gpurun_out/r02_band_k15_v2_details.txt 2>&1
tail -3 gpurun_out/ncu_band.log
