#!/bin/bash
# round-2 profile set (run on the GPU box): launch list of the bench command + full captures of the tile kernel (all 8 levels)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile -s 8 -c 8 -f -o gpurun_out/r2_prof_k_tile_all python tools/extract_time.py 320 2 > gpurun_out/ncu_tile_all.log 2>&1
tail -1 gpurun_out/ncu_tile_all.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_local_ba -s 1 -c 1 -f -o gpurun_out/r2_prof_k_local_ba python tools/ba_prof2.py > gpurun_out/ncu_ba.log 2>&1
tail -1 gpurun_out/ncu_ba.log
