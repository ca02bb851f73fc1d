#!/bin/bash
# full ncu capture of the level-0 k_tile launch of a 320-image batch -> gpurun_out/$1.ncu-rep
tag=${1:-r2_prof_k_tile}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tile -s 8 -c 1 -f -o gpurun_out/$tag python tools/extract_time.py 320 2 > gpurun_out/ncu_tile.log 2>&1
tail -2 gpurun_out/ncu_tile.log
