#!/bin/bash
# ncu --set full capture of one kernel of the bench step (run through gpurun): tools/profile_kernel.sh k_fast_cells
# LocalBA phases: B2S_BA_REPEAT=$((phase*256+8)) tools/profile_kernel.sh k_local_ba   (phase 1 Schur, 2 build, 3 errors)
k=${1:-k_fast_cells}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
tail -1 gpurun_out/ncu_$k.log
