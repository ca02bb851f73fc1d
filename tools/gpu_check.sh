#!/bin/bash
# Round-end style check on a B200 box (run through gpurun): GPU parity suite, smoke(), default bench line, ncu launch list.
set -u
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench_err.log && tail -c 600 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
