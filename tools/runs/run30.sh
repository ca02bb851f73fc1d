ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_v12.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms']['extract_match_ms'], d['phase_ms']['local_ba_batch_ms'])"
