python -m pytest tests/test_localba_gpu.py -x -q 2>&1 | tail -2
python tools/ba_prof.py 32 3 2>&1 | grep -E "Mcycles|device" | tail -2 | cut -c1-330
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms']['extract_match_ms'], d['phase_ms']['local_ba_batch_ms'])"
