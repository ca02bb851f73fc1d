python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'])"
