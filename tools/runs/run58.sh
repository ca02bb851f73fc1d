python -m pytest tests/test_shim_gpu.py tests/test_stream_gpu.py -x -q 2>&1 | tail -3
for cfg in "4 1" "4 2" "2 1" "2 2" "3 1"; do set -- $cfg; B2S_BA_NCTA=$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ba-depth $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('ncta $1 depth $2:', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), round(d['phase_ms']['local_ba_batch_ms'],2))"; done
