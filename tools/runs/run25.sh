python -m pytest tests/test_stream_gpu.py tests/test_extractor_gpu.py tests/test_matcher_gpu.py -x -q 2>&1 | tail -4
python tools/e2e_time.py 2>&1 | tail -3
for nc in 4 3 2; do B2S_BA_NCTA=$nc python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('ncta $nc', d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'])"; done
