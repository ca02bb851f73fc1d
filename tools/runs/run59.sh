python -m pytest tests/test_matcher_gpu.py -x -q -k "distinctive or sim3" 2>&1 | tail -4
