python -m pytest tests/test_vocabulary_gpu.py tests/test_matcher_gpu.py -x -q 2>&1 | tail -8
