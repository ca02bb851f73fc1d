python -m pytest tests/test_shim_gpu.py tests/test_vocabulary_gpu.py -x -q 2>&1 | tail -6
