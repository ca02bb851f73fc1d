python -m pytest tests/test_poseopt_gpu.py tests/test_localba_gpu.py -x -q 2>&1 | tail -12
