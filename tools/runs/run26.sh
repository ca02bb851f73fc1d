python -m pytest tests/test_localba_gpu.py tests/test_stream_gpu.py -x -q 2>&1 | tail -4
python tools/ba_prof.py 32 3 2>&1 | grep -E "Mcycles|batch" | tail -3 | cut -c1-330
