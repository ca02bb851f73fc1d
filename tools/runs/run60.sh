python -m pytest tests/test_matcher_gpu.py -x -q -k "distinctive or sim3" 2>&1 | tail -3
timeout 1000 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_matcher_gpu.py tests/test_stereo_gpu.py tests/test_poseopt_gpu.py tests/test_vocabulary_gpu.py tests/test_extractor_gpu.py -x -q > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -10
