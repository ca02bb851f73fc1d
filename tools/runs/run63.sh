python -m pytest tests/test_extractor_gpu.py -x -q -k "batch_upload" 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
