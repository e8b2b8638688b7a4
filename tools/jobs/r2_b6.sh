mkdir -p gpurun_out
python -m pytest tests/test_gpu_frontend.py tests/test_gpu_compact.py -x -q 2>&1 | tail -3
python tools/entry_profile.py 2e7 2>&1 | grep -v "INFO\|Warn"
python tools/entry_time.py 2e6 3 2>&1 | grep -v INFO | tail -4
timeout 300 python tools/entry_time.py 1e8 2 2>&1 | grep -v INFO | tail -3
