mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compact.py -x -q 2>&1 | tail -15
timeout 300 python tools/stage_time.py
