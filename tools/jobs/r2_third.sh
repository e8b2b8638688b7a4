mkdir -p gpurun_out
python -m pytest tests/test_gpu_compact.py -x -q 2>&1 | tail -5
python tools/stage_time.py
python tools/time_variants.py tools/stage_time.py nobig
