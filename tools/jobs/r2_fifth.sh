mkdir -p gpurun_out
python tools/stage_time.py
python tools/time_variants.py tools/stage_time.py p4
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
