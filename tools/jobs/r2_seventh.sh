mkdir -p gpurun_out
python tools/stage_time.py
PTB_KS_N=100000 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
