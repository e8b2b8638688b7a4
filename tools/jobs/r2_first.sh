mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo rc=$?; tail -c 3000 gpurun_out/r2_b1.json; tail -5 gpurun_out/r2_b1.err
