mkdir -p gpurun_out
python tools/stage_time.py 2>&1 | grep -v Warn
python -m pytest tests/test_gpu_compact.py tests/test_gpu_frontend.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
python tools/entry_time.py 2e6 3 2>&1 | grep -v INFO | tail -4
timeout 300 python tools/entry_time.py 5e7 1 2>&1 | grep -v INFO | tail -3
python bench.py --no-check --cpu-sample 100000 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["stage_ms_per_step"], d["e2e_entry"])
PY
