mkdir -p gpurun_out
python -m pytest tests/test_gpu_compact.py tests/test_expand_hits.py -x -q 2>&1 | tail -5
python tools/time_variants.py tools/stage_time.py libm lean 2>&1 | grep -v Warn
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02b_gputest.log 2>&1; echo test rc=$? ; tail -4 gpurun_out/r02b_gputest.log
python bench.py --no-entry --cpu-sample 100000 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["stage_ms_per_step"], d["shard_check"]["equal"])
PY
