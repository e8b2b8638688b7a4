mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02b_gputest.log 2>&1; echo test rc=$? ; tail -4 gpurun_out/r02b_gputest.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["bytes_per_electron"], d["roofline"]["stage_ms_per_step"], d["shard_check"]["equal"], d["e2e_entry"])
PY
df -h /tmp | tail -1; free -g | head -2; python tools/entry_time.py 2e6 2 2>&1 | tail -3; timeout 300 python tools/entry_time.py 5e7 1 2>&1 | tail -3
