mkdir -p gpurun_out
python -m pytest tests/test_gpu_compact.py -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-entry --reference-tree none --cpu-sample 100000 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b2.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["k_tracks_launch_ms"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/ncu1.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>5]
for r in rows[-40:]:
    print(r[4][:70], r[-1])
PY
