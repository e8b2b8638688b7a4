mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputest.log 2>&1; echo test rc=$? ; tail -4 gpurun_out/r02_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["stage_ms_per_step"], d["shard_check"]["equal"], d["e2e_entry"])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/r02_ncu_plain.json 2>/dev/null; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/r02_ncu_launch.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_digitise|k_tracks|k_gather|k_scan" --launch-skip 6 -c 18 -f -o gpurun_out/prof_r02c python bench.py --steps 2 --warmup 1 --no-entry --no-check --reference-tree none --cpu-sample 20000 > gpurun_out/r02_ncu_full.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/ | tail -12
