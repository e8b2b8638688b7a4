mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29601 tools/d2h_probe_ranks.py 2>/dev/null | tail -1 | tee gpurun_out/r2_d2h8.json
$TR --master-port 29602 tools/d2h_probe_ranks.py --bind 2>/dev/null | tail -1 | tee gpurun_out/r2_d2h8_bound.json
$TR --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_g8.json 2> gpurun_out/r2_g8.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_g8.json'))
print("g8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["shard_check"]["equal"], d["shard_check"]["rows"])
PY
$TR --master-port 29604 bench.py --gpus 8 --config c5 --warmup 3 > gpurun_out/r2_c5_g8.json 2> gpurun_out/r2_c5_g8.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_c5_g8.json'))
print("c5 g8", {k: d[k] for k in ("value","wall_s","device_s","rows_in_hdf5","steps")}, d["shard_check"]["equal"], d["shard_check"]["rows"], d["shard_check"]["how"])
PY
tail -2 gpurun_out/r2_c5_g8.err
nvidia-smi topo -m | head -14; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
