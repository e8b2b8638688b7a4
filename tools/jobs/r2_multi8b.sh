mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29603 bench.py --gpus 8 > gpurun_out/r2_g8.json 2> gpurun_out/r2_g8.err; echo rc=$?
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
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR4 --master-port 29605 bench.py --gpus 4 > gpurun_out/r2_g4.json 2> gpurun_out/r2_g4.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_g4.json'))
print("g4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["shard_check"]["equal"], d["shard_check"]["rows"])
PY
