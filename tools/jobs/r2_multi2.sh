mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_g2.json 2> gpurun_out/r2_g2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_g2.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["shard_check"]["equal"], d["shard_check"]["rows"], d["shard_check"]["how"])
PY
tail -3 gpurun_out/r2_g2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --electrons 4e8 --warmup 3 > gpurun_out/r2_c5_g2.json 2> gpurun_out/r2_c5_g2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_c5_g2.json'))
print({k: d[k] for k in ("value","wall_s","device_s","rows_in_hdf5","steps")}, d["shard_check"]["equal"], d["shard_check"]["rows"])
PY
tail -3 gpurun_out/r2_c5_g2.err
python bench.py --config c5 --electrons 4e8 --warmup 3 > gpurun_out/r2_c5_g1.json 2> gpurun_out/r2_c5_g1.err; echo rc=$?
python - <<'PY'
import json
a=json.load(open('gpurun_out/r2_c5_g1.json')); b=json.load(open('gpurun_out/r2_c5_g2.json'))
print("c5 1 GPU", a["value"], a["wall_s"], a["shard_check"]["equal"])
print("digests equal 1 vs 2 GPUs:", a["shard_check"]["sharded"] == b["shard_check"]["sharded"])
PY
