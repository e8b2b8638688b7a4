mkdir -p gpurun_out
python -m pytest tests/test_gpu_compact.py -x -q 2>&1 | tail -12
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-entry --no-check --reference-tree none --cpu-sample 20000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sampler on ', d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e']['value'])"
PTB_BENCH_NO_SAMPLER=1 python bench.py --steps 20 --warmup 5 --no-entry --no-check --reference-tree none --cpu-sample 20000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sampler off', d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['e2e']['value'])"
done
