set -x
O=gpurun_out/r02/final6; mkdir -p $O
timeout 200 python bench.py --workload cfg4_shard8 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg4_shard8.json 2> $O/bench_cfg4_shard8.err; tail -c 300 $O/bench_cfg4_shard8.err
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final6/bench_*.json')):
    d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
    print(f.split('/')[-1], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'bracket', round(d['bracket_ms_per_step'],4), 'wall', round(d['wall_ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['config']['l2'][:60])
PY
