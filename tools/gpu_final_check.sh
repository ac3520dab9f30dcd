set -x
O=gpurun_out/r02/final5; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/test_gpu.log; tail -3 $O/test_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; tail -c 400 $O/bench_cfg4.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02/final5/bench_cfg4.json') if l.startswith('{')][-1]
r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), r['chosen'], round(r['frac'],3), r['traffic'], d['parity']['abs_diff_px'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches'])
PY
