set -x
O=gpurun_out/r02/final4; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
timeout 300 python -m pytest tests/test_gpu_bootstrap.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > $O/bench_bootstrap64.json 2> $O/bench_bootstrap64.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final4/bench_*.json')):
    d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
    r=d.get('roofline',{})
    print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),3), r.get('chosen'), round(r.get('frac',0),3), r.get('traffic'), json.dumps(d.get('stage_ms',''))[:300])
PY
