set -x
O=gpurun_out/r02/final2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/test_gpu.log; tail -4 $O/test_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
for w in cfg2 cfg3 cfg4_shard8 sparse64 cfg4_intrinsics; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}.json 2> $O/bench_${w}.err
done
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err
timeout 300 python bench.py --workload triangulate_cfg4 --steps 10 --warmup 3 > $O/bench_triangulate.json 2> $O/bench_triangulate.err
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > $O/bench_bootstrap64.json 2> $O/bench_bootstrap64.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final2/bench_*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        r=d.get('roofline',{})
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),3), r.get('chosen'), round(r.get('frac',0),3), r.get('traffic'), d.get('parity',{}).get('abs_diff_px'), json.dumps(d.get('stage_ms',''))[:300])
    except Exception as e: print(f, 'ERR', e)
PY
