set -x
O=gpurun_out/r02/final3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/test_gpu.log; tail -4 $O/test_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 300 python bench.py --workload sparse64 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_sparse64.json 2> $O/bench_sparse64.err
CB_LM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pt_pass|pt_backsub|resjac_kernel|schur_syrk" -c 8 -o $O/prof_full python profiles/prof_solve.py cfg4 1 > $O/prof_full.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_sparse64.csv python profiles/prof_solve.py sparse64 2 > $O/prof_sparse64.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final3/bench_*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        r=d.get('roofline',{})
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),3), r.get('chosen'), round(r.get('frac',0),3), r.get('traffic'), d.get('parity',{}).get('abs_diff_px'))
    except Exception as e: print(f, 'ERR', e)
PY
