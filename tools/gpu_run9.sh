set -x
O=gpurun_out/r02/final
mkdir -p $O
python -c "import sys; sys.path.insert(0,'.'); from bench import kernel_source_hash; print(kernel_source_hash())" > $O/csrc_hash.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/test_gpu.log; tail -4 $O/test_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > $O/e2e_profile.log 2>&1; tail -9 $O/e2e_profile.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
for w in cfg2 cfg3 cfg4_shard8 sparse64 cfg4_intrinsics; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}.json 2> $O/bench_${w}.err
done
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err
timeout 300 python bench.py --workload triangulate_cfg4 --steps 10 --warmup 3 > $O/bench_triangulate.json 2> $O/bench_triangulate.err
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > $O/bench_bootstrap64.json 2> $O/bench_bootstrap64.err
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg4.csv python profiles/prof_solve.py cfg4 2 > $O/prof_cfg4.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg2.csv python profiles/prof_solve.py cfg2 2 > $O/prof_cfg2.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_shard8.csv python profiles/prof_solve.py cfg4_shard8 2 > $O/prof_shard8.log 2>&1
CB_LM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pt_pass|pt_backsub|resjac_kernel|schur_syrk" -c 8 -o $O/prof_full python profiles/prof_solve.py cfg4 1 > $O/prof_full.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final/bench_*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        r=d.get('roofline',{})
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d.get('e2e',{}).get('ms_per_step',0),3), r.get('chosen'), round(r.get('frac',0),3), 'hbm', round(d.get('roofline_hbm',{}).get('frac',0),3), round(d.get('roofline_hbm',{}).get('avg_launch_ms',0),4), 'sy', round(d.get('roofline_tensor',{}).get('avg_launch_ms',0),4), d.get('parity',{}).get('abs_diff_px'), d.get('nfev_per_step'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
