set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02/test_all3.log; tail -4 gpurun_out/r02/test_all3.log
for w in cfg4 cfg2 cfg4_shard8 cfg3 sparse64 cfg4_intrinsics; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench3_${w}.json 2> gpurun_out/r02/bench3_${w}.err
done
CB_PCG_MODE=1 timeout 300 python bench.py --workload cfg4_intrinsics --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench3_cfg4_intrinsics_pcgglobal.json 2> gpurun_out/r02/bench3_cfg4_intrinsics_pcgglobal.err
CB_SY_SPARSE=0 CB_CAM_ORDER=0 timeout 300 python bench.py --workload sparse64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench3_sparse64_dense.json 2> gpurun_out/r02/bench3_sparse64_dense.err
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02/bench3_cfg5.json 2> gpurun_out/r02/bench3_cfg5.err
timeout 300 python bench.py --workload triangulate_cfg4 --steps 10 --warmup 3 > gpurun_out/r02/bench3_triangulate.json 2> gpurun_out/r02/bench3_triangulate.err
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > gpurun_out/r02/bench3_bootstrap64.json 2> gpurun_out/r02/bench3_bootstrap64.err
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches3_cfg4.csv python profiles/prof_solve.py cfg4 2 > gpurun_out/r02/prof3_cfg4.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches3_sparse64.csv python profiles/prof_solve.py sparse64 2 > gpurun_out/r02/prof3_sparse64.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches3_cfg4i.csv python profiles/prof_solve.py cfg4_intrinsics 2 > gpurun_out/r02/prof3_cfg4i.log 2>&1
CB_LM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pt_pass|pt_backsub|resjac_kernel|schur_syrk" -c 8 -o gpurun_out/r02/prof3_full python profiles/prof_solve.py cfg4 1 > gpurun_out/r02/prof3_full.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench3_*.json')):
    try:
        d=json.load(open(f))
        if 'roofline_tensor' in d:
            print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['lm_loop']['trial_replay'][:12], 'pp', round(d['roofline']['avg_launch_ms'],4), round(d['roofline']['frac'],3), 'sy', round(d['roofline_tensor']['avg_launch_ms'],4), round(d['roofline_tensor']['frac'],3), d.get('parity',{}).get('abs_diff_px'), d['nfev_per_step'])
        else:
            print(f, round(d['value'],1), round(d['ms_per_step'],3), json.dumps(d.get('stage_ms', d.get('stages','')))[:300], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02/bench3_*.err | tail -40
